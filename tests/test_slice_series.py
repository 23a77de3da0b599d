"""The slice exponential of the smush kernels (csrc/slam_fwd1.cuh, herm_expm) restated in numpy: exp(-i A) = cos A - i sin A
for Hermitian A with ||A|| <= 0.7 from the degree-16/17 Taylor series, evaluated as Horner recurrences in K^2 = A^4 with
linear blocks (Paterson-Stockmeyer, block size 2: 9 products of commuting Hermitian matrices), then scaling-and-squaring.
Replaces qutip's Qobj.expm (scipy Pade) per time slice (src/slam/hamiltonian.py:136-143); checked here against
scipy.linalg.expm so that the coefficient table and the recurrence the kernel uses are pinned without a GPU."""
import math

import numpy as np
from scipy.linalg import expm

A_COEF = [(-1) ** j / math.factorial(2 * j) for j in range(9)]            # cos:      sum a_j K^j,  K = A^2
B_COEF = [(-1) ** j / math.factorial(2 * j + 1) for j in range(8)]        # sin A / A: sum b_j K^j


def herm_expm(H, dt, rho):
    """The kernel's algorithm: theta = dt / 2^s with ||theta H|| <= 0.7 (rho = the kernel's norm bound), series, s squarings."""
    s = 0
    if rho > 0.7:
        s = min(math.frexp(rho / 0.7)[1] - 1 + 1, 40)  # ilogb(rho / 0.7) + 1
    A = (dt / 2.0 ** s) * H
    I = np.eye(4)
    K = A @ A
    K2 = K @ K
    C = A_COEF[6] * I + A_COEF[7] * K + A_COEF[8] * K2
    P = B_COEF[6] * I + B_COEF[7] * K
    for i in (2, 1, 0):
        C = K2 @ C + A_COEF[2 * i] * I + A_COEF[2 * i + 1] * K
        P = K2 @ P + B_COEF[2 * i] * I + B_COEF[2 * i + 1] * K
    Y = C - 1j * (A @ P)
    for _ in range(s):
        Y = Y @ Y
    return Y


def _smush_generator(rng, scale):
    """General slice generator of hamiltonian.py:114-182: h13 = h02, h23 = h01, three diagonal Z terms."""
    gx, gy, gc, gg, gz1, gz2 = rng.uniform(-scale, scale, 6)
    pa, pb, pc, pg = rng.uniform(-np.pi, np.pi, 4)
    H = np.zeros((4, 4), dtype=complex)
    H[0, 1] = H[2, 3] = gy * np.exp(-1j * pb)
    H[0, 2] = H[1, 3] = gx * np.exp(-1j * pa)
    H[0, 3] = gg * np.exp(-1j * pg)
    H[1, 2] = gc * np.exp(-1j * pc)
    H = H + H.conj().T
    H[0, 0], H[1, 1], H[2, 2] = gz1 + gz2, gz1, gz2
    rho = abs(gx) + abs(gy) + max(abs(gc), abs(gg)) + abs(gz1) + abs(gz2)
    return H, rho


def test_series_at_the_scaling_threshold():
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(300):
        M = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        A = (M + M.conj().T) / 2
        A *= 0.7 / np.linalg.norm(A, 2)
        worst = max(worst, np.abs(herm_expm(A, 1.0, 0.7) - expm(-1j * A)).max())
    assert worst < 2e-15, worst


def test_scaling_and_squaring_over_the_reference_sampling_box():
    """Amplitudes of the reference's sampling box U(-4 pi, 4 pi) at dt = 0.25 (||dt H|| up to ~20): the result stays unitary and
    within 1e-13 of scipy's expm (north_star: template unitaries within 1e-10)."""
    rng = np.random.default_rng(1)
    for scale in (0.1, 1.0, 4 * np.pi):
        for _ in range(100):
            H, rho = _smush_generator(rng, scale)
            dt = 0.25
            Y = herm_expm(H, dt, dt * rho)
            assert np.abs(Y - expm(-1j * dt * H)).max() < 1e-13
            assert np.abs(Y.conj().T @ Y - np.eye(4)).max() < 1e-13


def test_norm_bound_covers_the_spectral_norm():
    """rho of smush_slice bounds ||H||_2 (row-sum bound of the generator's structure), so ||theta H|| <= 0.7 really holds."""
    rng = np.random.default_rng(2)
    for _ in range(500):
        H, rho = _smush_generator(rng, 5.0)
        assert np.linalg.norm(H, 2) <= rho * (1 + 1e-12)

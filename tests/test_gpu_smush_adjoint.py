"""Analytic adjoint gradient through parameter-bound smush gates (slam_adj1.cuh): K2 `slam_loss_grad` on GM_SMUSH
templates and the adjoint mode of the K5c optimiser.  Replaces scipy's (P+1)-evaluation finite-difference gradient
(opt.minimize(jac=None), src/slam/optimizer.py:270-278) for the templates of parallel_drive_volume.py:175-199 and
pd_playground.py:60-97 (gates: hamiltonian.py:114-182).

Oracle gradient: closed-form U3/RZ derivative matrices + Richardson-extrapolated central differences of the scipy-expm
gate matrix only (oracle.loss_and_grad), independent of the kernel's eigen-decomposition route.
Tolerances: loss 1e-12, gradient 1e-9 absolute (north_star: template unitaries within 1e-10)."""
import zlib

import numpy as np
import pytest
import torch

import oracle as O
from helpers import BASES, make_pair
from slam_decomposition_b200 import _lib, engine
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

pytestmark = pytest.mark.gpu

CASES = {
    # the Monte-Carlo / training template of parallel_drive_volume.py: (phi_c, phi_g, gx[T], gy[T]) free, no exterior 1Q
    "smush_sqiSwap_k2": dict(kind="smush", slots=("Q", "Q", np.pi / 2, 0.0, "Q", "Q", "Q", "Q", 0.5), k=2, T=2,
                             no_exterior_1q=True),
    "smush_CNOT_k1_T4": dict(kind="smush", slots=("Q", "Q", np.pi / 4, np.pi / 4) + ("Q",) * 8 + (1.0,), k=1, T=4,
                             no_exterior_1q=True),
    # every slot free, incl. the couplings and the duration, with exterior layers
    "smush_all_free_k2": dict(kind="smush", slots=("Q",) * 9, k=2, T=2),
    # ConversionGainSmush1QPhase (pd_playground): drive phases, Z terms and duration free
    "smush1q_all_free_k1_T3": dict(kind="smush1q", slots=("Q",) * 15, k=1, T=3),
    "smush1q_mixed_k3_T1": dict(kind="smush1q", slots=(0.3, "Q", -0.4, "Q", "Q", 0.7, "Q", 0.2, "Q", "Q", 0.1), k=3, T=1),
    # RZ-only 1Q layers (basisv2.py:267-272)
    "smush_vz_k2": dict(kind="smush", slots=("Q", "Q", np.pi / 2, 0.0, "Q", "Q", "Q", "Q", 0.5), k=2, T=2, vz_only=True),
}


def _inputs(orc, B, seed, span=2.0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-span, span, (B, orc.n_params))
    V = np.stack([O.haar_unitary(rng) for _ in range(3)])
    return X, V


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("cost", ["basic", "square", "basic_inverse"])
def test_loss_grad_matches_oracle(name, cost):
    desc, orc = make_pair(**CASES[name])
    B = 6
    X, V = _inputs(orc, B, seed=zlib.crc32(name.encode()) % 1000)
    ck = {"basic": _lib.COST_BASIC, "square": _lib.COST_SQUARE, "basic_inverse": _lib.COST_BASIC_INVERSE}[cost]
    loss, grad, tr = engine.loss_grad(desc, torch.as_tensor(X, device="cuda"), torch.as_tensor(V, device="cuda"),
                                      cost_kind=ck, want_trace=True)
    loss, grad, tr = loss.cpu().numpy(), grad.cpu().numpy(), tr.cpu().numpy()
    for b in range(B):
        l, g, T = O.loss_and_grad(orc, X[b], V[b % 3], cost, h_gate=2e-3, richardson=True)
        assert abs(loss[b] - l) < 1e-12
        assert abs(tr[b] - T) < 1e-11
        assert np.abs(grad[b] - g).max() < 1e-9, (name, b, np.abs(grad[b] - g).max())


def test_loss_only_and_large_amplitudes():
    """Loss-only launch equals the gradient launch's loss; amplitudes of the reference's sampling box U(-4 pi, 4 pi)
    (basisv2.py:157: ||dt H|| up to ~20) keep the gradient within 1e-8 of the oracle."""
    desc, orc = make_pair(**CASES["smush_sqiSwap_k2"])
    X, V = _inputs(orc, 8, seed=11, span=4 * np.pi)
    Xd, Vd = torch.as_tensor(X, device="cuda"), torch.as_tensor(V, device="cuda")
    l0, _, _ = engine.loss_grad(desc, Xd, Vd, want_grad=False)
    l1, g1, _ = engine.loss_grad(desc, Xd, Vd)
    assert (l0 - l1).abs().max().item() < 1e-13  # (series forward vs eigen-form forward of the gradient launch: rounding only)
    for b in range(8):
        l, g, _ = O.loss_and_grad(orc, X[b], V[b % 3], "basic", h_gate=1e-3, richardson=True)
        assert abs(l1[b].item() - l) < 1e-12
        assert np.abs(g1[b].cpu().numpy() - g).max() < 1e-8


def test_degenerate_slices():
    """Zero drives and zero phases: H has repeated eigenvalues (the divided differences take their sinc limit) and the
    Jacobi sweep starts from an already block-diagonal matrix."""
    desc, orc = make_pair(**CASES["smush_sqiSwap_k2"])
    X = np.zeros((2, orc.n_params))
    X[1, :6] = [0.3, -0.2, 0.5, 0.1, 0.9, -0.7]
    rng = np.random.default_rng(5)
    V = np.stack([O.haar_unitary(rng)])
    loss, grad, _ = engine.loss_grad(desc, torch.as_tensor(X, device="cuda"), torch.as_tensor(V, device="cuda"))
    for b in range(2):
        l, g, _ = O.loss_and_grad(orc, X[b], V[0], "basic", h_gate=2e-3, richardson=True)
        assert abs(loss[b].item() - l) < 1e-12
        assert np.abs(grad[b].cpu().numpy() - g).max() < 1e-9


def test_gradient_is_consistent_with_the_loss_kernel():
    """Directional derivative of the DEVICE loss (central difference along a random direction) equals grad . d."""
    desc, orc = make_pair(**CASES["smush1q_all_free_k1_T3"])
    rng = np.random.default_rng(2)
    B, P = 64, orc.n_params
    X = rng.uniform(-1.5, 1.5, (B, P))
    D = rng.standard_normal((B, P))
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    V = torch.as_tensor(np.stack([O.haar_unitary(rng) for _ in range(4)]), device="cuda")
    h = 1e-5
    Xd = torch.as_tensor(X, device="cuda")
    Dd = torch.as_tensor(D, device="cuda")
    _, g, _ = engine.loss_grad(desc, Xd, V)
    lp, _, _ = engine.loss_grad(desc, Xd + h * Dd, V, want_grad=False)
    lm, _, _ = engine.loss_grad(desc, Xd - h * Dd, V, want_grad=False)
    fd = (lp - lm) / (2 * h)
    assert ((g * Dd).sum(dim=1) - fd).abs().max().item() < 1e-8


def _smush_pair(name, k):
    gc, gg, t = BASES[name]
    T = round(t / 0.25)
    basis = pdv.smush_template(gc, gg, t, k)
    basis.spanning_range = range(k, k + 1)
    orc = O.OracleTemplate("smush", ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,), k=k, T=T, no_exterior_1q=True)
    return basis, orc


def test_adjoint_optimizer_converges_in_the_basin_and_uses_fewer_evaluations():
    """K5c with the adjoint gradient: started inside the basin (x_true + 0.2 noise) every target must reach 1e-10 (the
    reference's success threshold; its own finite-difference BFGS stalls near 5e-9), the returned Xk must reproduce the
    loss in the oracle, and the evaluation count must be far below the central-difference run's."""
    basis, orc = _smush_pair("sqiSwap", 2)
    rng = np.random.default_rng(3)
    Nt, R, P = 12, 4, orc.n_params
    X_true = rng.uniform(-1.5, 1.5, (Nt, P))
    V = torch.as_tensor(np.stack([orc.eval(x) for x in X_true]), device="cuda")
    x0 = torch.as_tensor(X_true[:, None, :] + 0.2 * rng.standard_normal((Nt, R, P)), device="cuda")
    opts = engine.opt_defaults()
    opts.f_far = 1e-4
    ev_adj = torch.zeros(1, dtype=torch.int64, device="cuda")
    ev_cen = torch.zeros(1, dtype=torch.int64, device="cuda")
    loss, x, iters = engine.fd_lbfgs_solve(basis.desc, V, R, opts, x0=x0, central="adjoint", evals=ev_adj)
    loss_c, _, _ = engine.fd_lbfgs_solve(basis.desc, V, R, opts, x0=x0, central=True, evals=ev_cen)
    best = loss.min(dim=1).values.cpu().numpy()
    best_c = loss_c.min(dim=1).values.cpu().numpy()
    # Without exterior 1Q gates the Hessian at a solution is singular, so some restarts creep (the "< 3 % progress per 32
    # steps" rule ends them between 1e-10 and 5e-9 in BOTH gradient modes; measured 2026-10: 9/12 targets below 1e-10,
    # worst 1.3e-9, identical to the central-difference run).  The reference's own BFGS stops at |g| < 1e-5, i.e. 1e-6..1e-8.
    assert (best <= 1e-10).mean() >= 0.6, best
    assert (best <= 5e-9).all(), best
    assert np.median(best) <= 10 * max(np.median(best_c), 1e-13), (best, best_c)
    xs, ls, Vh = x.cpu().numpy(), loss.cpu().numpy(), V.cpu().numpy()
    for i in range(Nt):
        r = int(ls[i].argmin())
        assert abs(O.cost(orc.eval(xs[i, r]), Vh[i], "basic") - ls[i, r]) < 1e-10
    assert ev_adj.item() * 5 < ev_cen.item(), (ev_adj.item(), ev_cen.item())


def test_template_optimizer_uses_the_adjoint_by_default():
    basis, orc = _smush_pair("sqiSwap", 1)
    rng = np.random.default_rng(17)
    V = np.stack([orc.eval(rng.uniform(-1.0, 1.0, orc.n_params)) for _ in range(6)])
    np.random.seed(1)
    opt = TemplateOptimizer(basis, BasicCost(), override_fail=True, training_restarts=32)
    assert opt.smush_adjoint
    out = opt.approximate_targets(V, range(1, 2))
    assert (out["loss"] <= 1e-9).mean() >= 0.8, out["loss"]
    for i in range(6):
        assert abs(O.cost(orc.eval(out["Xk"][i, : orc.n_params]), V[i], "basic") - out["loss"][i]) < 1e-10


def test_bounded_adjoint_run_respects_the_box():
    basis, orc = _smush_pair("sqiSwap", 1)
    rng = np.random.default_rng(9)
    Nt, R, P = 4, 8, orc.n_params
    V = torch.as_tensor(np.stack([orc.eval(rng.uniform(-0.8, 0.8, P)) for _ in range(Nt)]), device="cuda")
    lo = torch.full((P,), -1.0, dtype=torch.float64, device="cuda")
    hi = torch.full((P,), 1.0, dtype=torch.float64, device="cuda")
    opts = engine.opt_defaults()
    opts.lower, opts.upper = lo.data_ptr(), hi.data_ptr()
    opts.x0_lo, opts.x0_hi = -1.0, 1.0
    loss, x, _ = engine.fd_lbfgs_solve(basis.desc, V, R, opts, seed=4, central="adjoint")
    assert x.min().item() >= -1.0 and x.max().item() <= 1.0
    assert (loss.min(dim=1).values <= 1e-9).float().mean().item() >= 0.75


def test_bounded_adjoint_run_stops_only_at_kkt_points():
    """Box-constrained K5c (adjoint mode; the reference switches scipy to L-BFGS-B, optimizer.py:257-258): a quasi-Newton
    direction whose clamped segment is not a descent segment must fall back to projected steepest descent instead of
    ending the restart (SlamOptOpts.diag exposes the stop reason: 7 = no feasible descent).  With the start box wider
    than the bounds most initial points sit on a face of the box, which is where the premature stop used to happen."""
    basis, orc = _smush_pair("sqiSwap", 1)
    rng = np.random.default_rng(21)
    Nt, R, P = 64, 8, orc.n_params
    V = torch.as_tensor(np.stack([orc.eval(rng.uniform(-0.9, 0.9, P)) for _ in range(Nt)]), device="cuda")
    lo = torch.full((P,), -1.0, dtype=torch.float64, device="cuda")
    hi = torch.full((P,), 1.0, dtype=torch.float64, device="cuda")
    opts = engine.opt_defaults()
    opts.lower, opts.upper = lo.data_ptr(), hi.data_ptr()
    opts.x0_lo, opts.x0_hi = -3.0, 3.0  # clamped into the box: many coordinates start on a bound
    opts.early_exit = 0
    opts.diag = 1
    loss, x, it = engine.fd_lbfgs_solve(basis.desc, V, R, opts, seed=5, central="adjoint")
    reason = (it.cpu().numpy().astype(np.int64) >> 24).ravel()
    assert set(np.unique(reason)) <= {1, 2, 3, 4, 7, 8, 9}, np.unique(reason)
    assert x.min().item() >= -1.0 and x.max().item() <= 1.0
    # every "no feasible descent" stop is a KKT point of the box problem: zero projected gradient
    stop7 = np.nonzero(reason == 7)[0]
    if len(stop7):
        xs = x.reshape(-1, P)[torch.as_tensor(stop7, device="cuda")].contiguous()
        _, g, _ = engine.loss_grad(basis.desc, xs, V, tgt_idx=torch.as_tensor(stop7 // R, dtype=torch.int32, device="cuda"))
        xs_, g_ = xs.cpu().numpy(), g.cpu().numpy()
        free = ~(((xs_ <= -1.0) & (g_ > 0)) | ((xs_ >= 1.0) & (g_ < 0)))
        assert np.abs(g_ * free).max() < 1e-6
    assert (reason == 7).mean() < 0.05, np.bincount(reason)
    assert (loss.min(dim=1).values <= 1e-9).float().mean().item() >= 0.75


@pytest.mark.parametrize("gc,gg,t,k,form", [
    (np.pi / 2, 0.0, 0.25, 2, "reg16"),   # P = 14
    (np.pi / 2, 0.0, 0.5, 2, "reg20"),    # P = 18
    (np.pi / 2, 0.0, 0.75, 2, "reg24"),   # P = 22
    (np.pi / 4, np.pi / 4, 1.0, 2, "reg28"),  # P = 26
    (np.pi / 2, 0.0, 0.5, 3, "reg32"),    # P = 30
    (np.pi / 4, np.pi / 4, 0.5, 4, "local-array"),  # P = 42
])
def test_every_form_of_the_adjoint_optimizer_converges_in_the_basin(gc, gg, t, k, form):
    """K5c's adjoint mode picks a kernel by parameter count -- register form with the vector bound NQ = 16 / 20 / 24 / 28 / 32
    ((x, g) buffers in shared memory up to 28), local-array form above 32.  Each must solve in-basin starts of the
    parallel-drive Monte-Carlo template (parallel_drive_volume.py:175-199) and return parameters that reproduce the loss."""
    T = round(t / 0.25)
    basis = pdv.smush_template(gc, gg, t, k)
    orc = O.OracleTemplate("smush", ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,), k=k, T=T, no_exterior_1q=True)
    P = orc.n_params
    assert basis.desc.n_params == P
    rng = np.random.default_rng(zlib.crc32(form.encode()))
    Nt, R = 24, 2
    X_true = rng.uniform(-1.2, 1.2, (Nt, P))
    V = torch.as_tensor(np.stack([orc.eval(x) for x in X_true]), device="cuda")
    x0 = torch.as_tensor(X_true[:, None, :] + 0.05 * rng.standard_normal((Nt, R, P)), device="cuda")
    opts = engine.opt_defaults()
    opts.f_far = 1e-4
    loss, x, _ = engine.fd_lbfgs_solve(basis.desc, V, R, opts, x0=x0, central="adjoint")
    best = loss.min(dim=1).values.cpu().numpy()
    assert (best <= 1e-9).mean() >= 0.9, (form, np.sort(best)[-4:])
    i = int(np.argmin(best))
    r = int(loss[i].argmin().item())
    assert abs(O.cost(orc.eval(x[i, r].cpu().numpy()), V[i].cpu().numpy(), "basic") - best[i]) < 1e-10

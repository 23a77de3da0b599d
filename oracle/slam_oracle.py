"""numpy/scipy restatement of the SLAM hot path (CPU oracle, test infrastructure).

See ``oracle/__init__.py`` for the rules on who may import this.  All paths
cited below are relative to ``/root/reference/``.  The arithmetic owners in the
reference are third-party packages that are not vendored and not installed in
this image (qiskit-terra ~0.2x, qutip>=4.7, weylchamber>=0.4, see
``setup.cfg:11-14``); their published algorithms are restated here and anchored
on the reference's own call sites and recorded outputs.

Conventions
-----------
* float64 / complex128 everywhere; matrices are row-major ``[..., 4, 4]``.
* qiskit little-endian: a layer ``u(q0); u(q1)`` has matrix ``U3(q1) (x) U3(q0)``.
* Parameter vectors crossing the API (``Xk``) are in *lexicographic name order*
  (``P0,P1,P10,...,P2,...`` then ``Q...``), as ``QuantumCircuit.parameters`` sorts.
"""
from __future__ import annotations

import math
import random
import sys
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple, Union

import numpy as np
import scipy.linalg
import scipy.optimize
import scipy.stats

__all__ = [
    "u3", "rz", "layer_matrix", "riswap", "conversion_gain", "conversion_gain_expm",
    "smush", "smush_1qphase", "canonical_gate", "berkeley", "fsim", "CNOT", "SWAP", "ISWAP",
    "OracleTemplate", "trace_overlap", "cost_from_trace", "cost", "COST_KINDS",
    "c1c2c3", "c1c2c3_raw", "fold_c1", "g1g2g3", "g1g2g3_raw", "J_T_LI", "haar_unitary", "haar_sample_unitary",
    "loss_and_grad", "fd_gradient", "literal_objective", "literal_run", "LiteralResult",
    "philox4x32_10", "philox_uniform", "coverage_params", "coverage_points", "coverage_histogram",
    "bin_index", "near_bin_edge", "trajectory", "trajectory_segments", "F_eval", "F_lossgrad",
]

# --------------------------------------------------------------------------------------
# 1Q gates and layers  (qiskit UGate / RZGate; layer order: src/slam/basis.py:152-169)
# --------------------------------------------------------------------------------------


def u3(theta, phi, lam):
    """qiskit ``UGate(theta, phi, lam)`` matrix, batched over leading axes."""
    theta, phi, lam = np.broadcast_arrays(np.asarray(theta, float), np.asarray(phi, float), np.asarray(lam, float))
    c = np.cos(theta / 2)
    s = np.sin(theta / 2)
    out = np.empty(theta.shape + (2, 2), dtype=np.complex128)
    out[..., 0, 0] = c
    out[..., 0, 1] = -np.exp(1j * lam) * s
    out[..., 1, 0] = np.exp(1j * phi) * s
    out[..., 1, 1] = np.exp(1j * (phi + lam)) * c
    return out


def rz(lam):
    """qiskit ``RZGate(lam)`` = diag(e^{-i lam/2}, e^{+i lam/2}) (used when ``vz_only``; basisv2.py:267-272)."""
    lam = np.asarray(lam, float)
    out = np.zeros(lam.shape + (2, 2), dtype=np.complex128)
    out[..., 0, 0] = np.exp(-0.5j * lam)
    out[..., 1, 1] = np.exp(0.5j * lam)
    return out


def _kron2(a, b):
    """Batched kron of [...,2,2] x [...,2,2] -> [...,4,4]."""
    return np.einsum("...ij,...kl->...ikjl", a, b).reshape(a.shape[:-2] + (4, 4))


def layer_matrix(g_q0, g_q1):
    """4x4 matrix of a 1Q layer: gate on qubit 1 is the most-significant factor."""
    return _kron2(g_q1, g_q0)


# --------------------------------------------------------------------------------------
# 2Q gates
# --------------------------------------------------------------------------------------

CNOT = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)  # qiskit CXGate (ctrl q0)
SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
ISWAP = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def riswap(alpha):
    """``RiSwapGate(alpha).__array__`` (src/slam/utils/gates/custom_gates.py:582-595)."""
    alpha = np.asarray(alpha, float)
    c = np.cos(np.pi * alpha / 2)
    s = np.sin(np.pi * alpha / 2)
    out = np.zeros(alpha.shape + (4, 4), dtype=np.complex128)
    out[..., 0, 0] = 1
    out[..., 3, 3] = 1
    out[..., 1, 1] = c
    out[..., 2, 2] = c
    out[..., 1, 2] = 1j * s
    out[..., 2, 1] = 1j * s
    return out


def _cg_hamiltonian(phi_c, phi_g, gc, gg):
    """``ConversionGainPhaseHamiltonian.foo_H`` (src/slam/hamiltonian.py:95-101) as a dense 4x4.

    qutip ``create(2) = [[0,0],[1,0]]``; A = a (x) I, B = I (x) a.  A B^dag = |10><01| so
    H_c has ``H[2,1] = e^{i phi_c}``; A B = |11><00| so H_g has ``H[3,0] = e^{i phi_g}``.
    """
    H = np.zeros((4, 4), dtype=np.complex128)
    H[2, 1] += gc * np.exp(1j * phi_c)
    H[1, 2] += gc * np.exp(-1j * phi_c)
    H[3, 0] += gg * np.exp(1j * phi_g)
    H[0, 3] += gg * np.exp(-1j * phi_g)
    return H


def conversion_gain_expm(phi_c, phi_g, gc, gg, t):
    """``ConversionGainGate(p1,p2,g1,g2,t).__array__`` literally: expm(-i t H) (custom_gates.py:180-184,
    hamiltonian.py:31-33,108-111; positional quirk p1=phi_c, p2=phi_g, g1=gc, g2=gg)."""
    return scipy.linalg.expm(-1j * float(t) * _cg_hamiltonian(phi_c, phi_g, gc, gg))


def conversion_gain(phi_c, phi_g, gc, gg, t):
    """Closed form of :func:`conversion_gain_expm`, batched (two 2x2 rotations)."""
    phi_c, phi_g, gc, gg, t = np.broadcast_arrays(*[np.asarray(v, float) for v in (phi_c, phi_g, gc, gg, t)])
    out = np.zeros(phi_c.shape + (4, 4), dtype=np.complex128)
    cg, sg = np.cos(gg * t), np.sin(gg * t)
    cc, sc = np.cos(gc * t), np.sin(gc * t)
    out[..., 0, 0] = cg
    out[..., 3, 3] = cg
    out[..., 0, 3] = -1j * np.exp(-1j * phi_g) * sg
    out[..., 3, 0] = -1j * np.exp(1j * phi_g) * sg
    out[..., 1, 1] = cc
    out[..., 2, 2] = cc
    out[..., 1, 2] = -1j * np.exp(-1j * phi_c) * sc
    out[..., 2, 1] = -1j * np.exp(1j * phi_c) * sc
    return out


_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_I2 = np.eye(2, dtype=np.complex128)
_a_dag = np.array([[0, 0], [1, 0]], dtype=np.complex128)  # qutip create(2)
_A = np.kron(_a_dag, _I2)
_B = np.kron(_I2, _a_dag)


def _smush_hamiltonian(phi_a, phi_b, phi_c, phi_g, gc, gg, gz1, gz2, gx, gy):
    """``ConversionGainSmush1QPhase.foo_H`` (hamiltonian.py:156-163); with phi_a=phi_b=gz=0 it is
    ``ConversionGainSmush.foo_H`` (hamiltonian.py:122-127)."""
    Hx = np.exp(1j * phi_a) * _A + np.exp(-1j * phi_a) * _A.conj().T
    Hy = np.exp(1j * phi_b) * _B + np.exp(-1j * phi_b) * _B.conj().T
    Hz1 = _A.conj().T @ _A
    Hz2 = _B.conj().T @ _B
    return gx * Hx + gy * Hy + _cg_hamiltonian(phi_c, phi_g, gc, gg) + gz1 * Hz1 + gz2 * Hz2


def smush_1qphase(phi_a, phi_b, phi_c, phi_g, gc, gg, gz1, gz2, gx, gy, t, t_slices=None):
    """``ConversionGainSmush1QPhase.construct_U`` (hamiltonian.py:168-182): ordered product of
    per-slice ``expm(-i (t/N) H_i)``, later slices on the left.  ``t_slices`` (optional) overrides
    the per-slice duration list (used by the trajectory path, pd_playground.py:185-203)."""
    gx = np.atleast_1d(np.asarray(gx, float))
    gy = np.atleast_1d(np.asarray(gy, float))
    assert gx.shape == gy.shape
    N = gx.shape[0]
    total = np.eye(4, dtype=np.complex128)
    for it in range(N):
        dt = float(t) / N if t_slices is None else float(t_slices[it])
        H = _smush_hamiltonian(phi_a, phi_b, phi_c, phi_g, gc, gg, gz1, gz2, gx[it], gy[it])
        total = scipy.linalg.expm(-1j * dt * H) @ total
    return total


def smush(phi_c, phi_g, gc, gg, gx, gy, t):
    """``ConversionGainSmush.construct_U`` (hamiltonian.py:132-144)."""
    return smush_1qphase(0.0, 0.0, phi_c, phi_g, gc, gg, 0.0, 0.0, gx, gy, t)


def canonical_gate(c1, c2, c3):
    """weylchamber ``canonical_gate``: exp(i pi/2 (c1 XX + c2 YY + c3 ZZ)) (used at custom_gates.py:389,
    cost_function.py:180-188)."""
    H = c1 * np.kron(_X, _X) + c2 * np.kron(_Y, _Y) + c3 * np.kron(_Z, _Z)
    return scipy.linalg.expm(0.5j * np.pi * H)


def berkeley():
    """``BerkeleyGate`` = ``CanonicalGate(pi/4, pi/8, 0)`` which rescales by 2/pi (custom_gates.py:384-397)."""
    return canonical_gate(0.5, 0.25, 0.0)


def fsim(theta, phi):
    """``FSim.__array__`` (custom_gates.py:494-503); only used as a KAT input (App. B5)."""
    return np.array(
        [[1, 0, 0, 0], [0, np.cos(theta), -1j * np.sin(theta), 0], [0, -1j * np.sin(theta), np.cos(theta), 0],
         [0, 0, 0, np.exp(1j * phi)]], dtype=np.complex128)


# --------------------------------------------------------------------------------------
# Templates (CircuitTemplate src/slam/basis.py:51-170, CircuitTemplateV2 src/slam/basisv2.py:27-299)
# --------------------------------------------------------------------------------------

GATE_SLOTS = {
    # kind -> number of leading scalar slots (before gx/gy vectors), has vectors, trailing t
    "riswap": ("alpha",),
    "cg": ("phi_c", "phi_g", "gc", "gg", "t"),
    "smush": ("phi_c", "phi_g", "gc", "gg", "gx*", "gy*", "t"),
    "smush1q": ("phi_a", "phi_b", "phi_c", "phi_g", "gc", "gg", "gz1", "gz2", "gx*", "gy*", "t"),
    "fixed": (),
}

Slot = Union[float, str]  # a float constant, or "Q" = a fresh 2Q parameter (created in slot order)


def _lex_sorted(names: Sequence[str]) -> List[str]:
    return sorted(names)  # qiskit ParameterView sorts by name string


@dataclass
class OracleTemplate:
    """Structure of a template circuit after ``build(k)``.

    ``gate_kind``: one of GATE_SLOTS.  ``slots``: per-repetition slot list (floats or "Q"); if a
    single list is given it is reused for every repetition (fresh Q's each time, as
    basisv2.py:282-288 creates them).  ``T`` = number of time slices for smush kinds.
    ``fixed``: 4x4 matrix for ``gate_kind == "fixed"``.
    """

    gate_kind: str = "riswap"
    slots: Sequence[Slot] = (0.5,)
    k: int = 1
    T: int = 0
    no_exterior_1q: bool = False
    vz_only: bool = False
    fixed: np.ndarray | None = None
    # derived
    instructions: list = field(default_factory=list, init=False)
    names_creation: List[str] = field(default_factory=list, init=False)
    names_sorted: List[str] = field(default_factory=list, init=False)

    def __post_init__(self):
        if self.k <= 0:
            raise ValueError()  # basis.py:127-128
        p = 0
        q = 0
        n1 = 1 if self.vz_only else 3
        ins = []
        names = []

        def one_q_layer():
            nonlocal p
            for qubit in (0, 1):
                nm = [f"P{p + j}" for j in range(n1)]
                p += n1
                names.extend(nm)
                ins.append(("rz" if self.vz_only else "u", qubit, nm))

        for i in range(self.k):
            initial = i == 0
            final = i == self.k - 1
            if initial and not self.no_exterior_1q:
                one_q_layer()
            bound = []
            for s in self.slots:
                if isinstance(s, str):
                    nm = f"Q{q}"
                    q += 1
                    names.append(nm)
                    bound.append(nm)
                else:
                    bound.append(float(s))
            ins.append(("gate", self.gate_kind, bound))
            if not (final and self.no_exterior_1q):
                one_q_layer()
        self.instructions = ins
        self.names_creation = names
        self.names_sorted = _lex_sorted(names)

    @property
    def n_params(self) -> int:
        return len(self.names_sorted)

    # ---- gate matrix from bound slot values -------------------------------------------------
    def _gate_matrix(self, vals: Sequence[float]) -> np.ndarray:
        kind = self.gate_kind
        T = self.T
        if kind == "riswap":
            return riswap(vals[0])
        if kind == "cg":
            return conversion_gain(*vals)
        if kind == "smush":
            return smush(vals[0], vals[1], vals[2], vals[3], vals[4:4 + T], vals[4 + T:4 + 2 * T], vals[-1])
        if kind == "smush1q":
            return smush_1qphase(*vals[:8], vals[8:8 + T], vals[8 + T:8 + 2 * T], vals[-1])
        if kind == "fixed":
            return np.asarray(self.fixed, dtype=np.complex128)
        raise ValueError(kind)

    def eval(self, Xk: Sequence[float]) -> np.ndarray:
        """``basis.eval(Xk)`` (basis.py:102-116): bind in lexicographic order, multiply the chain.
        One 4x4 product per instruction, as ``Operator(circuit)`` composes them."""
        if len(Xk) != self.n_params:
            raise ValueError("parameter count mismatch")
        val = dict(zip(self.names_sorted, (float(v) for v in Xk)))
        U = np.eye(4, dtype=np.complex128)
        for op in self.instructions:
            if op[0] == "u":
                g = u3(*(val[n] for n in op[2]))
                M = np.kron(_I2, g) if op[1] == 0 else np.kron(g, _I2)
            elif op[0] == "rz":
                g = rz(val[op[2][0]])
                M = np.kron(_I2, g) if op[1] == 0 else np.kron(g, _I2)
            else:
                M = self._gate_matrix([val[s] if isinstance(s, str) else s for s in op[2]])
            U = M @ U
        return U

    def eval_batch(self, X: np.ndarray) -> np.ndarray:
        """Vectorised over the leading axis (closed-form gates only; smush falls back to a loop)."""
        X = np.asarray(X, float)
        if self.gate_kind in ("smush", "smush1q"):
            return np.stack([self.eval(x) for x in X])
        idx = {n: i for i, n in enumerate(self.names_sorted)}
        B = X.shape[0]
        U = np.broadcast_to(np.eye(4, dtype=np.complex128), (B, 4, 4)).copy()
        eye2 = np.broadcast_to(_I2, (B, 2, 2))
        for op in self.instructions:
            if op[0] in ("u", "rz"):
                if op[0] == "u":
                    g = u3(*(X[:, idx[n]] for n in op[2]))
                else:
                    g = rz(X[:, idx[op[2][0]]])
                M = _kron2(eye2, g) if op[1] == 0 else _kron2(g, eye2)
            else:
                cols = [X[:, idx[s]] if isinstance(s, str) else np.full(B, s) for s in op[2]]
                if self.gate_kind == "riswap":
                    M = riswap(cols[0])
                elif self.gate_kind == "cg":
                    M = conversion_gain(*cols)
                else:
                    M = np.broadcast_to(np.asarray(self.fixed, np.complex128), (B, 4, 4))
            U = M @ U
        return U

    # ---- random initial points ---------------------------------------------------------------
    def parameter_guess(self, rng: np.random.Generator, v2: bool = False) -> np.ndarray:
        """x0 distribution: ``CircuitTemplate`` U[0,2pi)^P (basis.py:111); V2 U(-4pi,4pi) (basisv2.py:157-167)."""
        if v2:
            return rng.uniform(-4 * np.pi, 4 * np.pi, self.n_params)
        return rng.random(self.n_params) * 2 * np.pi


# --------------------------------------------------------------------------------------
# Cost functionals (src/slam/cost_function.py:133-221)
# --------------------------------------------------------------------------------------

COST_KINDS = ("basic", "square", "basic_inverse", "weyl_euclidean", "makhlin_euclidean", "makhlin_functional",
              "basic_reduced", "square_reduced")


def trace_overlap(current_u, target_u):
    """T = Tr(V^dag U), batched."""
    return np.einsum("...ij,...ij->...", np.conj(target_u), current_u)


def cost_from_trace(T, kind: str):
    a = np.abs(T)
    if kind == "basic":  # cost_function.py:140-145  (NOT squared)
        return 1 - a / 4
    if kind == "square":  # cost_function.py:169-173
        return 1 - (a ** 2 + 4) / 20
    if kind == "basic_inverse":  # cost_function.py:133-137
        return a / 4
    raise ValueError(kind)


def cost(current_u, target_u, kind: str = "basic") -> float:
    """``objective.unitary_fidelity(current_u, target_u)`` for every in-scope functional."""
    if kind in ("basic", "square", "basic_inverse"):
        return float(cost_from_trace(trace_overlap(current_u, target_u), kind))
    if kind == "weyl_euclidean":  # cost_function.py:199-206
        return float(np.linalg.norm(np.array(c1c2c3(target_u)) - np.array(c1c2c3(current_u))))
    if kind == "makhlin_euclidean":  # cost_function.py:209-216
        return float(np.linalg.norm(np.array(g1g2g3(target_u)) - np.array(g1g2g3(current_u))))
    if kind == "makhlin_functional":  # cost_function.py:219-221
        return float(J_T_LI(target_u, current_u))
    if kind in ("basic_reduced", "square_reduced"):  # cost_function.py:176-189
        ct = canonical_gate(*c1c2c3(target_u))
        cc = canonical_gate(*c1c2c3(current_u))
        return cost(cc, ct, kind.split("_")[0])
    raise ValueError("Unrecognized Cost Function")  # optimizer.py:211


# --------------------------------------------------------------------------------------
# Weyl-chamber coordinates and Makhlin invariants (weylchamber>=0.4; Childs et al. PRA 68 052311)
# --------------------------------------------------------------------------------------

_SYSY = np.array([[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]], dtype=np.complex128)
_QMAGIC = (1 / np.sqrt(2)) * np.array([[1, 0, 0, 1j], [0, 1j, 1, 0], [0, 1j, -1, 0], [1, 0, 0, -1j]], dtype=np.complex128)
_WEYL_M = np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1]], dtype=float)


def c1c2c3_raw(U):
    """Un-rounded Weyl coordinates (units of pi), batched over leading axes.  SURVEY App. A.6."""
    U = np.asarray(U, dtype=np.complex128)
    Ut = _SYSY @ np.swapaxes(U, -1, -2) @ _SYSY
    det = np.linalg.det(U)
    ev = np.linalg.eigvals((U @ Ut) / np.sqrt(det.astype(np.complex128))[..., None, None])
    two_S = np.angle(ev) / np.pi
    two_S = np.where(two_S <= -0.5, two_S + 2.0, two_S)
    S = -np.sort(-two_S / 2.0, axis=-1)  # descending
    n = np.rint(S.sum(axis=-1)).astype(int)
    ar = np.arange(4)
    S = S - (ar < n[..., None])
    # roll(S, -n): out[i] = S[(i+n) % 4]
    S = np.take_along_axis(S, (ar + n[..., None]) % 4, axis=-1)
    c = S[..., :3] @ _WEYL_M.T
    flip = c[..., 2] < 0
    c1 = np.where(flip, 1 - c[..., 0], c[..., 0])
    c3 = np.where(flip, -c[..., 2], c[..., 2])
    return np.stack([c1 + 0.0, c[..., 1] + 0.0, c3 + 0.0], axis=-1)


def c1c2c3(U, ndigits: int = 8) -> Tuple[float, float, float]:
    """``weylchamber.c1c2c3(U)``: 8-decimal rounded tuple (single matrix)."""
    c = c1c2c3_raw(np.asarray(U))
    return tuple(round(float(v) + 0.0, ndigits) for v in c)


def fold_c1(c):
    """Mirror-fold c1 > 1/2 -> 1 - c1 (pd_playground.py:199-202, parallel_drive_volume.py:295-304)."""
    c = np.array(c, dtype=float, copy=True)
    c[..., 0] = np.where(c[..., 0] > 0.5, 1.0 - c[..., 0], c[..., 0])
    return c


def g1g2g3_raw(U):
    """Un-rounded Makhlin invariants (g1, g2, g3), batched.  SURVEY App. A.7."""
    U = np.asarray(U, dtype=np.complex128)
    UB = _QMAGIC.conj().T @ U @ _QMAGIC
    detU = np.linalg.det(UB)
    m = np.swapaxes(UB, -1, -2) @ UB
    tr = np.trace(m, axis1=-2, axis2=-1)
    tr2 = np.trace(m @ m, axis1=-2, axis2=-1)
    g12 = tr ** 2 / (16.0 * detU)
    g3 = (tr ** 2 - tr2) / (4.0 * detU)
    return np.stack([g12.real + 0.0, g12.imag + 0.0, g3.real + 0.0], axis=-1)


def g1g2g3(U, ndigits: int = 8) -> Tuple[float, float, float]:
    g = g1g2g3_raw(np.asarray(U))
    return tuple(round(float(v) + 0.0, ndigits) for v in g)


def J_T_LI(O, U) -> float:
    """weylchamber ``J_T_LI(O, U, form='g')`` on the 8-dp rounded invariants (bit-pinned by KAT B5b)."""
    return float(np.sum(np.abs(np.array(g1g2g3(O)) - np.array(g1g2g3(U))) ** 2))


# --------------------------------------------------------------------------------------
# Samplers (src/slam/sampler.py:62-71 + qiskit random_unitary -> scipy unitary_group)
# --------------------------------------------------------------------------------------


def haar_unitary(rng: np.random.Generator, n: int | None = None) -> np.ndarray:
    """Haar U(4) exactly as scipy.stats.unitary_group.rvs(4) draws it."""
    if n is None:
        return scipy.stats.unitary_group.rvs(4, random_state=rng)
    return np.stack([scipy.stats.unitary_group.rvs(4, random_state=rng) for _ in range(n)])


def haar_sample_unitary(seed=None) -> np.ndarray:
    """``HaarSample(seed)._get_unitary()``: re-seeds python ``random`` on every call (sampler.py:67-71)."""
    random.seed(seed)
    s = random.randint(0, sys.maxsize)
    return scipy.stats.unitary_group.rvs(4, random_state=np.random.default_rng(s))


# --------------------------------------------------------------------------------------
# Analytic loss + gradient (second, independent oracle for the CUDA adjoint kernel)
# --------------------------------------------------------------------------------------


def _du3(theta, phi, lam):
    """Partial derivatives of u3 wrt (theta, phi, lam): three [2,2] matrices."""
    c = np.cos(theta / 2)
    s = np.sin(theta / 2)
    el, ep = np.exp(1j * lam), np.exp(1j * phi)
    dth = 0.5 * np.array([[-s, -el * c], [ep * c, -ep * el * s]])
    dph = np.array([[0, 0], [1j * ep * s, 1j * ep * el * c]])
    dla = np.array([[0, -1j * el * s], [0, 1j * ep * el * c]])
    return dth, dph, dla


def loss_and_grad(tmpl: OracleTemplate, Xk, target_u, kind: str = "basic", h_gate: float = 1e-6,
                  richardson: bool = False):
    """Loss and analytic gradient (API parameter order) by explicit derivative matrices.

    1Q parameters use closed-form derivative matrices; 2Q (``Q``) parameters
    use a central difference of the *gate matrix only* (step ``h_gate``), which is exact to O(h^2)
    and independent of the CUDA kernels' derivatives (closed form for RiSwap / ConversionGain, Daleckii-Krein
    through the slice exponentials for the smush gates).  ``richardson=True`` combines steps h and h/2 to O(h^4)
    (use h_gate ~ 2e-3: truncation ~1e-12, round-off ~1e-13).
    """
    Xk = np.asarray(Xk, float)
    val = dict(zip(tmpl.names_sorted, Xk))
    mats = []
    for op in tmpl.instructions:
        if op[0] == "u":
            g = u3(*(val[n] for n in op[2]))
            mats.append(np.kron(_I2, g) if op[1] == 0 else np.kron(g, _I2))
        elif op[0] == "rz":
            g = rz(val[op[2][0]])
            mats.append(np.kron(_I2, g) if op[1] == 0 else np.kron(g, _I2))
        else:
            mats.append(tmpl._gate_matrix([val[s] if isinstance(s, str) else s for s in op[2]]))
    n = len(mats)
    pre = [np.eye(4, dtype=np.complex128)]
    for M in mats:
        pre.append(M @ pre[-1])
    suf = [np.eye(4, dtype=np.complex128)]
    for M in reversed(mats):
        suf.append(suf[-1] @ M)
    suf = suf[::-1]  # suf[i] = mats[n-1] ... mats[i]
    U = pre[n]
    Vh = np.conj(np.asarray(target_u)).T
    T = np.trace(Vh @ U)
    dT = {}
    for i, op in enumerate(tmpl.instructions):
        env = pre[i] @ Vh @ suf[i + 1]  # Tr(V^dag suf dM pre) = Tr(dM env)
        if op[0] == "u":
            ds = _du3(*(val[nm] for nm in op[2]))
            for nm, d in zip(op[2], ds):
                D = np.kron(_I2, d) if op[1] == 0 else np.kron(d, _I2)
                dT[nm] = dT.get(nm, 0) + np.trace(D @ env)
        elif op[0] == "rz":
            lam = val[op[2][0]]
            d = np.diag([-0.5j * np.exp(-0.5j * lam), 0.5j * np.exp(0.5j * lam)])
            D = np.kron(_I2, d) if op[1] == 0 else np.kron(d, _I2)
            dT[op[2][0]] = dT.get(op[2][0], 0) + np.trace(D @ env)
        else:
            for j, s in enumerate(op[2]):
                if isinstance(s, str):
                    v0 = [val[q] if isinstance(q, str) else q for q in op[2]]

                    def cdiff(h):
                        vp, vm = list(v0), list(v0)
                        vp[j] += h
                        vm[j] -= h
                        return (tmpl._gate_matrix(vp) - tmpl._gate_matrix(vm)) / (2 * h)

                    D = (4 * cdiff(h_gate / 2) - cdiff(h_gate)) / 3 if richardson else cdiff(h_gate)
                    dT[s] = dT.get(s, 0) + np.trace(D @ env)
    a = abs(T)
    dabs = {nm: (np.conj(T) * d).real / a for nm, d in dT.items()}
    if kind == "basic":
        loss = 1 - a / 4
        g = {nm: -v / 4 for nm, v in dabs.items()}
    elif kind == "square":
        loss = 1 - (a * a + 4) / 20
        g = {nm: -2 * a * v / 20 for nm, v in dabs.items()}
    elif kind == "basic_inverse":
        loss = a / 4
        g = {nm: v / 4 for nm, v in dabs.items()}
    else:
        raise ValueError(kind)
    grad = np.array([g[nm] for nm in tmpl.names_sorted])
    return float(loss), grad, complex(T)


def fd_gradient(f, x, eps: float = 1.4901161193847656e-08):
    """scipy's default 2-point forward difference (what ``opt.minimize(jac=None)`` uses; optimizer.py:270-278)."""
    x = np.asarray(x, float)
    f0 = f(x)
    g = np.empty_like(x)
    for i in range(x.size):
        xp = x.copy()
        xp[i] += eps
        g[i] = (f(xp) - f0) / eps
    return f0, g


# --------------------------------------------------------------------------------------
# Literal restatement of the optimiser loop (src/slam/optimizer.py:188-313) -- the timed CPU baseline
# --------------------------------------------------------------------------------------


def literal_objective(tmpl: OracleTemplate, target_u, kind: str = "basic"):
    """``objective_func`` closure (optimizer.py:191-214): one template evaluation + one cost per call."""
    Vh = np.conj(np.asarray(target_u)).T

    def f(xk):
        U = tmpl.eval(xk)
        if kind in ("basic", "square", "basic_inverse"):
            a = abs(np.trace(Vh @ U))
            if kind == "basic":
                return 1 - a / 4
            if kind == "square":
                return 1 - (a * a + 4) / 20
            return a / 4
        return cost(U, target_u, kind)

    return f


@dataclass
class LiteralResult:
    best_result: float
    best_Xk: np.ndarray
    best_cycles: int
    nfev: int
    success_label: int


def literal_run(make_template, target_u, k_range, restarts: int = 5, kind: str = "basic",
                success_threshold: float = 1e-10, rng: np.random.Generator | None = None, v2: bool = False,
                method: str = "BFGS", maxiter: int = 2500) -> LiteralResult:
    """``TemplateOptimizer._run`` (optimizer.py:188-313): k-loop x restart-loop x scipy BFGS with
    finite-difference gradients and the reference's early exits.  ``make_template(k)`` -> OracleTemplate."""
    rng = rng or np.random.default_rng()
    best_result, best_Xk, best_cycles = None, None, -1
    nfev = 0
    for k in k_range:
        tmpl = make_template(k)
        f = literal_objective(tmpl, target_u, kind)
        for _ in range(restarts):
            x0 = tmpl.parameter_guess(rng, v2=v2)
            res = scipy.optimize.minimize(fun=f, method=method, x0=x0, options={"maxiter": maxiter})
            nfev += res.nfev
            if best_result is None or res.fun < best_result:
                best_result, best_Xk, best_cycles = float(res.fun), res.x, k
            if best_result < success_threshold:
                break
        if best_result < success_threshold:
            break
    return LiteralResult(best_result, best_Xk, best_cycles, nfev, int(best_result <= success_threshold))


# --------------------------------------------------------------------------------------
# Counter-based RNG shared with the CUDA coverage kernel: Philox4x32-10 (Salmon et al., SC'11)
# --------------------------------------------------------------------------------------

_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)
PHILOX_STREAM_TAG = 0x51A3B200


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(*[np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3, k0, k1)])
    c0, c1, c2, c3, k0, k1 = (a.copy() for a in (c0, c1, c2, c3, k0, k1))
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r != 9:
                k0 = k0 + _PH_W0
                k1 = k1 + _PH_W1
    return c0, c1, c2, c3


def philox_uniform(seed: int, sample, n_params: int) -> np.ndarray:
    """u[sample, j] in [0,1), 53-bit: counter = (sample_lo, sample_hi, j//2, TAG), key = (seed_lo, seed_hi);
    words (w0,w1) -> param 2*(j//2), (w2,w3) -> param 2*(j//2)+1; u = ((w_hi<<32 | w_lo) >> 11) * 2^-53."""
    sample = np.asarray(sample, dtype=np.uint64)
    s_lo = (sample & _MASK32).astype(np.uint32)[:, None]
    s_hi = (sample >> np.uint64(32)).astype(np.uint32)[:, None]
    nblk = (n_params + 1) // 2
    blk = np.arange(nblk, dtype=np.uint32)[None, :]
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w0, w1, w2, w3 = philox4x32_10(s_lo, s_hi, blk, np.uint32(PHILOX_STREAM_TAG), k0, k1)

    def to_u(hi, lo):
        v = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
        return (v >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    u = np.empty((sample.shape[0], 2 * nblk))
    u[:, 0::2] = to_u(w0, w1)
    u[:, 1::2] = to_u(w2, w3)
    return u[:, :n_params]


# --------------------------------------------------------------------------------------
# Coverage-set Monte-Carlo (src/slam/utils/gates/parallel_drive_volume.py:175-225, 292-307)
# --------------------------------------------------------------------------------------


def coverage_params(seed: int, first_sample: int, n_samples: int, n_params: int, lo: float, hi: float) -> np.ndarray:
    """params[n, j] = lo + (hi - lo) * u[n, j] (separate multiply and add, no FMA), j in API order."""
    u = philox_uniform(seed, np.arange(first_sample, first_sample + n_samples, dtype=np.uint64), n_params)
    return lo + (hi - lo) * u


def coverage_points(tmpl: OracleTemplate, params: np.ndarray) -> np.ndarray:
    """N x [template eval -> c1c2c3 -> mirror fold]; un-rounded folded coordinates."""
    U = tmpl.eval_batch(params)
    return fold_c1(c1c2c3_raw(U))


def bin_index(c: np.ndarray, nbins: int = 128) -> np.ndarray:
    """Flat bin index on the folded chamber box [0,1/2]^3 with `nbins` cells per axis."""
    ijk = np.clip(np.floor(c * (2.0 * nbins)).astype(np.int64), 0, nbins - 1)
    return (ijk[..., 0] * nbins + ijk[..., 1]) * nbins + ijk[..., 2]


def near_bin_edge(c: np.ndarray, nbins: int = 128, tol: float = 1e-9) -> np.ndarray:
    s = c * (2.0 * nbins)
    return np.any(np.abs(s - np.rint(s)) < tol * 2.0 * nbins, axis=-1)


def coverage_histogram(tmpl: OracleTemplate, seed: int, first_sample: int, n_samples: int, lo: float, hi: float,
                       nbins: int = 128) -> np.ndarray:
    p = coverage_params(seed, first_sample, n_samples, tmpl.n_params, lo, hi)
    idx = bin_index(coverage_points(tmpl, p), nbins)
    return np.bincount(idx, minlength=nbins ** 3).astype(np.int64)


# --------------------------------------------------------------------------------------
# Parallel-drive Weyl trajectory (src/slam/utils/pd_playground.py:169-208)
# --------------------------------------------------------------------------------------


def trajectory(phases, gc, gg, gz1, gz2, gx_vec, gy_vec, dt: float, R: int = 5):
    """``ParallelDrivenGateWidget.iterate_time``: for end = 1..N, earlier slices at full ``dt``, the
    last one at ``linspace(0, dt, R)``; returns (coords [N,R,3] folded & 8-dp rounded as the
    reference stores them, final unitary)."""
    pa, pb, pc, pg = phases
    N = len(gx_vec)
    coords = np.zeros((N, R, 3))
    prefix = np.eye(4, dtype=np.complex128)
    U = prefix
    for end in range(1, N + 1):
        for r, t in enumerate(np.linspace(0, dt, R)):
            Ui = smush_1qphase(pa, pb, pc, pg, gc, gg, gz1, gz2, [gx_vec[end - 1]], [gy_vec[end - 1]], t)
            U = Ui @ prefix
            c = list(c1c2c3(U))
            if c[0] > 0.5:
                c[0] = -1 * c[0] + 1
            coords[end - 1, r] = c
        prefix = U  # last t == dt
    return coords, U


def trajectory_segments(gate_rows, gx_vec, gy_vec, dt: float, R: int = 5):
    """Composed widgets ``pdgw + pdgw2 (+ ...)`` (``ParallelDrivenGateWidget.__add__``, pd_playground.py:46-58: the circuits
    of the operands are concatenated, so every slice keeps the phases / couplings of the widget it came from) followed by
    ``iterate_time`` (pd_playground.py:179-208).  ``gate_rows`` [N, 8] = (phase_a, phase_b, phase_c, phase_g, gc, gg, gz1,
    gz2) per slice.  Returns (coords [N,R,3] folded & rounded, final unitary)."""
    gate_rows = np.asarray(gate_rows, float)
    N = len(gx_vec)
    coords = np.zeros((N, R, 3))
    prefix = np.eye(4, dtype=np.complex128)
    U = prefix
    for end in range(1, N + 1):
        g = gate_rows[end - 1]
        for r, t in enumerate(np.linspace(0, dt, R)):
            Ui = smush_1qphase(g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], [gx_vec[end - 1]], [gy_vec[end - 1]], t)
            U = Ui @ prefix
            c = list(c1c2c3(U))
            if c[0] > 0.5:
                c[0] = -1 * c[0] + 1
            coords[end - 1, r] = c
        prefix = U
    return coords, U


# --------------------------------------------------------------------------------------
# Algorithmic FLOP counts -- shared by bench.py's roofline
# --------------------------------------------------------------------------------------


def F_eval(k: int) -> int:
    return 512 * 2 * k + 124 * (k + 1) + 128


def F_lossgrad(k: int) -> int:
    return F_eval(k) + 512 * (3 * (k + 1) + 2 * k) + 768 * (k + 1)

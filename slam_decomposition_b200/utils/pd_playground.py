"""Parallel-driven gate trajectories (reference: src/slam/utils/pd_playground.py:17-339).

``ParallelDrivenGateWidget`` strings N copies of a one-slice ``ConversionGainSmush1QPhaseGate`` (t = 0.1 each) and
follows the Weyl-chamber trajectory while the last slice's duration sweeps ``linspace(0, dt, R)``.  The reference
rebuilds and re-exponentiates the whole prefix for every point (O(N^2 R) qutip expm calls); the kernel
(``slam_pd_trajectory`` / ``slam_pd_trajectory_slices``) keeps the prefix product and does one slice exponential + one
Weyl evaluation per point.  Widgets add (``pdgw + pdgw2``, pd_playground.py:46-58) into multi-segment pulses whose
drive phases and couplings change from segment to segment; ``pulse_template`` / ``search_pulse`` lower such a pulse, with
free and TIED parameters, to a template descriptor and run the outer search of
``scripts/parallel_drive_swap/parallel_drive_swap.ipynb`` (cells 7, 9, 10: scipy Nelder-Mead over a handful of pulse
parameters, one start at a time) as a batch of thousands of starts in the device Nelder-Mead kernel.
The plotting / ipywidget glue of the reference is out of scope; ``coordinate_list`` keeps its format
(list over slices of lists of [c1, c2, c3]) so ``visualize.coordinate_2dlist_weyl`` could consume it unchanged.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _lib, engine
from ..circuit import Parameter, TemplateCircuit, lower
from .gates.custom_gates import ConversionGainSmush1QPhaseGate

_FIELDS = ("phase_a", "phase_b", "phase_c", "phase_g", "gc", "gg", "gz1", "gz2")


def trajectories(gate, gx, gy, dt: float = 0.1, R: int = 5, fold: bool = True, round8: bool = True):
    """Batched form: gate [B,8] = (phase_a, phase_b, phase_c, phase_g, gc, gg, gz1, gz2) per trajectory, or [B,N,8] per
    slice; gx/gy [B,N] -> (coords [B,N,R,3], final unitaries [B,4,4]) as CUDA tensors."""
    dev = engine.require_cuda()
    as_t = lambda a: a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, float), device=dev)
    gate, gx, gy = as_t(gate), as_t(gx), as_t(gy)
    if gate.dim() == 3:
        return engine.pd_trajectory(gate, gx, gy, dt, R=R, fold=fold, round8=round8)
    return engine.pd_trajectory(gate.reshape(-1, 8), gx.reshape(gate.reshape(-1, 8).shape[0], -1),
                                gy.reshape(gate.reshape(-1, 8).shape[0], -1), dt, R=R, fold=fold, round8=round8)


class ParallelDrivenGateWidget:
    def __init__(self, N=10, gc=np.pi / 2, gg=0, gz1=0, gz2=0, phase_a=0, phase_b=0, phase_c=0, phase_g=0) -> None:
        self.N = N
        self.gc, self.gg, self.gz1, self.gz2 = gc, gg, gz1, gz2
        self.phase_a, self.phase_b, self.phase_c, self.phase_g = phase_a, phase_b, phase_c, phase_g
        self.t = 0.1
        self.timesteps = 1
        self.duration_1q = self.t / self.timesteps
        self._segments: Optional[List["ParallelDrivenGateWidget"]] = None  # set on sums of widgets
        self.construct_basis()
        self.prepare_parameters(0, 0)
        self.fig = None

    def __add__(self, other: "ParallelDrivenGateWidget") -> "ParallelDrivenGateWidget":
        """pd_playground.py:46-58: the circuits are concatenated (parameters of the left operand renamed), N adds; every
        slice keeps the phases and couplings of the widget it came from."""
        ret = ParallelDrivenGateWidget()
        ret._segments = (self._segments or [self]) + (other._segments or [other])
        ret.N = self.N + other.N
        ret.construct_basis()
        ret.prepare_parameters(0, 0)
        return ret

    # the reference composes N qiskit copies of the one-slice template; here the "circuit" is the per-slice gate table
    def _row(self):
        return [float(getattr(self, f)) for f in _FIELDS]

    def construct_basis(self):
        if self._segments is None:
            self._gate = np.array(self._row(), dtype=np.float64)
            self._gate_rows = np.tile(self._gate, (self.N, 1))
        else:
            self._gate_rows = np.concatenate([np.tile(np.array(w._row(), dtype=np.float64), (w.N, 1)) for w in self._segments])
            self._gate = self._gate_rows[0]

    def prepare_parameters(self, q0, q1):
        self._gx = np.full(self.N, float(q0))
        self._gy = np.full(self.N, float(q1))

    def prepare_parameters_nonuniform(self, g0_vector, g1_vector):
        assert len(g0_vector) == len(g1_vector) == self.N
        self._gx = np.asarray(g0_vector, dtype=np.float64)
        self._gy = np.asarray(g1_vector, dtype=np.float64)

    def _run(self, R, want_coords):
        self.construct_basis()
        dev = engine.require_cuda()
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a)[None], device=dev)
        gate = t(self._gate_rows) if self._segments is not None else t(self._gate)
        return engine.pd_trajectory(gate, t(self._gx), t(self._gy), self.duration_1q, R=R, fold=True, round8=True,
                                    want_coords=want_coords)

    def solve_end(self):
        """All slices at full duration -> final 4x4 unitary (pd_playground.py:169-177)."""
        _, U = self._run(2, False)  # R = 2: linspace(0, dt, 2)[-1] == dt; only the end point of each slice is evaluated
        return U[0].cpu().numpy()

    def iterate_time(self, R=5):
        """Weyl trajectory (pd_playground.py:179-208): sets coordinate_list, end_segment_list, final_unitary."""
        coords, U = self._run(R, True)
        c = coords[0].cpu().numpy()
        self.coordinate_list = [[list(map(float, c[s, r])) for r in range(R)] for s in range(self.N)]
        self.end_segment_list = [self.coordinate_list[s][-1] for s in range(self.N)]
        self.final_unitary = U[0].cpu().numpy()

    def widget_wrap(self, q0, q1, pa, pb, pc, pg, gz1, gz2):
        self.gz1, self.gz2 = gz1, gz2
        self.phase_a, self.phase_b, self.phase_c, self.phase_g = pa, pb, pc, pg
        self.construct_basis()
        self.prepare_parameters(q0, q1)
        self.iterate_time()

    def widget_wrap_nonuniform(self, g0_vector, g1_vector):
        self.prepare_parameters_nonuniform(g0_vector, g1_vector)
        self.iterate_time()


class ImprovedCX(ParallelDrivenGateWidget):
    """pd_playground.py:247-283: the parallel-driven CX (gx = 3 on every slice of the default iSWAP drive), with the two 1Q
    markers appended to the trajectory and the undriven baseline (I -> sqrt(iSWAP) -> CX along the chamber floor)."""

    def __init__(self):
        super().__init__()
        self.prepare_parameters_nonuniform([3] * self.N, [0] * self.N)
        self.iterate_time()
        R = 5
        self.coordinate_list.append([[0, 0, 0]] * R)
        self.coordinate_list.append([(0.5, 0, 0)] * R)
        i_steps = 25  # R * t / duration_1q
        baseline = [[[0.25 / (i_steps - 1) * i, 0.25 / (i_steps - 1) * i, 0] for i in range(i_steps)],
                    [[0.25 + 0.25 / (i_steps - 1) * i, 0.25 - 0.25 / (i_steps - 1) * i, 0] for i in range(i_steps)],
                    [[0, 0, 0]] * R, [[0.25, 0.25, 0]] * R, [[0.5, 0, 0]] * R]
        self.baseline_coords = baseline


class ImprovedSWAP(ParallelDrivenGateWidget):
    """pd_playground.py:286-339: gx = gy = pi on every slice, then the tail found in decomp_trajectory.ipynb (two U3 gates on
    qubit 0 -- the reference appends both to qubit 0), whose coordinate is appended once per np.linspace(0, R) sample (50
    points, each a flat [c1, c2, c3] entry, as the reference's loop does), then the two 1Q markers."""

    TAIL_U3 = ((7.84862563826406, 9.44285614361501, 2.30856826810552), (7.85928560541358, 9.44027709402712, -3.9923157086907))

    def __init__(self):
        super().__init__()
        self.prepare_parameters_nonuniform([np.pi] * self.N, [np.pi] * self.N)
        self.iterate_time()
        R = 5
        U = self.final_unitary

        def u3(theta, phi, lam):
            c, s = np.cos(theta / 2), np.sin(theta / 2)
            return np.array([[c, -np.exp(1j * lam) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]])

        for tri in self.TAIL_U3:
            U = np.kron(np.eye(2), u3(*tri)) @ U  # qubit 0 = least significant factor
        from ..weyl import c1c2c3

        c = list(c1c2c3(U))
        if c[0] > 0.5:
            c[0] = -1 * c[0] + 1
        self.extended_unitary = U
        for _ in np.linspace(0, R):
            self.coordinate_list.append(c)
        self.coordinate_list.append([[0, 0, 0]] * 5)
        self.coordinate_list.append([(0.5, 0.5, 0.5)] * 5)


# ----------------------------------------------------------------------------------------------------------------------
# multi-segment pulses with free / tied parameters and their batched outer search
# ----------------------------------------------------------------------------------------------------------------------
def pulse_template(segments: Sequence[dict], dt: float = 0.1):
    """Lower a multi-segment pulse to a template descriptor: one ``ConversionGainSmush1QPhaseGate`` slice (T = 1, t = dt) per
    time step, no 1Q gates.  ``segments``: dicts with ``N`` (slices) and the fields phase_a, phase_b, phase_c, phase_g, gc,
    gg, gz1, gz2, gx, gy (default 0); a field is a number, a ``Parameter`` (shared by every slice of the segment -- the
    amplitudes tied across slices of notebook cell 7 -- and by every segment that uses the same Parameter), or for gx / gy a
    sequence of N numbers / Parameters (free per-slice amplitudes, cell 9).  Returns (desc, parameter names in Xk order)."""
    qc = TemplateCircuit(2)
    for seg in segments:
        n = int(seg["N"])
        for s in range(n):
            vals = [seg.get(f, 0.0) for f in _FIELDS]
            gx, gy = seg.get("gx", 0.0), seg.get("gy", 0.0)
            gx = gx[s] if isinstance(gx, (list, tuple, np.ndarray)) else gx
            gy = gy[s] if isinstance(gy, (list, tuple, np.ndarray)) else gy
            qc.append(ConversionGainSmush1QPhaseGate(*vals, [gx], [gy], dt), (0, 1))
    desc, names, numeric = lower(qc, no_exterior_1q=True)
    assert numeric.size == 0
    return desc, names


def search_pulse(segments: Sequence[dict], target, objective, n_starts: int = 4096, lo: float = -2 * np.pi,
                 hi: float = 2 * np.pi, x0=None, seed: int = 0, dt: float = 0.1, max_iter: int = 2500):
    """The outer search of parallel_drive_swap.ipynb cells 7 / 9 / 10 -- ``scipy.optimize.minimize(cost_function, p0,
    method="Nelder-Mead")`` with ``cost_function(p) = objective.unitary_fidelity(target, pulse(p).solve_end())`` -- for
    ``n_starts`` starting points at once in the device Nelder-Mead kernel (starts ~ U(lo, hi) as in cell 10, or the rows of
    ``x0``).  Returns dict(loss [n_starts], x [n_starts, P], names, best_loss, best_x)."""
    desc, names = pulse_template(segments, dt=dt)
    dev = engine.require_cuda()
    V = torch.as_tensor(np.asarray(target, dtype=np.complex128)[None], device=dev)
    nm = engine.nm_defaults()
    nm.cost_kind, nm.max_iter, nm.early_exit = int(objective.cost_kind), int(max_iter), 0
    nm.x0_lo, nm.x0_hi = float(lo), float(hi)
    x0_t = None
    if x0 is not None:
        x0_t = torch.as_tensor(np.asarray(x0, dtype=np.float64).reshape(1, -1, desc.n_params), device=dev)
        n_starts = x0_t.shape[1]
    loss, x, iters = engine.nm_solve(desc, V, int(n_starts), nm, x0=x0_t, seed=int(seed))
    loss, x = loss[0].cpu().numpy(), x[0].cpu().numpy()
    i = int(np.nanargmin(loss))
    return {"loss": loss, "x": x, "names": names, "best_loss": float(loss[i]), "best_x": x[i], "iters": iters[0].cpu().numpy()}

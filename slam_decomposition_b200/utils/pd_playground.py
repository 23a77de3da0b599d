"""Parallel-driven gate trajectories (reference: src/slam/utils/pd_playground.py:17-208).

``ParallelDrivenGateWidget`` strings N copies of a one-slice ``ConversionGainSmush1QPhaseGate`` (t = 0.1 each) and
follows the Weyl-chamber trajectory while the last slice's duration sweeps ``linspace(0, dt, R)``.  The reference
rebuilds and re-exponentiates the whole prefix for every point (O(N^2 R) qutip expm calls); the kernel
(``slam_pd_trajectory``) keeps the prefix product and does one slice exponential + one Weyl evaluation per point.
The plotting / ipywidget glue of the reference is out of scope; ``coordinate_list`` keeps its format
(list over slices of lists of [c1, c2, c3]) so ``visualize.coordinate_2dlist_weyl`` could consume it unchanged.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine


def trajectories(gate, gx, gy, dt: float = 0.1, R: int = 5, fold: bool = True, round8: bool = True):
    """Batched form: gate [B,8] = (phase_a, phase_b, phase_c, phase_g, gc, gg, gz1, gz2), gx/gy [B,N] ->
    (coords [B,N,R,3], final unitaries [B,4,4]) as CUDA tensors."""
    dev = engine.require_cuda()
    as_t = lambda a: a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, float), device=dev)
    gate, gx, gy = as_t(gate), as_t(gx), as_t(gy)
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
        self.construct_basis()
        self.prepare_parameters(0, 0)
        self.fig = None

    # the reference composes N qiskit copies of the one-slice template; here the "circuit" is the parameter table
    def construct_basis(self):
        self._gate = np.array([self.phase_a, self.phase_b, self.phase_c, self.phase_g, self.gc, self.gg, self.gz1, self.gz2],
                              dtype=np.float64)

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
        return engine.pd_trajectory(t(self._gate), t(self._gx), t(self._gy), self.duration_1q, R=R, fold=True, round8=True,
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

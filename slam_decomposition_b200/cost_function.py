"""Cost functionals with the reference's class names (src/slam/cost_function.py:117-221).

``unitary_fidelity(current_u, target_u)`` keeps the reference signature (two 4x4 arrays -> float,
0 = best) and is evaluated on the device; ``cost_kind`` tells the batched kernels which functional
to fuse into the loss/gradient epilogue.  ``BasicCost`` is 1 - |Tr(V^dag U)|/4 (NOT squared).
Out of scope: entanglement costs (3-qubit state vectors), ContinuousUnitaryCostFunction,
LineSegmentDistanceCost.
"""
from __future__ import annotations

from abc import ABC

import numpy as np
import torch

from . import _lib, engine
from . import weyl as _weyl


class UnitaryCostFunction(ABC):
    cost_kind = None  # SlamCostKind for trace-based functionals; None = not fused into K2/K5

    def __init__(self):
        self.normalization = 1

    def unitary_fidelity(self, current_u, target_u):
        raise NotImplementedError

    # batched form used by the optimizer shims: current [B,4,4], target [B,4,4] or [4,4] -> [B] tensor
    def unitary_fidelity_batch(self, current_u: torch.Tensor, target_u: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class _TraceCost(UnitaryCostFunction):
    def _from_abs(self, a):
        raise NotImplementedError

    def unitary_fidelity_batch(self, current_u, target_u):
        cur = _weyl._to_dev(current_u)
        tgt = _weyl._to_dev(target_u)
        # T = Tr(V^dag U) on the device (elementwise conj-multiply-reduce over 16 entries)
        T = (tgt.conj() * cur).sum(dim=(-2, -1))
        return self._from_abs(T.abs())

    def unitary_fidelity(self, current_u, target_u):
        return float(self.unitary_fidelity_batch(current_u, target_u)[0].item())


class BasicCostInverse(_TraceCost):
    cost_kind = _lib.COST_BASIC_INVERSE

    def _from_abs(self, a):
        return a / 4


class BasicCost(_TraceCost):
    cost_kind = _lib.COST_BASIC

    def _from_abs(self, a):
        return 1 - a / 4


class SquareCost(_TraceCost):
    cost_kind = _lib.COST_SQUARE

    def _from_abs(self, a):
        return 1 - (a ** 2 + 4) / 20


def _canonical_batch(c: torch.Tensor) -> torch.Tensor:
    """canonical_gate(c1,c2,c3) for a batch of coordinates, closed form on the device."""
    a, b, cc = (np.pi / 2) * c[:, 0], (np.pi / 2) * c[:, 1], (np.pi / 2) * c[:, 2]
    m = torch.zeros((c.shape[0], 4, 4), dtype=torch.complex128, device=c.device)
    ep, em = torch.exp(1j * cc), torch.exp(-1j * cc)
    m[:, 0, 0] = m[:, 3, 3] = ep * torch.cos(a - b)
    m[:, 0, 3] = m[:, 3, 0] = 1j * ep * torch.sin(a - b)
    m[:, 1, 1] = m[:, 2, 2] = em * torch.cos(a + b)
    m[:, 1, 2] = m[:, 2, 1] = 1j * em * torch.sin(a + b)
    return m


class BasicReducedCost(BasicCost):
    cost_kind = _lib.COST_BASIC_REDUCED

    def unitary_fidelity_batch(self, current_u, target_u):
        ct = _canonical_batch(_weyl.c1c2c3_batch(target_u, round8=True))
        cc = _canonical_batch(_weyl.c1c2c3_batch(current_u, round8=True))
        return super().unitary_fidelity_batch(cc, ct)


class SquareReducedCost(SquareCost):
    cost_kind = _lib.COST_SQUARE_REDUCED

    def unitary_fidelity_batch(self, current_u, target_u):
        ct = _canonical_batch(_weyl.c1c2c3_batch(target_u, round8=True))
        cc = _canonical_batch(_weyl.c1c2c3_batch(current_u, round8=True))
        return super().unitary_fidelity_batch(cc, ct)


class _InvariantCost(UnitaryCostFunction):
    def unitary_fidelity(self, current_u, target_u):
        if (4, 4) != tuple(np.shape(current_u)):
            raise ValueError("Weyl chamber only for 2Q gates")
        return float(self.unitary_fidelity_batch(current_u, target_u)[0].item())


class WeylEuclideanCost(_InvariantCost):
    cost_kind = _lib.COST_WEYL_EUCLIDEAN

    def unitary_fidelity_batch(self, current_u, target_u):
        ct = _weyl.c1c2c3_batch(target_u, round8=True)
        cc = _weyl.c1c2c3_batch(current_u, round8=True)
        return torch.linalg.vector_norm(ct - cc, dim=-1)


class MakhlinEuclideanCost(_InvariantCost):
    cost_kind = _lib.COST_MAKHLIN_EUCLIDEAN

    def unitary_fidelity_batch(self, current_u, target_u):
        gt = _weyl.g1g2g3_batch(target_u, round8=True)
        gc = _weyl.g1g2g3_batch(current_u, round8=True)
        return torch.linalg.vector_norm(gt - gc, dim=-1)


class MakhlinFunctionalCost(_InvariantCost):
    cost_kind = _lib.COST_MAKHLIN_FUNCTIONAL

    def unitary_fidelity_batch(self, current_u, target_u):
        gt = _weyl.g1g2g3_batch(target_u, round8=True)
        gc = _weyl.g1g2g3_batch(current_u, round8=True)
        return ((gt - gc).abs() ** 2).sum(dim=-1)

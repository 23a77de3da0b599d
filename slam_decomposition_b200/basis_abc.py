"""``VariationalTemplate`` base class and ``DataDictEntry`` (reference: src/slam/basis_abc.py).

Kept: the attribute surface ``TemplateOptimizer`` reads (``preseeded``, ``seed``, ``spanning_range``,
``use_polytopes``) and ``target_invariant`` (= Weyl coordinates, computed by the K3 kernel).  The
pickle/KD-tree preseed store of the reference is out of scope (dead unless monodromy polytopes are on,
basis_abc.py:37-43), so ``preseeded`` is always False here.
"""
from __future__ import annotations

from abc import ABC
from dataclasses import dataclass
from random import uniform

import numpy as np


@dataclass
class DataDictEntry:
    success_label: int
    loss_result: float
    Xk: list
    cycles: int


class VariationalTemplate(ABC):
    def __init__(self, preseed: bool = False, use_polytopes: bool = False):
        if use_polytopes:
            raise NotImplementedError("monodromy polytope coverage is out of scope of the CUDA hot path (SURVEY 2.1 #13)")
        if getattr(self, "spanning_range", None) is None:
            raise NotImplementedError
        self.use_polytopes = False
        self.preseeded = False  # reference: preseed and use_polytopes
        self.seed = None
        self.data_dict = {}
        self.coordinate_tree = None

    def eval(self, Xk):
        raise NotImplementedError

    def parameter_guess(self, temperature=0):
        if self.preseeded and self.seed is not None:
            return [el * uniform(1 - 0.05 * temperature, 1 + 0.05 * temperature) for el in self.seed]
        return None

    def assign_seed(self, Xk):
        self.seed = Xk

    def target_invariant(self, target_U):
        """Weyl-chamber coordinates (c1, c2, c3) of a 2Q target, 8-dp rounded as weylchamber.c1c2c3."""
        target_U = np.asarray(target_U)
        if (4, 4) != target_U.shape:
            return (-1, -1, -1, -1)
        from .weyl import c1c2c3

        return c1c2c3(target_U)

    def undo_invariant_transform(self, target_U):
        raise NotImplementedError

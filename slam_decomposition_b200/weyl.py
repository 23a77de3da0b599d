"""Drop-in for the ``weylchamber`` calls on the hot path (``c1c2c3``, ``g1g2g3``, ``J_T_LI``,
``canonical_gate``), evaluated by the K3 CUDA kernel.  Single matrices go through the same batched
kernel as a batch of one; batched variants take/return torch tensors."""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .utils.gates.custom_gates import _canonical_matrix


def _to_dev(U) -> torch.Tensor:
    if isinstance(U, torch.Tensor):
        t = U.to(torch.complex128)
        if not t.is_cuda:
            t = t.to(engine.require_cuda())
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(U, dtype=np.complex128)), device=engine.require_cuda())
    return t.reshape(-1, 4, 4)


def c1c2c3_batch(U, fold: bool = False, round8: bool = False) -> torch.Tensor:
    c, _ = engine.weyl(_to_dev(U), fold=fold, round8=round8)
    return c


def g1g2g3_batch(U, round8: bool = False) -> torch.Tensor:
    _, g = engine.weyl(_to_dev(U), round8=round8, want_c=False, want_g=True)
    return g


def c1c2c3(U, ndigits: int = 8):
    """``weylchamber.c1c2c3(U)`` -> (c1, c2, c3) rounded to 8 decimals."""
    c = c1c2c3_batch(U, round8=(ndigits == 8))[0].tolist()
    if ndigits != 8:
        c = [round(v + 0.0, ndigits) for v in c]
    return tuple(c)


def g1g2g3(U, ndigits: int = 8):
    g = g1g2g3_batch(U, round8=(ndigits == 8))[0].tolist()
    if ndigits != 8:
        g = [round(v + 0.0, ndigits) for v in g]
    return tuple(g)


def J_T_LI(O, U, form: str = "g") -> float:
    """weylchamber ``J_T_LI`` (form 'g'): sum |g_i(O) - g_i(U)|^2 on the 8-dp rounded invariants."""
    if form != "g":
        raise NotImplementedError("only the default form 'g' is used by the reference (cost_function.py:221)")
    g = g1g2g3_batch(torch.stack([_to_dev(O)[0], _to_dev(U)[0]]), round8=True).cpu().numpy()
    return float(np.sum(np.abs(g[0] - g[1]) ** 2))


def canonical_gate(c1, c2, c3) -> np.ndarray:
    """exp(i pi/2 (c1 XX + c2 YY + c3 ZZ)) (constant table; closed form)."""
    return _canonical_matrix(c1, c2, c3)

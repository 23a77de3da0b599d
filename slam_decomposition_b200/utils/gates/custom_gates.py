"""2Q basis-gate objects with the reference's names and constructor signatures.

Mirrors the in-scope classes of ``src/slam/utils/gates/custom_gates.py`` (RiSwapGate :561-606,
ConversionGainGate :163-212, ConversionGainSmushGate :215-257, ConversionGainSmush1QPhaseGate
:260-313, CanonicalGate :384-392, BerkeleyGate :395-400).  Objects only carry parameters; their
matrices are produced by the CUDA kernels (``__array__`` runs ``slam_template_eval`` on a
one-gate template), so there is no host-side expm anywhere.
"""
from __future__ import annotations

from typing import List

import numpy as np

from ...circuit import Gate, Parameter, ParameterValueType, TemplateCircuit

__all__ = [
    "RiSwapGate", "ConversionGainGate", "ConversionGainSmushGate", "ConversionGainSmush1QPhaseGate", "CanonicalGate",
    "BerkeleyGate", "FixedGate", "UnitaryGate", "CXGate", "CZGate", "SwapGate", "iSwapGate", "CPhaseGate", "FSim", "SYC",
]


def _numeric(params) -> bool:
    return all(isinstance(p, (int, float, np.floating, np.integer)) for p in params)


class _DeviceMatrixGate(Gate):
    """A parameterised 2Q gate whose matrix comes from the device kernels."""

    def __array__(self, dtype=None, copy=None):
        if self.is_parameterized():
            raise TypeError(f"{self.name}: cannot take the matrix of a gate with unbound parameters")
        from ... import engine  # local import: gate objects are usable without a GPU until a matrix is needed
        from ...circuit import lower
        import torch

        qc = TemplateCircuit(2)
        qc.append(self, (0, 1))
        desc, _, _ = lower(qc, no_exterior_1q=True)
        dev = engine.require_cuda()
        U = engine.template_eval(desc, torch.zeros((1, 0), dtype=torch.float64, device=dev))
        out = U[0].cpu().numpy()
        return out.astype(dtype) if dtype is not None else out

    def to_matrix(self):
        return self.__array__()


class RiSwapGate(_DeviceMatrixGate):
    r"""n-th root of iSWAP: ``RiSwapGate(1/2)`` is :math:`\sqrt{iSWAP}` (custom_gates.py:561-606)."""

    def __init__(self, alpha: ParameterValueType):
        super().__init__("riswap", 2, [alpha], label="riswap")
        if _numeric(self.params):
            self.duration = self.cost()

    def cost(self):
        return float(self.params[0])

    def fidelity(self):
        if float(self.params[0]) <= (1 / 20):
            return 0
        base = 0.999
        return max(1 - (1 - base) * float(self.params[0]), 0)

    def __str__(self):
        return RiSwapGate.latex_string(self.params)

    @staticmethod
    def latex_string(gate_params=None):
        if gate_params is None:
            return r"$\sqrt[n]{iSwap}$"
        return r"$\sqrt[" + str(int(1 / gate_params[0])) + r"]{iSwap}$"


class ConversionGainGate(_DeviceMatrixGate):
    """exp(-i t (gc H_c(phi_c) + gg H_g(phi_g))).  Positional order (p1, p2, g1, g2, t_el) =
    (phi_c, phi_g, gc, gg, t) -- the reference's positional quirk."""

    def __init__(self, p1: ParameterValueType, p2: ParameterValueType, g1: ParameterValueType, g2: ParameterValueType,
                 t_el: ParameterValueType = 1):
        super().__init__("2QGate", 2, [p1, p2, g1, g2, t_el], "2QGate")
        if _numeric(self.params):
            self.duration = self.cost()
            self.name2 = str(self)

    def __str__(self):
        g1, g2, t = self.params[2], self.params[3], self.params[4]
        return f"2QGate({g1:.8f}, {g2:.8f}, {t:.8f})"

    def cost(self):
        norm = np.pi / 2
        return (abs(self.params[2]) + abs(self.params[3])) * self.params[-1] / norm

    def normalize_duration(self, new_duration):
        old = self.duration
        t = self.params[-1]
        self.params[2] = self.params[2] * t / new_duration
        self.params[3] = self.params[3] * t / new_duration
        self.params[-1] = new_duration
        assert abs(self.cost() - old) < 1e-12


class ConversionGainSmushGate(_DeviceMatrixGate):
    """Time-sliced conversion/gain drive with simultaneous 1Q X drives gx[i], gy[i] (custom_gates.py:215-257)."""

    def __init__(self, pc, pg, gc, gg, gx: List[ParameterValueType], gy: List[ParameterValueType], t_el=1):
        self.xy_len = len(gx)
        assert len(gx) == len(gy)
        self.t_el = t_el
        super().__init__("2QSmushGate", 2, [pc, pg, gc, gg, *gx, *gy, t_el], "2QSmushGate")
        if _numeric(self.params[0:4]):
            self.duration = self.cost()

    def cost(self):
        norm = np.pi / 2
        return (abs(self.params[2]) + abs(self.params[3])) * self.params[-1] / norm


class ConversionGainSmush1QPhaseGate(_DeviceMatrixGate):
    """As the smush gate, plus 1Q drive phases and Z detunings (custom_gates.py:260-313)."""

    def __init__(self, pa, pb, pc, pg, gc, gg, gz1, gz2, gx: List[ParameterValueType], gy: List[ParameterValueType], t_el=1):
        self.xy_len = len(gx)
        assert len(gx) == len(gy)
        self.t_el = t_el
        super().__init__("2QSmushGate1QPhase", 2, [pa, pb, pc, pg, gc, gg, gz1, gz2, *gx, *gy, t_el], "2QSmushGate1QPhase")
        if _numeric(self.params[:8]):
            self.duration = self.cost()

    def cost(self):
        norm = np.pi / 2
        return (abs(self.params[4]) + abs(self.params[5])) * self.params[-1] / norm


class FixedGate(Gate):
    """A constant 2Q gate given by its 4x4 matrix (a table, not a computation)."""

    def __init__(self, name: str, matrix, label=None):
        super().__init__(name, 2, [], label or name)
        m = np.asarray(matrix, dtype=np.complex128)
        if m.shape != (4, 4):
            raise ValueError("FixedGate needs a 4x4 matrix")
        self._matrix = m

    def to_matrix_const(self):
        return self._matrix

    def to_matrix(self):
        return self._matrix.copy()

    def __array__(self, dtype=None, copy=None):
        return self._matrix.astype(dtype) if dtype is not None else self._matrix.copy()


class UnitaryGate(FixedGate):
    def __init__(self, matrix, label=None):
        super().__init__("unitary", matrix, label)


def _canonical_matrix(c1, c2, c3):
    """exp(i pi/2 (c1 XX + c2 YY + c3 ZZ)) in closed form (weylchamber.canonical_gate; the three
    terms commute: XX+YY and XX-YY generate rotations of {01,10} and {00,11}, ZZ is diagonal)."""
    a, b, c = (np.pi / 2) * np.array([c1, c2, c3], dtype=float)
    m = np.zeros((4, 4), dtype=np.complex128)
    m[0, 0] = m[3, 3] = np.exp(1j * c) * np.cos(a - b)
    m[0, 3] = m[3, 0] = 1j * np.exp(1j * c) * np.sin(a - b)
    m[1, 1] = m[2, 2] = np.exp(-1j * c) * np.cos(a + b)
    m[1, 2] = m[2, 1] = 1j * np.exp(-1j * c) * np.sin(a + b)
    return m


class CanonicalGate(FixedGate):
    """``CanonicalGate(alpha, beta, gamma)`` rescales by 2/pi before calling canonical_gate (custom_gates.py:384-389)."""

    def __init__(self, alpha, beta, gamma, name="can"):
        c = [2 * x / np.pi for x in (alpha, beta, gamma)]
        super().__init__(name, _canonical_matrix(*c))
        self.params = [alpha, beta, gamma]
        self.data = self._matrix


class BerkeleyGate(CanonicalGate):
    def __init__(self):
        super().__init__(np.pi / 4, np.pi / 8, 0, name="B")

    def __str__(self):
        return "B"


class CXGate(FixedGate):
    def __init__(self):
        super().__init__("cx", [[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]])


class CZGate(FixedGate):
    def __init__(self):
        super().__init__("cz", np.diag([1, 1, 1, -1]))


class SwapGate(FixedGate):
    def __init__(self):
        super().__init__("swap", [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])


class iSwapGate(FixedGate):
    def __init__(self):
        super().__init__("iswap", [[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])


class CPhaseGate(FixedGate):
    def __init__(self, theta):
        super().__init__("cp", np.diag([1, 1, 1, np.exp(1j * theta)]))
        self.params = [theta]


class FSim(FixedGate):
    """``FSim(theta, phi)`` matrix table (custom_gates.py:489-503)."""

    def __init__(self, theta, phi):
        super().__init__("fsim", [[1, 0, 0, 0], [0, np.cos(theta), -1j * np.sin(theta), 0],
                                  [0, -1j * np.sin(theta), np.cos(theta), 0], [0, 0, 0, np.exp(1j * phi)]])
        self.params = [theta, phi]


class SYC(FSim):
    def __init__(self):
        super().__init__(np.pi / 2, np.pi / 6)

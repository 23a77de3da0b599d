"""2Q Hamiltonian factories (reference: src/slam/hamiltonian.py:84-182).  ``construct_U`` keeps the
reference's static-factory signature and positional conventions; the exponentials are evaluated by the
device kernels (closed form for conversion/gain, per-slice scaling-and-squaring for the smush families),
not by qutip/scipy ``expm``.  3Q Hamiltonians (Circulator, DeltaConversionGain) are out of scope.
"""
from __future__ import annotations

from abc import ABC
from inspect import signature

import numpy as np


class Hamiltonian(ABC):
    @staticmethod
    def construct_U(*args):
        raise NotImplementedError

    @classmethod
    def n_params(cls) -> int:
        return len(signature(cls.construct_U).parameters)


def _gate_matrix(gate) -> np.ndarray:
    return np.asarray(gate.__array__(), dtype=np.complex128)


class ConversionGainPhaseHamiltonian(Hamiltonian):
    @staticmethod
    def construct_U(gc, gg, phi_c, phi_g, t=1):
        # positional quirk kept: the four leading arguments are forwarded as (phi_c, phi_g, gc, gg)
        # (hamiltonian.py:95-111)
        from .utils.gates.custom_gates import ConversionGainGate

        return _gate_matrix(ConversionGainGate(float(gc), float(gg), float(phi_c), float(phi_g), float(t)))


class ConversionGainHamiltonian(Hamiltonian):
    @staticmethod
    def construct_U(gc, gg):
        from .utils.gates.custom_gates import ConversionGainGate

        return _gate_matrix(ConversionGainGate(0.0, 0.0, float(gc), float(gg), 1.0))


class ConversionGainSmush(Hamiltonian):
    @staticmethod
    def construct_U(phi_c, phi_g, gc, gg, gxvector, gyvector, t=1):
        from .utils.gates.custom_gates import ConversionGainSmushGate

        assert len(gxvector) == len(gyvector)
        return _gate_matrix(ConversionGainSmushGate(float(phi_c), float(phi_g), float(gc), float(gg),
                                                    [float(v) for v in gxvector], [float(v) for v in gyvector], float(t)))


class ConversionGainSmush1QPhase(Hamiltonian):
    @staticmethod
    def construct_U(phi_a, phi_b, phi_c, phi_g, gc, gg, gz1, gz2, gxvector, gyvector, t=1):
        from .utils.gates.custom_gates import ConversionGainSmush1QPhaseGate

        assert len(gxvector) == len(gyvector)
        return _gate_matrix(ConversionGainSmush1QPhaseGate(
            float(phi_a), float(phi_b), float(phi_c), float(phi_g), float(gc), float(gg), float(gz1), float(gz2),
            [float(v) for v in gxvector], [float(v) for v in gyvector], float(t)))

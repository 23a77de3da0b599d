"""``CircuitTemplate`` and ``HamiltonianTemplate`` (reference: src/slam/basis.py:24-170).

A template is: 1Q layer, then k x (2Q basis gate, 1Q layer); ``build(k)`` creates the symbolic circuit
with parameters ``P0, P1, ...`` in creation order (q0 then q1 per layer) and lowers it to the
``SlamTemplateDesc`` the CUDA kernels read.  ``eval(Xk)`` binds ``Xk`` in ``circuit.parameters`` order
(sorted by name) and returns the 4x4 unitary computed by ``slam_template_eval``.
``MixedOrderBasisCircuitTemplate`` (monodromy polytopes) is out of scope.
"""
from __future__ import annotations

from itertools import cycle
from typing import Sequence

import numpy as np
import torch

from . import engine
from .basis_abc import VariationalTemplate
from .circuit import Parameter, TemplateCircuit, lower
from .hamiltonian import Hamiltonian
from .utils.gates.custom_gates import RiSwapGate


class _Cycler:
    """``itertools.cycle`` with a readable position.  The reference creates its gate / edge cycles ONCE in ``__init__``
    (basis.py:68-72, basisv2.py:61-64), so they keep advancing across successive ``build()`` calls: with more than one base
    gate the sequence a build sees depends on how many gates earlier builds consumed.  The position is exposed so that the
    optimiser's probing builds (which the reference does not do) can be undone."""

    def __init__(self, items):
        self.items = list(items)
        self.pos = 0

    def __next__(self):
        v = self.items[self.pos % len(self.items)]
        self.pos += 1
        return v

    def __iter__(self):
        return self


class _CircuitTemplateBase(VariationalTemplate):
    """Shared machinery of CircuitTemplate / CircuitTemplateV2: device evaluation of the built circuit."""

    vz_only = False

    def _lower(self):
        self.desc, self.param_names, numeric = lower(self.circuit, vz_only=self.vz_only,
                                                     no_exterior_1q=bool(self.no_exterior_1q))
        if numeric.size:
            raise ValueError("template circuits must keep their 1Q angles symbolic")

    # ---- single evaluation (reference signature) ------------------------------------------------
    def eval(self, Xk) -> np.ndarray:
        """``Operator(circuit.assign_parameters(Xk)).data`` -> complex128 [4,4] (fresh array)."""
        Xk = np.asarray(Xk, dtype=np.float64).reshape(1, -1)
        dev = engine.require_cuda()
        U = engine.template_eval(self.desc, torch.as_tensor(Xk, device=dev))
        return U[0].cpu().numpy()

    # ---- batched evaluation (what the GPU is for) -----------------------------------------------
    def eval_batch(self, X: torch.Tensor) -> torch.Tensor:
        """X [B,P] (CUDA float64) -> U [B,4,4] complex128 on the device."""
        return engine.template_eval(self.desc, X)

    def assign_Xk(self, Xk):
        return self.circuit.assign_parameters({p: float(v) for p, v in zip(self.circuit.parameters, Xk)})

    def _reset(self):
        self.cycles = 0
        self.circuit = TemplateCircuit(self.n_qubits)
        self._p_index = 0
        self._q_index = 0

    def _init_cycles(self):
        self.gate_2q_base = _Cycler(self._base_gates)
        self.gate_2q_edges = _Cycler([_Cycler(e) for e in self._edge_params])

    def cycle_state(self):
        """Positions of the gate / edge cycles (see _Cycler)."""
        return (self.gate_2q_base.pos, self.gate_2q_edges.pos, tuple(c.pos for c in self.gate_2q_edges.items))

    def set_cycle_state(self, state):
        self.gate_2q_base.pos, self.gate_2q_edges.pos = state[0], state[1]
        for c, pos in zip(self.gate_2q_edges.items, state[2]):
            c.pos = pos

    def _next_1q(self, n):
        out = [Parameter(f"P{self._p_index + j}") for j in range(n)]
        self._p_index += n
        return out

    def build(self, n_repetitions):
        self._reset()
        if n_repetitions <= 0:
            raise ValueError()
        for i in range(n_repetitions):
            self._build_cycle(initial=(i == 0), final=(i == n_repetitions - 1))
        self._lower()


class HamiltonianTemplate(VariationalTemplate):
    """``eval(Xk) = h.construct_U(*Xk)`` (basis.py:24-48)."""

    def __init__(self, h: Hamiltonian):
        self.h = h
        self.n_qubits = 2
        self.spanning_range = range(1)
        self.using_bounds = False
        self.using_constraints = False
        self.bounds_list = None
        self.constraint_func = None
        super().__init__(preseed=False, use_polytopes=False)

    def get_spanning_range(self, target_u):
        return range(1, 2)

    x0_bounds = (0.0, 1.0)  # np.random.random(p_len) (basis.py:48)

    def build(self, n_repetitions):
        """Nothing to extend (optimizer.py:240-248 only builds circuit templates).  For the scalar-argument
        conversion/gain Hamiltonians the template is lowered to a one-gate descriptor whose parameters H0, H1, ... are
        construct_U's positional arguments, so the device optimisers can run it."""
        from .hamiltonian import ConversionGainHamiltonian, ConversionGainPhaseHamiltonian
        from .utils.gates.custom_gates import ConversionGainGate

        h = self.h if isinstance(self.h, type) else type(self.h)
        n = h.n_params()
        ps = [Parameter(f"H{j}") for j in range(n)]
        if issubclass(h, ConversionGainPhaseHamiltonian):
            gate = ConversionGainGate(ps[0], ps[1], ps[2], ps[3], ps[4])  # positional quirk kept
        elif issubclass(h, ConversionGainHamiltonian):
            gate = ConversionGainGate(0.0, 0.0, ps[0], ps[1], 1.0)
        else:
            raise NotImplementedError(f"{h.__name__}: vector-argument Hamiltonians cannot be driven by a flat Xk "
                                      "(construct_U(*Xk) fails in the reference as well)")
        self.circuit = TemplateCircuit(2)
        self.circuit.append(gate, (0, 1))
        self.no_exterior_1q = True
        self.desc, self.param_names, _ = lower(self.circuit, no_exterior_1q=True)

    def eval(self, Xk):
        return np.asarray(self.h.construct_U(*Xk))

    def parameter_guess(self, t=1):
        parent = super().parameter_guess(t)
        if parent is not None:
            return parent
        return np.random.random(self.h.n_params())


class CircuitTemplate(_CircuitTemplateBase):
    def __init__(self, n_qubits=2, base_gates=None, edge_params=None, no_exterior_1q=False, use_polytopes=False,
                 maximum_span_guess=5, preseed=False):
        if n_qubits != 2:
            raise NotImplementedError("the CUDA hot path covers 2-qubit templates (4x4 unitaries)")
        base_gates = [RiSwapGate(1 / 2)] if base_gates is None else list(base_gates)
        edge_params = [[(0, 1)]] if edge_params is None else edge_params
        self.n_qubits = n_qubits
        self.no_exterior_1q = no_exterior_1q
        self._base_gates = base_gates
        self._edge_params = edge_params
        self.using_bounds = False
        self.bounds_list = None
        self.using_constraints = False
        self.constraint_func = None
        self.spanning_range = None
        if not use_polytopes:
            self.spanning_range = range(1, maximum_span_guess + 1)
            self.coverage = None
        super().__init__(preseed=preseed, use_polytopes=use_polytopes)
        self._init_cycles()
        self._reset()
        self.trotter = False

    def get_spanning_range(self, target_u):
        return self.spanning_range

    def parameter_guess(self, t=0):
        """np array of U[0, 2pi) values, one per parameter (basis.py:106-111)."""
        parent = super().parameter_guess(t)
        if parent is not None:
            return parent
        return np.random.random(len(self.circuit.parameters)) * 2 * np.pi

    x0_bounds = (0.0, 2 * np.pi)

    def _build_cycle(self, initial=False, final=False):
        if initial and not self.no_exterior_1q:
            for qubit in range(self.n_qubits):
                self.circuit.u(*self._next_1q(3), qubit)
        gate = next(self.gate_2q_base)
        edge = next(next(self.gate_2q_edges))
        self.circuit.append(gate, edge)
        if not (final and self.no_exterior_1q):
            for qubit in edge:
                self.circuit.u(*self._next_1q(3), qubit)
        self.cycles += 1

"""``CircuitTemplateV2`` (reference: src/slam/basisv2.py:27-299): the 2Q gate is a *class or callable*
that receives fresh ``Q{j}`` parameters on every repetition (continuous 2Q search, smush vectors via
``param_vec_expand``), optional RZ-only 1Q layers, per-parameter bounds and a circuit-cost constraint.
"""
from __future__ import annotations

from inspect import signature

import numpy as np

from .basis import _CircuitTemplateBase
from .circuit import Parameter
from .utils.gates.custom_gates import ConversionGainGate, RiSwapGate


class CircuitTemplateV2(_CircuitTemplateBase):
    def __init__(self, n_qubits=2, base_gates=None, edge_params=None, no_exterior_1q=False, use_polytopes=False,
                 maximum_span_guess=5, preseed=False, vz_only=False, param_vec_expand=None):
        if n_qubits != 2:
            raise NotImplementedError("the CUDA hot path covers 2-qubit templates (4x4 unitaries)")
        base_gates = [RiSwapGate] if base_gates is None else list(base_gates)
        edge_params = [[(0, 1)]] if edge_params is None else edge_params
        self.n_qubits = n_qubits
        self.no_exterior_1q = no_exterior_1q
        self.param_vec_expand = param_vec_expand
        if self.param_vec_expand is not None:
            assert len(base_gates) == 1
        self.vz_only = bool(vz_only)
        self._base_gates = base_gates
        self._edge_params = edge_params
        self.bounds = {}
        self.bounds_list = []
        self.constraint_func = None
        self.using_bounds = False
        self.using_constraints = False
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

    def reconstruct(self, ret):
        self.build(ret.cycles)
        print("Cost:", self.circuit_cost(Xk=ret.Xk))
        return self.assign_Xk(ret.Xk)

    # ---- abstract circuit score used for constraints (basisv2.py:98-127) -----------------------
    def circuit_cost(self, Xk):
        cost = 0
        for ins in self.assign_Xk(Xk):
            g = ins.operation
            c = 0.0
            if g.name == "riswap":
                c = RiSwapGate(g.params[0]).cost()
            elif g.name in ("2QGate", "2QSmushGate"):
                a = [float(el) for el in g.params]
                c = ConversionGainGate(a[0], a[1], a[2], a[3], a[-1]).cost()
            cost += c
        return cost

    def circuit_fidelity(self, Xk):
        fidelity = 1.0
        for ins in self.assign_Xk(Xk):
            g = ins.operation
            c = 1.0
            if g.name == "riswap":
                c = RiSwapGate(g.params[0]).cost()  # (sic) the reference multiplies by cost(), basisv2.py:137-140
            fidelity = fidelity * c
        return fidelity

    default_bound = (-4 * np.pi, 4 * np.pi)

    def parameter_guess(self, t=0):
        """One uniform draw per parameter from its bound, default (-4pi, 4pi) (basisv2.py:148-171)."""
        parent = super().parameter_guess(t)
        if parent is not None:
            return parent
        random_list = []
        self.bounds_list = []
        for parameter in self.circuit.parameters:
            cbound = self.bounds.get(parameter.name, self.default_bound)
            self.bounds_list.append(cbound)
            if cbound is None:
                cbound = self.default_bound
            random_list.append(np.random.uniform(cbound[0], cbound[1], 1)[0])
        if not self.using_bounds:
            self.bounds_list = None
        return random_list

    def x0_bound_arrays(self):
        """(lo[P], hi[P]) in API order for device-side initial-point generation."""
        lo, hi = [], []
        for parameter in self.circuit.parameters:
            b = self.bounds.get(parameter.name, self.default_bound) or self.default_bound
            b = (self.default_bound[0] if b[0] is None else b[0], self.default_bound[1] if b[1] is None else b[1])
            b = (float(b[0]), float(b[1]))
            lo.append(min(b))
            hi.append(max(b))
        return np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)

    def add_bound(self, parameter_name, max=None, min=None):
        self.bounds[parameter_name] = (min, max)
        if not any(p.name == parameter_name for p in self.circuit.parameters):
            raise ValueError("Parameter Name not found")
        self.using_bounds = True

    def set_constraint(self, param_max_cost):
        self.constraint_func = {"type": "ineq", "fun": lambda x: param_max_cost - self.circuit_cost(x)}
        self.constraint_max = float(param_max_cost)  # read by the device optimiser (augmented-Lagrangian term)
        self.using_constraints = True

    def remove_constraint(self):
        self.constraint_func = None
        self.using_constraints = False

    def _next_2q(self, n):
        out = [Parameter(f"Q{self._q_index + j}") for j in range(n)]
        self._q_index += n
        return out

    def _build_cycle(self, initial=False, final=False):
        n1 = 1 if self.vz_only else 3

        def one_q(qubit):
            ps = self._next_1q(n1)
            if self.vz_only:
                self.circuit.rz(ps[0], qubit)
            else:
                self.circuit.u(*ps, qubit)

        if initial and not self.no_exterior_1q:
            for qubit in range(self.n_qubits):
                one_q(qubit)
        gate = next(self.gate_2q_base)
        edge = next(next(self.gate_2q_edges))
        num2qparams = len(signature(gate).parameters)
        if self.param_vec_expand is not None:
            num2qparams = sum(self.param_vec_expand)
        gate_instance = gate(*self._next_2q(num2qparams))
        self.circuit.append(gate_instance, edge)
        if not (final and self.no_exterior_1q):
            for qubit in edge:
                one_q(qubit)
        self.cycles += 1

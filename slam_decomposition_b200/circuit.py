"""Minimal symbolic circuit used by the template shims.

The reference keeps a template as a qiskit ``QuantumCircuit`` holding ``Parameter`` objects
(src/slam/basis.py:118-169).  qiskit is not a dependency here: this module provides just enough of
that surface (``Parameter``, ``circuit.parameters`` sorted by name, ``assign_parameters``,
iteration over instructions) for the template classes, and lowers a circuit to the
``SlamTemplateDesc`` the CUDA kernels consume.  No numerics happen here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import SLAM_MAX_K, SLAM_MAX_SLOTS, SlamTemplateDesc


class Parameter:
    """Named symbolic parameter (stand-in for ``qiskit.circuit.Parameter``)."""

    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = str(name)

    def __repr__(self):
        return f"Parameter({self.name})"

    def __str__(self):
        return self.name

    def __eq__(self, other):
        return isinstance(other, Parameter) and other.name == self.name

    def __hash__(self):
        return hash(("Parameter", self.name))


ParameterValueType = Union[Parameter, float, int]


class Gate:
    """Base of every gate object: a name, a qubit count and a flat ``params`` list whose entries
    are floats or :class:`Parameter` (same attribute names as ``qiskit.circuit.Gate``)."""

    def __init__(self, name: str, num_qubits: int, params: Sequence[ParameterValueType], label: str | None = None):
        self.name = name
        self.num_qubits = num_qubits
        self.params = list(params)
        self.label = label

    def is_parameterized(self) -> bool:
        return any(isinstance(p, Parameter) for p in self.params)

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(map(str, self.params))})"


class UGate(Gate):
    def __init__(self, theta, phi, lam):
        super().__init__("u", 1, [theta, phi, lam])


class RZGate(Gate):
    def __init__(self, lam):
        super().__init__("rz", 1, [lam])


@dataclass
class CircuitInstruction:
    operation: Gate
    qubits: Tuple[int, ...]

    def __iter__(self):  # old qiskit tuple-iteration `for instr, qargs, cargs in qc`
        yield self.operation
        yield self.qubits
        yield ()

    def __getitem__(self, i):
        return (self.operation, self.qubits, ())[i]


class TemplateCircuit:
    """Ordered list of 1Q/2Q instructions on ``n_qubits`` (2 for everything on the hot path)."""

    def __init__(self, n_qubits: int = 2):
        self.num_qubits = n_qubits
        self.data: List[CircuitInstruction] = []
        self._bound: Dict[str, float] | None = None  # set on circuits returned by assign_parameters

    # -- construction ----------------------------------------------------------------------
    def u(self, theta, phi, lam, qubit: int):
        self.data.append(CircuitInstruction(UGate(theta, phi, lam), (qubit,)))

    def rz(self, lam, qubit: int):
        self.data.append(CircuitInstruction(RZGate(lam), (qubit,)))

    def append(self, gate: Gate, qargs: Sequence[int]):
        self.data.append(CircuitInstruction(gate, tuple(qargs)))

    def __iter__(self):
        return iter(self.data)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return self.data[i]

    # -- parameters ------------------------------------------------------------------------
    @property
    def parameters(self) -> List[Parameter]:
        """Unbound parameters sorted by name string, as qiskit's ``ParameterView`` orders them
        (``P0, P1, P10, ..., P2, ...`` then ``Q...``)."""
        seen: Dict[str, Parameter] = {}
        for ins in self.data:
            for p in ins.operation.params:
                if isinstance(p, Parameter):
                    seen.setdefault(p.name, p)
        return [seen[n] for n in sorted(seen)]

    @property
    def num_parameters(self) -> int:
        return len(self.parameters)

    def assign_parameters(self, mapping) -> "TemplateCircuit":
        """Return a bound copy.  ``mapping``: dict Parameter->value, or a sequence in API order."""
        if not isinstance(mapping, dict):
            mapping = dict(zip(self.parameters, mapping))
        values = {(k.name if isinstance(k, Parameter) else str(k)): v for k, v in mapping.items()}
        out = TemplateCircuit(self.num_qubits)
        for ins in self.data:
            g = ins.operation
            new = object.__new__(type(g))
            new.__dict__.update(g.__dict__)
            new.params = [values.get(p.name, p) if isinstance(p, Parameter) else p for p in g.params]
            out.data.append(CircuitInstruction(new, ins.qubits))
        return out

    bind_parameters = assign_parameters

    def draw(self, *_, **__) -> str:
        lines = []
        for ins in self.data:
            lines.append(f"{ins.operation.name}({', '.join(str(p) for p in ins.operation.params)}) q{list(ins.qubits)}")
        return "\n".join(lines)


# ------------------------------------------------------------------------------------------
# lowering to the C descriptor
# ------------------------------------------------------------------------------------------

_KIND_OF = {"riswap": _lib.GATE_RISWAP, "2QGate": _lib.GATE_CG, "2QSmushGate": _lib.GATE_SMUSH,
            "2QSmushGate1QPhase": _lib.GATE_SMUSH_1QPHASE}


def lower(circuit: TemplateCircuit, vz_only: bool = False, no_exterior_1q: bool = False):
    """circuit -> (SlamTemplateDesc, bound_values).

    ``bound_values`` is a float64 vector in API order if every parameter is already bound to a
    number *and no symbolic parameter remains* ... otherwise ``None``.  Fully numeric circuits are
    lowered with their 1Q angles re-expressed as parameters (x = the numbers), so the same kernels
    evaluate them.
    """
    if circuit.num_qubits != 2:
        raise NotImplementedError("the CUDA path evaluates 2-qubit templates (4x4 unitaries) only")
    names = [p.name for p in circuit.parameters]
    numeric: List[float] = []  # numeric 1Q angles become pseudo-parameters appended after the symbolic ones
    index = {n: i for i, n in enumerate(names)}

    def pidx(v) -> int:
        if isinstance(v, Parameter):
            return index[v.name]
        numeric.append(float(v))
        return len(names) + len(numeric) - 1

    desc = SlamTemplateDesc()
    for i in range(SLAM_MAX_K + 1):
        for s in range(6):
            desc.p1q[i][s] = -1
    for g in range(SLAM_MAX_K):
        for s in range(SLAM_MAX_SLOTS):
            desc.slot_param[g][s] = -1
    layer = 0  # 1Q gates seen so far belong to this layer (= number of 2Q gates seen)
    kind = None
    T = 0
    n_slots = 0
    for ins in circuit.data:
        g = ins.operation
        if g.num_qubits == 1:
            q = ins.qubits[0]
            if layer > SLAM_MAX_K:
                raise ValueError("too many template repetitions")
            if g.name == "u":
                if vz_only:
                    raise ValueError("vz_only template holds a U gate")
                if desc.p1q[layer][3 * q] != -1:
                    raise ValueError("two 1Q gates on the same qubit within one layer")
                for j in range(3):
                    desc.p1q[layer][3 * q + j] = pidx(g.params[j])
            elif g.name == "rz":
                if not vz_only:
                    raise ValueError("RZ gate in a U3 template; build the template with vz_only=True")
                desc.p1q[layer][3 * q] = pidx(g.params[0])
            else:
                raise NotImplementedError(f"1Q gate {g.name}")
        else:
            if tuple(ins.qubits) != (0, 1):
                raise NotImplementedError("2Q gates must act on edge (0, 1)")
            if layer >= SLAM_MAX_K:
                raise ValueError(f"template has more than {SLAM_MAX_K} 2Q gates")
            if g.name in _KIND_OF:
                this_kind = _KIND_OF[g.name]
                slots = list(g.params)
                this_T = getattr(g, "xy_len", 0)
            elif hasattr(g, "to_matrix_const"):
                this_kind = _lib.GATE_FIXED
                slots = []
                this_T = 0
                m = np.ascontiguousarray(g.to_matrix_const(), dtype=np.complex128).reshape(16)
                flat = m.view(np.float64)
                if kind is not None and any(desc.fixed_gate[j] != flat[j] for j in range(32)):
                    raise NotImplementedError("templates mixing different fixed 2Q gates")
                for j in range(32):
                    desc.fixed_gate[j] = float(flat[j])
            else:
                raise NotImplementedError(f"2Q gate {g.name} has no CUDA lowering")
            if kind is None:
                kind, T, n_slots = this_kind, this_T, len(slots)
            elif (kind, T, n_slots) != (this_kind, this_T, len(slots)):
                raise NotImplementedError("templates mixing different 2Q gate families")
            if len(slots) > SLAM_MAX_SLOTS:
                raise ValueError("2Q gate has too many scalar slots")
            for s, v in enumerate(slots):
                if isinstance(v, Parameter):
                    desc.slot_param[layer][s] = index[v.name]
                else:
                    desc.slot_const[layer][s] = float(v)
            layer += 1
    if layer == 0:
        raise ValueError("template holds no 2Q gate")  # build(n_repetitions <= 0) raises in the reference
    desc.gate_kind = kind
    desc.k = layer
    desc.T = T
    desc.n_slots = n_slots
    desc.n_params = len(names) + len(numeric)
    desc.no_exterior_1q = int(no_exterior_1q)
    desc.vz_only = int(vz_only)
    if desc.n_params > _lib.SLAM_MAX_PARAMS:
        raise ValueError("too many parameters")
    return desc, names, np.asarray(numeric, dtype=np.float64)

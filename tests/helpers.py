"""Build the same template twice: as the product's TemplateCircuit/SlamTemplateDesc and as the oracle's
OracleTemplate, so parity tests feed identical Xk vectors to both."""
import numpy as np

from oracle import OracleTemplate
from slam_decomposition_b200.circuit import Parameter, TemplateCircuit, lower
from slam_decomposition_b200.utils.gates.custom_gates import (
    ConversionGainGate,
    ConversionGainSmush1QPhaseGate,
    ConversionGainSmushGate,
    FixedGate,
    RiSwapGate,
)

BASES = {  # gc, gg, t   (parallel_drive_volume.py:91-96)
    "iSwap": (np.pi / 2, 0.0, 1.0),
    "sqiSwap": (np.pi / 2, 0.0, 0.5),
    "CNOT": (np.pi / 4, np.pi / 4, 1.0),
    "sqCNOT": (np.pi / 4, np.pi / 4, 0.5),
    "B": (3 * np.pi / 8, np.pi / 8, 1.0),
    "sqB": (3 * np.pi / 8, np.pi / 8, 0.5),
}


def _gate(kind, vals, T):
    if kind == "riswap":
        return RiSwapGate(vals[0])
    if kind == "cg":
        return ConversionGainGate(*vals)
    if kind == "smush":
        return ConversionGainSmushGate(vals[0], vals[1], vals[2], vals[3], vals[4:4 + T], vals[4 + T:4 + 2 * T], vals[-1])
    if kind == "smush1q":
        return ConversionGainSmush1QPhaseGate(*vals[:8], vals[8:8 + T], vals[8 + T:8 + 2 * T], vals[-1])
    raise ValueError(kind)


def make_pair(kind="riswap", slots=(0.5,), k=3, T=0, no_exterior_1q=False, vz_only=False, fixed=None):
    """-> (desc, oracle_template).  `slots`: floats or "Q" (fresh 2Q parameter), as OracleTemplate."""
    orc = OracleTemplate(kind, tuple(slots), k=k, T=T, no_exterior_1q=no_exterior_1q, vz_only=vz_only, fixed=fixed)
    qc = TemplateCircuit(2)
    p = 0
    q = 0
    n1 = 1 if vz_only else 3

    def layer():
        nonlocal p
        for qubit in (0, 1):
            ps = [Parameter(f"P{p + j}") for j in range(n1)]
            p += n1
            if vz_only:
                qc.rz(ps[0], qubit)
            else:
                qc.u(*ps, qubit)

    for i in range(k):
        if i == 0 and not no_exterior_1q:
            layer()
        if kind == "fixed":
            qc.append(FixedGate("fixed", fixed), (0, 1))
        else:
            vals = []
            for s in slots:
                if isinstance(s, str):
                    vals.append(Parameter(f"Q{q}"))
                    q += 1
                else:
                    vals.append(float(s))
            qc.append(_gate(kind, vals, T), (0, 1))
        if not (i == k - 1 and no_exterior_1q):
            layer()
    desc, names, numeric = lower(qc, vz_only=vz_only, no_exterior_1q=no_exterior_1q)
    assert numeric.size == 0
    assert names == orc.names_sorted, (names, orc.names_sorted)
    return desc, orc


def b11_case(kats):
    """KAT B11 (local_smush_test.ipynb cell 5): the slot tuple of the solved template and its parameter values by name."""
    t = kats["B11"]["template"]
    slots = tuple(t["slots"])
    vals = {}
    p = 0
    for tri in kats["B11"]["u3_triples"]:
        for v in tri:
            vals[f"P{p}"] = v
            p += 1
    for q, v in enumerate(list(kats["B11"]["gx"]) + list(kats["B11"]["gy"])):
        vals[f"Q{q}"] = v
    return slots, t["k"], t["T"], vals

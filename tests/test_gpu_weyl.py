"""K3 parity: Weyl-chamber coordinates and Makhlin invariants vs the oracle (weylchamber algorithm)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from slam_decomposition_b200 import engine

pytestmark = pytest.mark.gpu
TOL = 1e-10  # north_star: Weyl coordinates within 1e-10 absolute
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.complex128), device="cuda")


def _coords(U, **kw):
    c, g = engine.weyl(_dev(U), want_g=True, **kw)
    return c.cpu().numpy(), g.cpu().numpy()


def test_golden_gates_rounded():
    named = {
        "CX": O.CNOT, "SWAP": O.SWAP, "I": np.eye(4), "iSWAP": O.ISWAP, "sqrt_iSWAP": O.riswap(0.5),
        "B": O.berkeley(), "FSIM": O.fsim(KATS["B5"]["fsim_theta"], KATS["B5"]["fsim_phi"]),
    }
    U = np.stack(list(named.values()))
    c, g = _coords(U, round8=True)
    for i, (name, M) in enumerate(named.items()):
        assert tuple(c[i]) == O.c1c2c3(M), name
        assert tuple(g[i]) == O.g1g2g3(M), name
    assert tuple(c[0]) == tuple(KATS["B4"]["CX"]) and tuple(c[1]) == tuple(KATS["B4"]["SWAP"])
    assert tuple(c[6]) == tuple(KATS["B5"]["c1c2c3"])
    for name in ("I", "iSWAP", "sqrt_iSWAP"):
        i = list(named).index(name)
        assert tuple(g[i]) == tuple(KATS["B10"][name])


def test_basis_gates_b3():
    for name, gd in KATS["B3"]["gates"].items():
        G = O.conversion_gain(0, 0, gd["gc"], gd["gg"], gd["t"])
        c, _ = _coords(G[None], fold=True, round8=True)
        assert np.allclose(c[0], gd["c1c2c3_folded"], atol=1e-8), name


def test_haar_batch_unrounded():
    rng = np.random.default_rng(0)
    U = O.haar_unitary(rng, 5000)
    c, g = _coords(U)
    co = O.c1c2c3_raw(U)
    go = O.g1g2g3_raw(U)
    # compare folded coordinates (raw c1 mirrors when c3 crosses 0: SURVEY hard parts)
    d = np.abs(O.fold_c1(c) - O.fold_c1(co)).max(axis=1)
    assert np.sum(d > TOL) == 0, d.max()
    assert np.abs(g - go).max() < TOL
    cf, _ = _coords(U, fold=True)
    assert np.abs(cf - O.fold_c1(co)).max() < TOL


def _dress(rng, M, n):
    out = []
    for _ in range(n):
        k1 = np.kron(O.u3(*rng.uniform(0, 7, 3)), O.u3(*rng.uniform(0, 7, 3)))
        k2 = np.kron(O.u3(*rng.uniform(0, 7, 3)), O.u3(*rng.uniform(0, 7, 3)))
        out.append(np.exp(1j * rng.uniform(0, 7)) * (k1 @ M @ k2))
    return out


def _assert_matches_oracle_or_class(Us, c, who=None):
    """Device vs oracle at TOL.  weylchamber's own algorithm (which the oracle restates literally) is unstable exactly
    at the SWAP corner: all four eigenphases sit on its `two_S <= -0.5` branch point and round-off can send it to the
    out-of-chamber representative (1.5, -0.5, 0.5) (the reference works around this, speed_limit_pass.py:369-377).
    Rows where the oracle leaves the chamber are therefore checked against the invariants instead: the device
    coordinates must lie in the chamber and canonical_gate(c) must have the Makhlin invariants of U."""
    co = O.fold_c1(O.c1c2c3_raw(Us))
    bad = np.abs(c - co).max(axis=1) >= TOL
    for i in np.nonzero(bad)[0]:
        in_chamber = -1e-12 <= co[i, 2] <= co[i, 1] + 1e-12 and co[i, 1] <= co[i, 0] + 1e-12 and co[i, 0] <= 0.5 + 1e-12
        assert not in_chamber, (i, c[i], co[i])  # a genuine disagreement
        assert -1e-12 <= c[i, 2] <= c[i, 1] + 1e-12 <= c[i, 0] + 2e-12 <= 0.5 + 3e-12, c[i]
        assert np.abs(np.array(O.g1g2g3_raw(O.canonical_gate(*c[i]))) - np.array(O.g1g2g3_raw(Us[i]))).max() < 1e-9
        assert np.abs(c[i] - 0.5).max() < 1e-3, c[i]  # only ever near the SWAP corner
    return int(bad.sum())


CLASSES = {  # chamber points every basis gate of this domain sits on (parallel_drive_volume.py:91-96 and the chamber corners)
    "I": (0.0, 0.0, 0.0), "CNOT": (0.5, 0.0, 0.0), "SWAP": (0.5, 0.5, 0.5), "iSWAP": (0.5, 0.5, 0.0),
    "sqiSWAP": (0.25, 0.25, 0.0), "B": (0.5, 0.25, 0.0), "sqCNOT": (0.25, 0.0, 0.0), "sqB": (0.25, 0.125, 0.0),
}


def test_degenerate_and_locally_equivalent_inputs(capsys):
    """Degenerate spectra (identity / CNOT / SWAP / iSWAP / sqrt-iSWAP / B / sqCNOT / sqB classes) dressed with random
    local gates and a global phase: the folded coordinates must equal the class point to north_star's 1e-10 (the
    eigenphases of the unitary m = U_B^T U_B are perfectly conditioned; the joint-Jacobi kernel reaches ~1e-15 here)."""
    rng = np.random.default_rng(1)
    base = {"I": np.eye(4), "CNOT": O.CNOT, "SWAP": O.SWAP, "iSWAP": O.ISWAP, "sqiSWAP": O.riswap(0.5), "B": O.berkeley(),
            "sqCNOT": O.conversion_gain(0, 0, np.pi / 4, np.pi / 4, 0.5), "sqB": O.conversion_gain(0, 0, 3 * np.pi / 8, np.pi / 8, 0.5)}
    Us, ref, who = [], [], []
    for name, M in base.items():
        assert np.abs(O.fold_c1(O.c1c2c3_raw(M)) - np.array(CLASSES[name])).max() < 1e-15, name
        for U in _dress(rng, M, 200):
            Us.append(U)
            ref.append(CLASSES[name])
            who.append(name)
    Us, ref, who = np.stack(Us), np.array(ref), np.array(who)
    c, g = _coords(Us, fold=True)
    err = np.abs(c - ref).max(axis=1)
    with capsys.disabled():
        print("\n  max |dc| per class: " + ", ".join(f"{n}={err[who == n].max():.1e}" for n in base))
    assert err.max() < TOL
    _assert_matches_oracle_or_class(Us, c)
    assert np.abs(g - O.g1g2g3_raw(Us)).max() < TOL


def test_near_degenerate_inputs_along_chamber_edges():
    """Class points offset by 1e-12 .. 1e-4 along every chamber edge direction (near-degenerate eigenphase pairs, and
    points just outside the chamber that map back in), dressed with local gates: device vs oracle at 1e-10."""
    rng = np.random.default_rng(3)
    dirs = [(1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 1, 0), (0, 1, 1), (0, 0, 1)]
    dirs = dirs + [tuple(-v for v in d) for d in dirs]
    Us = []
    for p in CLASSES.values():
        for eps in (1e-12, 1e-9, 1e-8, 1e-7, 1e-6, 1e-4):
            for d in dirs:
                q = np.array(p) + eps * np.array(d)
                Us.extend(_dress(rng, O.canonical_gate(*q), 4))
    Us = np.stack(Us)
    c, g = _coords(Us, fold=True)
    _assert_matches_oracle_or_class(Us, c)
    assert np.abs(g - O.g1g2g3_raw(Us)).max() < TOL


def test_template_outputs_b1_b2():
    from helpers import make_pair

    desc, orc = make_pair("riswap", (0.5,), k=3)
    vals = {}
    p = 0
    for tri in KATS["B1"]["u3_triples"]:
        for v in tri:
            vals[f"P{p}"] = v
            p += 1
    x = torch.as_tensor(np.array([[vals[n] for n in orc.names_sorted]]), device="cuda")
    U = engine.template_eval(desc, x)
    c, _ = engine.weyl(U, round8=True)
    assert tuple(c[0].tolist()) == tuple(KATS["B1"]["c1c2c3"])


def test_ragged_sizes():
    rng = np.random.default_rng(2)
    for B in (1, 127, 128, 129, 1000):
        U = O.haar_unitary(rng, B)
        c, _ = _coords(U, fold=True)
        assert np.abs(c - O.fold_c1(O.c1c2c3_raw(U))).max() < TOL

"""Pin the CPU oracle against every recorded output the reference tree holds for this path
(SURVEY.md Appendix B -> tests/golden/kats.json).  CPU only."""
import json
import os

import numpy as np
import pytest
import scipy.linalg

import oracle as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))


def _b1_template_and_x(k=3, n_triples=8):
    t = O.OracleTemplate("riswap", (0.5,), k=k)
    vals = {}
    p = 0
    for tri in KATS["B1"]["u3_triples"][:n_triples]:
        for v in tri:
            vals[f"P{p}"] = v
            p += 1
    return t, [vals[n] for n in t.names_sorted]


def test_b1_decomp_trajectory_coords_and_square_cost():
    t, x = _b1_template_and_x()
    U = t.eval(x)
    assert O.c1c2c3(U) == tuple(KATS["B1"]["c1c2c3"])
    assert O.cost(U, O.SWAP, "square") == pytest.approx(KATS["B1"]["square_cost_vs_swap"], rel=1e-6)
    assert np.allclose(U.conj().T @ U, np.eye(4), atol=1e-14)


def test_b2_prefix_of_b1():
    t, x = _b1_template_and_x(k=2, n_triples=6)
    assert O.c1c2c3(t.eval(x)) == tuple(KATS["B2"]["c1c2c3"])


def test_b3_basis_gate_chamber_points_and_closed_form():
    for name, g in KATS["B3"]["gates"].items():
        G = O.conversion_gain(0, 0, g["gc"], g["gg"], g["t"])
        assert np.abs(G - O.conversion_gain_expm(0, 0, g["gc"], g["gg"], g["t"])).max() < 1e-15
        c = O.fold_c1(np.array(O.c1c2c3(G)))
        assert np.allclose(c, g["c1c2c3_folded"], atol=1e-8), name


def test_conversion_gain_closed_form_with_phases():
    rng = np.random.default_rng(5)
    for _ in range(20):
        pc, pg, gc, gg, t = rng.uniform(-3, 3, 5)
        assert np.abs(O.conversion_gain(pc, pg, gc, gg, t) - O.conversion_gain_expm(pc, pg, gc, gg, t)).max() < 5e-15
    # sign quirk: CG(0,0,pi/2,0,1/2) has -i/sqrt(2) off-diagonals, RiSwap(1/2) has +i/sqrt(2)
    assert np.isclose(O.conversion_gain(0, 0, np.pi / 2, 0, 0.5)[1, 2], -1j / np.sqrt(2))
    assert np.isclose(O.riswap(0.5)[1, 2], 1j / np.sqrt(2))


def test_b4_b5_standard_gates():
    assert O.c1c2c3(O.CNOT) == tuple(KATS["B4"]["CX"])
    assert O.c1c2c3(O.SWAP) == tuple(KATS["B4"]["SWAP"])
    assert O.c1c2c3(O.fsim(KATS["B5"]["fsim_theta"], KATS["B5"]["fsim_phi"])) == tuple(KATS["B5"]["c1c2c3"])
    assert O.c1c2c3(O.berkeley()) == (0.5, 0.25, 0.0)


def test_b5b_makhlin_functional_bit_pattern():
    a = np.array(KATS["B5b"]["g_target"])
    b = np.array(KATS["B5b"]["g_current"])
    assert float(np.sum(np.abs(a - b) ** 2)) == KATS["B5b"]["J_T_LI"]
    # and through the functional itself: a gate whose rounded invariants are b
    assert O.J_T_LI(O.SWAP, O.SWAP) == 0.0


def test_b11_smush_template_solution_recorded_by_the_reference():
    """The only reference-held record of the smush Hamiltonian path: the solved circuit drawn in
    scripts/local_smush_test.ipynb cell 5 (MakhlinFunctionalCost vs CX: 'Best Loss=0.0', 'Cost: 1.0')."""
    from helpers import b11_case

    slots, k, T, vals = b11_case(KATS)
    tmpl = O.OracleTemplate("smush", slots, k=k, T=T)
    U = tmpl.eval([vals[n] for n in tmpl.names_sorted])
    assert np.abs(U @ U.conj().T - np.eye(4)).max() < 1e-13
    assert O.J_T_LI(O.CNOT, U) == KATS["B11"]["makhlin_functional_vs_CX"] == 0.0
    assert O.cost(U, O.CNOT, "makhlin_functional") == 0.0
    c = O.c1c2c3(U)
    assert c[0] == KATS["B11"]["c1"] and 0 <= c[1] < KATS["B11"]["c2_c3_below"] and 0 <= c[2] < KATS["B11"]["c2_c3_below"]
    # circuit cost of the smush gate: (|gc| + |gg|) t / (pi/2)  (custom_gates.py:252-257)
    gc, gg, t = slots[2], slots[3], slots[-1]
    assert (abs(gc) + abs(gg)) * t / (np.pi / 2) == KATS["B11"]["circuit_cost"]


def test_b10_makhlin_invariants():
    gates = {"I": np.eye(4), "CNOT": O.CNOT, "SWAP": O.SWAP, "iSWAP": O.ISWAP, "sqrt_iSWAP": O.riswap(0.5)}
    for name, G in gates.items():
        assert O.g1g2g3(G) == tuple(KATS["B10"][name]), name


def test_b7_readme_target_is_reachable_at_k2():
    """README: Haar target with these coordinates is solved by 2 sqrt(iSWAP)s; use canonical_gate as surrogate."""
    c = KATS["B7"]["target_c1c2c3"]
    V = O.canonical_gate(*c)
    assert np.allclose(O.c1c2c3(V), c, atol=1e-8)
    res = O.literal_run(lambda k: O.OracleTemplate("riswap", (0.5,), k=k), V, range(2, 3), restarts=5,
                        rng=np.random.default_rng(3))
    assert res.best_result < 1e-8 and res.best_cycles == 2


def test_philox_random123_vectors():
    for v in KATS["philox4x32_10"]["vectors"]:
        out = O.philox4x32_10(*[int(c, 16) for c in v["ctr"]], *[int(k, 16) for k in v["key"]])
        assert [int(o) for o in out] == [int(o, 16) for o in v["out"]]


def test_magic_basis_spectrum_equals_childs_spectrum():
    rng = np.random.default_rng(0)
    for _ in range(10):
        U = O.haar_unitary(rng)
        UB = O.slam_oracle._QMAGIC.conj().T @ U @ O.slam_oracle._QMAGIC
        m = UB.T @ UB
        ev1 = np.sort_complex(np.linalg.eigvals(m))
        ev2 = np.sort_complex(np.linalg.eigvals(U @ (O.slam_oracle._SYSY @ U.T @ O.slam_oracle._SYSY)))
        assert np.allclose(ev1, ev2, atol=1e-12)


def test_weyl_coords_in_chamber_and_local_invariance():
    rng = np.random.default_rng(1)
    U = O.haar_unitary(rng, 200)
    c = O.c1c2c3_raw(U)
    assert np.all(c[:, 0] >= -1e-12) and np.all(c[:, 0] < 1 + 1e-12)
    assert np.all(c[:, 1] <= 0.5 + 1e-12) and np.all(c[:, 2] <= c[:, 1] + 1e-12) and np.all(c[:, 2] >= -1e-12)
    # invariance under local unitaries
    for i in range(20):
        k1 = np.kron(O.u3(*rng.uniform(0, 6, 3)), O.u3(*rng.uniform(0, 6, 3)))
        k2 = np.kron(O.u3(*rng.uniform(0, 6, 3)), O.u3(*rng.uniform(0, 6, 3)))
        assert np.allclose(O.c1c2c3_raw(k1 @ U[i] @ k2), c[i], atol=1e-10)
        assert np.allclose(O.g1g2g3_raw(k1 @ U[i] @ k2), O.g1g2g3_raw(U[i]), atol=1e-10)


def test_param_order_is_lexicographic():
    t = O.OracleTemplate("riswap", (0.5,), k=2)
    assert t.names_sorted[:4] == ["P0", "P1", "P10", "P11"]
    t2 = O.OracleTemplate("smush", ("Q", "Q", 1.0, 0.5, "Q", "Q", "Q", "Q", 0.5), k=2, T=2, no_exterior_1q=True)
    assert t2.n_params == 6 * (2 - 1) + 2 * (2 + 2 * 2)
    assert t2.names_sorted[-1].startswith("Q") and t2.names_sorted[0] == "P0"


def test_eval_batch_matches_literal_eval():
    rng = np.random.default_rng(2)
    for kind, slots in (("riswap", (0.5,)), ("cg", (0.3, -0.2, np.pi / 4, np.pi / 4, 0.5)), ("riswap", ("Q",)),
                        ("cg", ("Q", "Q", "Q", "Q", "Q"))):
        t = O.OracleTemplate(kind, slots, k=3)
        X = rng.uniform(-7, 7, (5, t.n_params))
        Ub = t.eval_batch(X)
        for i in range(5):
            assert np.abs(Ub[i] - t.eval(X[i])).max() < 1e-14


def test_analytic_gradient_matches_finite_differences():
    rng = np.random.default_rng(3)
    V = O.haar_unitary(rng)
    for kind, slots, kw in (("riswap", (0.5,), {}), ("cg", (0.0, 0.0, np.pi / 4, np.pi / 4, 0.5), {}),
                            ("cg", ("Q", "Q", "Q", "Q", "Q"), {}), ("riswap", ("Q",), {"vz_only": True}),
                            ("riswap", (0.5,), {"no_exterior_1q": True})):
        t = O.OracleTemplate(kind, slots, k=3, **kw)
        x = rng.uniform(0, 2 * np.pi, t.n_params)
        for cost_kind in ("basic", "square", "basic_inverse"):
            loss, g, _ = O.loss_and_grad(t, x, V, cost_kind)
            f = O.literal_objective(t, V, cost_kind)
            assert abs(loss - f(x)) < 1e-14
            h = 1e-6
            gfd = np.array([(f(x + h * e) - f(x - h * e)) / (2 * h) for e in np.eye(t.n_params)])
            assert np.abs(g - gfd).max() < 1e-8, (kind, cost_kind)


def test_smush_matches_plain_cg_when_drives_off():
    U = O.smush(0.3, -0.4, 1.1, 0.2, [0.0, 0.0], [0.0, 0.0], 0.7)
    assert np.abs(U - O.conversion_gain(0.3, -0.4, 1.1, 0.2, 0.7)).max() < 1e-14
    # general slice is unitary
    U = O.smush_1qphase(0.1, 0.2, 0.3, 0.4, 1.0, 0.5, 0.3, -0.2, [1.0, -2.0, 0.5], [0.3, 0.1, -1.0], 1.5)
    assert np.allclose(U.conj().T @ U, np.eye(4), atol=1e-13)


def test_haar_sample_reseeds_every_call():
    a = O.haar_sample_unitary(7)
    b = O.haar_sample_unitary(7)
    assert np.array_equal(a, b)  # sampler.py:67-71 quirk
    assert np.allclose(a.conj().T @ a, np.eye(4), atol=1e-13)


def test_coverage_stream_is_reproducible_and_shardable():
    t = O.OracleTemplate("cg", (0.0, 0.0, np.pi / 2, 0.0, 0.5), k=2, no_exterior_1q=True)
    h_all = O.coverage_histogram(t, 2023, 0, 400, 0.0, 2 * np.pi, nbins=16)
    h_a = O.coverage_histogram(t, 2023, 0, 150, 0.0, 2 * np.pi, nbins=16)
    h_b = O.coverage_histogram(t, 2023, 150, 250, 0.0, 2 * np.pi, nbins=16)
    assert h_all.sum() == 400 and np.array_equal(h_all, h_a + h_b)
    u = O.philox_uniform(2023, np.arange(4, dtype=np.uint64), 7)
    assert u.shape == (4, 7) and np.all(u >= 0) and np.all(u < 1)


def test_flop_model():
    assert [O.F_eval(k) for k in (1, 3, 6)] == [1400, 3696, 7140]
    assert [O.F_lossgrad(k) for k in (1, 3, 6)] == [7032, 15984, 29412]

"""Host-side logic that needs no GPU: symbolic circuit, parameter ordering, descriptor lowering, samplers."""
import numpy as np
import pytest

import oracle as O
from helpers import BASES, make_pair
from slam_decomposition_b200 import _lib
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.basisv2 import CircuitTemplateV2
from slam_decomposition_b200.circuit import Parameter, TemplateCircuit, lower
from slam_decomposition_b200.distributed import shard_range
from slam_decomposition_b200.sampler import GateSample, HaarBatchSample, HaarSample
from slam_decomposition_b200.utils.gates.custom_gates import (ConversionGainGate, ConversionGainSmushGate, CXGate, RiSwapGate,
                                                             SwapGate)


def test_circuit_template_structure_and_param_order():
    b = CircuitTemplate(maximum_span_guess=4, preseed=False)
    assert list(b.spanning_range) == [1, 2, 3, 4]
    for k in (1, 2, 3, 4):
        b.build(k)
        orc = O.OracleTemplate("riswap", (0.5,), k=k)
        assert [p.name for p in b.circuit.parameters] == orc.names_sorted
        assert b.desc.n_params == 6 * (k + 1) and b.desc.k == k and b.desc.gate_kind == _lib.GATE_RISWAP
        assert len(b.parameter_guess()) == 6 * (k + 1)
    # lexicographic order: P0, P1, P10, P11, ...
    b.build(2)
    assert [p.name for p in b.circuit.parameters][:4] == ["P0", "P1", "P10", "P11"]
    # layer 0 qubit 0 holds P0..P2 -> API indices of those names
    names = [p.name for p in b.circuit.parameters]
    assert [b.desc.p1q[0][j] for j in range(6)] == [names.index(f"P{j}") for j in range(6)]
    with pytest.raises(ValueError):
        b.build(0)


def test_no_exterior_and_vz_only_lowering():
    b = CircuitTemplateV2(base_gates=[RiSwapGate], no_exterior_1q=True, vz_only=True)
    b.build(3)
    d = b.desc
    assert d.vz_only == 1 and d.p1q[0][0] == -1 and d.p1q[3][0] == -1 and d.p1q[1][0] >= 0 and d.p1q[1][1] == -1
    assert d.n_params == 2 * 2 + 3  # two interior RZ layers + three Q parameters
    assert all(d.slot_param[g][0] >= 0 for g in range(3))


def test_v2_smush_template_parameter_count_matches_reference_formula():
    """P(k,T) = 6(k-1) + k(2+2T) for the coverage template (parallel_drive_volume.py:175-198)."""
    for base, T in (("sqiSwap", 2), ("iSwap", 4), ("sqCNOT", 2)):
        gc, gg, t = BASES[base]

        def pp2(*vargs, gc=gc, gg=gg, t=t, T=T):
            return ConversionGainSmushGate(vargs[0], vargs[1], gc, gg, vargs[2:2 + T], vargs[2 + T:], t_el=t)

        for k in (1, 3, 6):
            b = CircuitTemplateV2(n_qubits=2, base_gates=[pp2], no_exterior_1q=1, vz_only=0, param_vec_expand=[2, T, T])
            b.build(k)
            assert b.desc.n_params == 6 * (k - 1) + k * (2 + 2 * T)
            assert b.desc.gate_kind == _lib.GATE_SMUSH and b.desc.T == T and b.desc.n_slots == 5 + 2 * T
            x = b.parameter_guess()
            assert len(x) == b.desc.n_params and all(-4 * np.pi <= v <= 4 * np.pi for v in x)
            assert b.bounds_list is None


def test_assign_parameters_and_bound_circuit():
    b = CircuitTemplate(preseed=False)
    b.build(1)
    x = np.arange(12, dtype=float)
    qc = b.assign_Xk(x)
    assert qc.num_parameters == 0
    names = [p.name for p in b.circuit.parameters]
    first_u = qc[0].operation
    assert first_u.params == [x[names.index("P0")], x[names.index("P1")], x[names.index("P2")]]
    desc, names2, numeric = lower(qc)
    assert names2 == [] and numeric.size == 12 and desc.n_params == 12


def test_lower_rejects_unsupported_structures():
    qc = TemplateCircuit(2)
    qc.u(Parameter("P0"), Parameter("P1"), Parameter("P2"), 0)
    with pytest.raises(ValueError):
        lower(qc)  # no 2Q gate
    qc.append(RiSwapGate(0.5), (0, 1))
    qc.append(ConversionGainGate(0, 0, 1, 1, 1), (0, 1))
    with pytest.raises(NotImplementedError):
        lower(qc)  # mixed gate families
    qc3 = TemplateCircuit(3)
    with pytest.raises(NotImplementedError):
        lower(qc3)
    with pytest.raises(NotImplementedError):
        CircuitTemplate(n_qubits=3)
    with pytest.raises(NotImplementedError):
        CircuitTemplate(use_polytopes=True)


def test_helpers_pair_is_consistent():
    desc, orc = make_pair("cg", ("Q", 0.2, np.pi / 4, "Q", 0.5), k=3)
    assert desc.n_params == orc.n_params == 24 + 6
    assert desc.slot_param[0][0] >= 0 and desc.slot_param[0][1] == -1 and desc.slot_const[0][1] == 0.2


def test_samplers():
    a, b2 = list(HaarSample(seed=5, n_samples=2))
    assert np.array_equal(a, b2)  # re-seeded on every call (sampler.py:67-71)
    assert np.array_equal(a, O.haar_sample_unitary(5))
    hs = list(HaarBatchSample(seed=1, n_samples=3))
    assert not np.allclose(hs[0], hs[1])
    assert np.allclose(list(GateSample(CXGate()))[0], O.CNOT)
    assert np.allclose(list(GateSample(SwapGate(), n_samples=2))[1], O.SWAP)
    assert GateSample(SwapGate()).n_qubits == 2


def test_gate_costs_follow_the_reference():
    assert RiSwapGate(0.5).cost() == 0.5
    g = ConversionGainGate(0, 0, np.pi / 4, np.pi / 4, 1.0)
    assert g.cost() == pytest.approx(1.0)
    assert str(g).startswith("2QGate(0.78539816")
    s = ConversionGainSmushGate(0, 0, np.pi / 2, 0, [1, 2], [3, 4], 0.5)
    assert s.xy_len == 2 and s.cost() == pytest.approx(0.5) and len(s.params) == 9


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 100000, 10 ** 9):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def test_circuit_cost_batch_matches_the_per_sample_shim():
    """The device optimiser's batched circuit_cost (descriptor slot tables, torch) against CircuitTemplateV2.circuit_cost
    (basisv2.py:98-127) for RiSwap (alpha) and ConversionGain ((|gc| + |gg|) t / (pi/2)) templates."""
    import torch

    from slam_decomposition_b200.optimizer import TemplateOptimizer

    rng = np.random.default_rng(0)
    b = CircuitTemplateV2(base_gates=[RiSwapGate])
    b.build(3)
    X = rng.uniform(-1, 1, (5, b.desc.n_params))
    got = TemplateOptimizer._circuit_cost_batch(b.desc, torch.as_tensor(X)).numpy()
    assert np.allclose(got, [b.circuit_cost(x) for x in X], atol=1e-14)

    def cg(*v):
        return ConversionGainGate(v[0], v[1], v[2], v[3], 0.5)

    b2 = CircuitTemplateV2(base_gates=[cg], param_vec_expand=[4])
    b2.build(2)
    X = rng.uniform(-2, 2, (5, b2.desc.n_params))
    got = TemplateOptimizer._circuit_cost_batch(b2.desc, torch.as_tensor(X)).numpy()
    assert np.allclose(got, [b2.circuit_cost(x) for x in X], atol=1e-13)
    b2.set_constraint(1.5)
    assert b2.using_constraints and b2.constraint_max == 1.5
    assert np.isclose(b2.constraint_func["fun"](X[0]), 1.5 - b2.circuit_cost(X[0]))
    b2.remove_constraint()
    assert not b2.using_constraints


def test_hull_coverage_reproduces_the_reference_base_volume_of_two_sqrt_iswaps():
    """Post-processing of a coverage histogram (parallel_drive_volume.py:343-396), on the host: a cloud that fills the
    known span of two sqrt(iSWAP) gates, {c1 >= c2 + c3} inside the half-chamber, must give the Haar fraction the
    reference recorded as its base volume (data/extended_results.json: sqiSwap k=2 -> 0.7901173636843226), contain CNOT
    and B but not SWAP -- and a flat cloud (the c3 = 0 plane of two CNOTs) has zero volume."""
    import json
    import os

    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

    ref = 0.7901173636843226
    pts, w = pdv.half_chamber_grid(256)
    sel = pts[:, 0] >= pts[:, 1] + pts[:, 2]
    assert abs((w * sel).sum() / w.sum() - ref) < 2e-4  # the Haar density and the grid quadrature themselves
    # (a) exact sample points: a dense cloud in the span (the hull of samples is an inner approximation)
    rng = np.random.default_rng(0)
    cloud = pts[sel][rng.choice(int(sel.sum()), 200_000, replace=False)]
    out = pdv.hull_coverage(cloud, plain_cloud=cloud, grid=128)
    assert abs(out["extended_vol"] - ref) < 5e-3 and abs(out["base_vol"] - ref) < 5e-3
    assert (out["has_CNOT"], out["has_SWAP"], out["has_B"]) == (True, False, True)
    # (b) a histogram of the same cloud: voxel approximation, good to a voxel's worth of volume
    nb = 128
    hist = np.bincount(np.ravel_multi_index(np.floor(cloud * 2 * nb).astype(int).T, (nb, nb, nb)), minlength=nb ** 3)
    out = pdv.hull_coverage(hist, nbins=nb, grid=128)
    assert abs(out["extended_vol"] - ref) < 0.02 and out["base_vol"] == 0.0
    # (c) a flat cloud (the c3 = 0 plane that two CNOTs span) has zero volume and contains nothing by itself
    flat = cloud.copy()
    flat[:, 2] = 0.0
    out = pdv.hull_coverage(flat, grid=64)
    assert out["extended_vol"] == 0.0 and out["base_vol"] == 0.0
    assert (out["has_CNOT"], out["has_SWAP"], out["has_B"]) == (False, False, False)
    # with exact base flags (base_reachable) the base set decides
    out = pdv.hull_coverage(flat, grid=64, base_flags={"CNOT": True, "SWAP": False, "B": True})
    assert (out["has_CNOT"], out["has_SWAP"], out["has_B"]) == (True, False, True)
    # the reference's recorded table is shipped as a golden fixture for the GPU-side comparison
    tab = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "extended_results.json")))
    assert tab["sqiSwap"]["2"][0] == ref


def test_descriptor_cache_key_follows_the_gate_cycle_phase():
    """TemplateOptimizer caches the template descriptors of a repeated chained sweep under (basis, sizes, gate-cycle phase).
    The reference's gate cycles keep advancing across builds (basis.py:68-72): with one base gate the phase never changes
    (cache hit on the next sweep), with two it depends on how many gates earlier builds consumed -- and a build from a
    different phase really yields a different descriptor."""
    from slam_decomposition_b200.optimizer import TemplateOptimizer

    one = CircuitTemplate(base_gates=[RiSwapGate(1 / 2)], maximum_span_guess=3, preseed=False)
    p0 = TemplateOptimizer._cycle_phase(one)
    one.build(3)
    d1 = bytes(one.desc)
    assert TemplateOptimizer._cycle_phase(one) == p0
    one.build(3)
    assert bytes(one.desc) == d1

    two = CircuitTemplate(base_gates=[RiSwapGate(1 / 2), RiSwapGate(1 / 3)], maximum_span_guess=3, preseed=False)
    q0 = TemplateOptimizer._cycle_phase(two)
    two.build(3)  # consumes three gates of a cycle of two: the phase moves
    e1 = bytes(two.desc)
    assert TemplateOptimizer._cycle_phase(two) != q0
    two.build(3)
    assert bytes(two.desc) != e1

"""K6 parity: fused Philox -> template -> Weyl -> fold -> bin Monte-Carlo vs the oracle on the same stream."""
import numpy as np
import pytest
import torch

import oracle as O
from helpers import BASES, make_pair
from slam_decomposition_b200 import engine

pytestmark = pytest.mark.gpu


def _check(desc, orc, seed, first, n, lo, hi, nbins):
    hist, coords = engine.coverage_mc(desc, seed, first, n, lo, hi, nbins=nbins, want_coords=True)
    hist, coords = hist.cpu().numpy(), coords.cpu().numpy()
    params = O.coverage_params(seed, first, n, orc.n_params, lo, hi)
    ref = O.coverage_points(orc, params)
    assert np.abs(coords - ref).max() < 1e-10  # identical sample stream -> identical points
    edge = O.near_bin_edge(ref, nbins, tol=1e-9)
    got_bins = O.bin_index(coords, nbins)
    ref_bins = O.bin_index(ref, nbins)
    assert np.array_equal(got_bins[~edge], ref_bins[~edge])  # bit-exact membership away from bin edges
    assert hist.sum() == n
    ref_hist = np.bincount(ref_bins, minlength=nbins ** 3)
    assert np.abs(hist - ref_hist).sum() <= 2 * edge.sum()
    return edge.mean()


@pytest.mark.parametrize("base", ["sqiSwap", "CNOT", "B"])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_plain_template_coverage(base, k):
    """Variant A: fixed basis gate, interior U3 parameters ~ U[0, 2pi)."""
    gc, gg, t = BASES[base]
    if k == 1:
        desc, orc = make_pair("cg", (0.0, 0.0, gc, gg, t), k=1)  # exterior layers only: a single chamber point
    else:
        desc, orc = make_pair("cg", (0.0, 0.0, gc, gg, t), k=k, no_exterior_1q=True)
    # (CNOT k=2 lives entirely on the c3 = 0 face, i.e. on a bin edge: membership there is checked through the
    # |hist - ref| <= 2 * n_edge bound inside _check)
    _check(desc, orc, 2023, 0, 3000, 0.0, 2 * np.pi, 128)


@pytest.mark.parametrize("base,T", [("sqiSwap", 2), ("CNOT", 4)])
def test_smush_template_coverage(base, T):
    """Variant B: the reference's smush template, every parameter ~ U(-4pi, 4pi) (parallel_drive_volume.py:175-222)."""
    gc, gg, t = BASES[base]
    slots = ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,)
    desc, orc = make_pair("smush", slots, k=2, T=T, no_exterior_1q=True)
    _check(desc, orc, 7, 1000, 400, -4 * np.pi, 4 * np.pi, 64)


def test_shards_add_up_and_stream_is_position_independent():
    gc, gg, t = BASES["sqiSwap"]
    desc, orc = make_pair("cg", (0.0, 0.0, gc, gg, t), k=2, no_exterior_1q=True)
    n = 100000
    whole, _ = engine.coverage_mc(desc, 11, 0, n, 0.0, 2 * np.pi, nbins=32)
    parts = torch.zeros_like(whole)
    for lo, hi in ((0, 12345), (12345, 70000), (70000, n)):
        engine.coverage_mc(desc, 11, lo, hi - lo, 0.0, 2 * np.pi, nbins=32, hist=parts)
    assert torch.equal(whole, parts)
    assert whole.sum().item() == n
    # sqrt(iSWAP) k=2 cannot leave the c3 = 0 ... actually covers a 3-D region; sanity: all mass inside the chamber
    idx = torch.nonzero(whole).flatten()
    k3 = idx % 32
    k2 = (idx // 32) % 32
    assert (k3 <= k2).all()  # c3 <= c2 inside the Weyl chamber


def test_coverage_fraction_sanity_against_extended_results():
    """B8 (loose): plain sqrt(iSWAP) k=2 covers ~79% of the Haar volume; the Haar-weighted voxel estimate from
    the MC cloud must be in that neighbourhood (the reference value is an N=3000 convex hull)."""
    gc, gg, t = BASES["sqiSwap"]
    desc, _ = make_pair("cg", (0.0, 0.0, gc, gg, t), k=2, no_exterior_1q=True)
    nb = 32
    hist, _ = engine.coverage_mc(desc, 5, 0, 2_000_000, 0.0, 2 * np.pi, nbins=nb)
    occ = (hist.reshape(nb, nb, nb) > 0).cpu().numpy()
    # Haar density on folded Weyl coordinates (units of pi): |prod_{i<j} sin(pi(ci+cj)) sin(pi(ci-cj))| restricted to the chamber
    g = (np.arange(nb) + 0.5) / (2 * nb)
    c1, c2, c3 = np.meshgrid(g, g, g, indexing="ij")
    inside = (c2 <= c1) & (c3 <= c2)
    dens = np.abs(np.sin(np.pi * (c1 + c2)) * np.sin(np.pi * (c1 - c2)) * np.sin(np.pi * (c1 + c3)) * np.sin(np.pi * (c1 - c3))
                  * np.sin(np.pi * (c2 + c3)) * np.sin(np.pi * (c2 - c3)))
    frac = (dens * inside * occ).sum() / (dens * inside).sum()
    assert 0.70 < frac < 0.90, frac


def test_reference_style_coverage_api():
    """The user-facing helpers that mirror parallel_drive_volume.py: template construction, cloud, fold, volume."""
    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

    gc, gg, t, name, iters = pdv.GATE_LIST[1]  # sqiSwap
    basis = pdv.smush_template(gc, gg, t, k=2)
    T = round(t / pdv.duration_1q)
    assert basis.desc.n_params == 6 * (2 - 1) + 2 * (2 + 2 * T)
    pts = pdv.coverage_points(basis, n_samples=pdv.N, seed=1).cpu().numpy()
    slots = ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,)
    _, orc = make_pair("smush", slots, k=2, T=T, no_exterior_1q=True)
    ref = O.coverage_points(orc, O.coverage_params(1, 0, pdv.N, orc.n_params, -4 * np.pi, 4 * np.pi))
    assert np.abs(pts - ref).max() < 1e-10
    left, right = pdv.mirror_fold([(0.7, 0.2, 0.1), (0.3, 0.2, 0.1)])
    assert np.allclose(left, [[0.3, 0.2, 0.1], [0.3, 0.2, 0.1]]) and np.allclose(right, [[0.7, 0.2, 0.1], [0.7, 0.2, 0.1]])
    # parallel drive extends the k=2 coverage of sqrt(iSWAP) beyond the plain template's (extended_results.json: 0.79 -> 0.83)
    nb = 32
    h_plain = pdv.coverage_histogram(pdv.plain_template(gc, gg, t, 2), 1_000_000, seed=3, nbins=nb)
    h_smush = pdv.coverage_sweep(basis, 1_000_000, seed=3, nbins=nb)
    v_plain, v_smush = pdv.haar_volume_fraction(h_plain, nb), pdv.haar_volume_fraction(h_smush, nb)
    assert h_smush.sum().item() == 1_000_000
    # (a finite uniform(-4pi,4pi) cloud need not fill its reachable set as densely as the plain template's, so only
    # the plain-template figure is compared with the reference's 0.79)
    assert 0.70 < v_plain < 0.90 and 0.5 < v_smush <= 1.0


def test_hull_volumes_and_flags_reproduce_the_reference_table():
    """Post-processing parity: the reference's study protocol -- N = 3000 random parallel-drive instances per
    (gate, k), convex hull per mirror side, union with the base set, Haar volume, CNOT / SWAP / B membership
    (parallel_drive_volume.py:140-410) -- against its recorded table src/slam/data/extended_results.json (golden fixture).
    The reference's numbers are single draws of a stochastic estimate (and its hulls are far from converged at N = 3000: at
    N = 2e6 the same hulls reach 0.79 / 0.92 / 0.86 / 0.44 for iSwap k=1 / sqiSwap k=2 / CNOT k=2 / B k=1), so the mean over
    seeds is compared with the tolerance of that scatter; base volumes (exact polytopes in the reference) are tighter."""
    import json
    import os

    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "extended_results.json")))
    gates = {name: (gc, gg, t) for gc, gg, t, name, _ in pdv.GATE_LIST}
    #        gate      k  tol(extended)  tol(base)
    rows = [("sqiSwap", 2, 0.02, 0.005), ("CNOT", 2, 0.05, 0.002), ("sqCNOT", 4, 0.02, 0.02), ("sqCNOT", 5, 0.006, 0.006),
            ("B", 1, 0.03, 0.002), ("iSwap", 1, 0.07, 0.002), ("sqCNOT", 3, 0.01, 0.01), ("sqB", 2, 0.01, 0.01)]
    for name, k, tol_e, tol_b in rows:
        r = ref[name][str(k)]
        out = [pdv.coverage_study(*gates[name], k, seed=100 + 7 * s, n_base=400_000, exact_flags=(s == 0)) for s in range(6)]
        ext = float(np.mean([o[1] for o in out]))
        assert abs(out[0][0] - r[0]) <= tol_b, (name, k, "base", out[0][0], r[0])
        assert abs(ext - r[1]) <= tol_e, (name, k, "extended", ext, r[1])
        assert [bool(v) for v in out[0][2:]] == [bool(v) for v in r[2:]], (name, k, out[0][2:], r[2:])
    assert ref["sqCNOT"]["5"][1] >= 0.995 and np.mean([o[1] for o in out]) > 0  # (sanity of the fixture itself)
    # membership flags of every (gate, k) row of the table that is not the forced full-coverage row
    for gc, gg, t, name, iters in pdv.GATE_LIST:
        for k in range(1, iters):
            flags = pdv.coverage_study(gc, gg, t, k, seed=3, n_base=50_000)[2:]
            assert [bool(v) for v in flags] == [bool(v) for v in ref[name][str(k)][2:]], (name, k, flags)

"""BASELINE.json-sized runs checked through size-independent properties (the oracle cannot follow at these sizes)."""
import numpy as np
import pytest
import torch

import oracle as O
from helpers import BASES, make_pair
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate

pytestmark = pytest.mark.gpu


def _haar(n, seed):
    rng = np.random.default_rng(seed)
    z = (rng.standard_normal((n, 4, 4)) + 1j * rng.standard_normal((n, 4, 4))) / np.sqrt(2.0)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r, axis1=-2, axis2=-1)
    return np.ascontiguousarray(q * (d / np.abs(d))[:, None, :])


@pytest.mark.parametrize("name", ["sqiSwap", "CNOT", "B"])
def test_coverage_1e7_samples_configs1(name):
    """configs[1]: 1e7 random samples per (basis, k): counts are conserved, shards add up exactly (histogram is a
    checksum of the whole stream), every occupied bin lies inside the Weyl chamber, and a 3000-sample prefix is
    bit-identical to the oracle's bins away from bin edges."""
    gc, gg, t = BASES[name]
    n = 10_000_000
    for k in (2, 3):
        basis = pdv.plain_template(gc, gg, t, k)
        whole = pdv.coverage_histogram(basis, n, seed=2023)
        assert whole.sum().item() == n
        parts = torch.zeros_like(whole)
        for lo, hi in ((0, 3_333_333), (3_333_333, 9_000_001), (9_000_001, n)):
            pdv.coverage_histogram(basis, hi - lo, seed=2023, first_sample=lo, hist=parts)
        assert torch.equal(whole, parts)
        idx = torch.nonzero(whole).flatten()
        b3, b2 = idx % 128, (idx // 128) % 128
        assert (b3 <= b2).all()  # c3 <= c2 everywhere in the chamber (c2 <= c1 only holds on the un-folded half)
        # prefix vs oracle
        _, orc = make_pair("cg", (0.0, 0.0, gc, gg, t), k=k, no_exterior_1q=True)
        ref = O.coverage_points(orc, O.coverage_params(2023, 0, 3000, orc.n_params, 0.0, 2 * np.pi))
        h3, _ = engine.coverage_mc(basis.desc, 2023, 0, 3000, 0.0, 2 * np.pi, nbins=128)
        edge = O.near_bin_edge(ref, 128, tol=1e-9)
        refh = np.bincount(O.bin_index(ref, 128), minlength=128 ** 3)
        assert np.abs(h3.cpu().numpy() - refh).sum() <= 2 * edge.sum()


def test_haar_sweep_1e5_targets_configs2():
    """configs[2]: 1e5 Haar targets x 16 restarts onto sqCNOT templates k <= 6 with early exit.  Checked through:
    every target solved below 1e-10, the per-k fractions (a property of the basis gate), a random sample of the
    returned parameter vectors re-evaluated by the oracle, and run-to-run reproducibility of the statistics."""
    Nt = 100_000
    V = _haar(Nt, 42)
    gate = ConversionGainGate(0.0, 0.0, *BASES["sqCNOT"])
    opt = TemplateOptimizer(CircuitTemplate(base_gates=[gate], maximum_span_guess=6, preseed=False), BasicCost(),
                            override_fail=True, training_restarts=16)
    np.random.seed(0)
    out = opt.approximate_targets(V)
    assert out["success"].all() and out["loss"].max() <= 1e-10
    frac = np.bincount(out["cycles"], minlength=7)[1:] / Nt
    assert frac[0] == 0 and frac[1] < 1e-3 and 0.47 < frac[2] < 0.53 and 0.44 < frac[3] < 0.49 and frac[4] < 0.06
    rng = np.random.default_rng(1)
    for i in rng.integers(0, Nt, 12):
        k = int(out["cycles"][i])
        tmpl = O.OracleTemplate("cg", (0.0, 0.0, *BASES["sqCNOT"]), k=k)
        assert out["n_params"][i] == tmpl.n_params
        assert O.cost(tmpl.eval(out["Xk"][i, : tmpl.n_params]), V[i], "basic") <= 1e-9
    # the loss+grad kernel agrees with the optimiser's reported losses on all 1e5 solutions of one k at once
    k = 3
    sel = np.nonzero(out["cycles"] == k)[0]
    desc, _ = make_pair("cg", (0.0, 0.0, *BASES["sqCNOT"]), k=k)
    X = torch.as_tensor(np.ascontiguousarray(out["Xk"][sel, : desc.n_params]), device="cuda")
    Vd = torch.as_tensor(V[sel], device="cuda")
    loss, grad, _ = engine.loss_grad(desc, X, Vd, tgt_idx=torch.arange(len(sel), dtype=torch.int32, device="cuda"))
    assert np.abs(loss.cpu().numpy() - out["loss"][sel]).max() < 1e-12
    assert grad.abs().max().item() < 1e-4  # stationary points
    np.random.seed(1)
    out2 = opt.approximate_targets(V)
    frac2 = np.bincount(out2["cycles"], minlength=7)[1:] / Nt
    assert np.abs(frac - frac2).max() < 5e-3


def test_streaming_loss_grad_2e22_rows_linearity_in_the_target():
    """Eval micro-benchmark size (millions of rows): T = Tr(V^dag U) is linear in V, so for V3 = a V1 + b V2
    the traces must satisfy T3 = conj(a) T1 + conj(b) T2 on every row."""
    desc, orc = make_pair("riswap", (0.5,), k=3)
    B = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.rand((B, orc.n_params), device="cuda", dtype=torch.float64, generator=g) * (2 * np.pi)
    V12 = torch.as_tensor(_haar(2, 3), device="cuda")
    a, b = 0.3 - 0.4j, -0.7 + 0.2j
    V3 = (a * V12[0] + b * V12[1])[None]
    _, _, T1 = engine.loss_grad(desc, X, V12[0:1].contiguous(), want_grad=False, want_trace=True)
    _, _, T2 = engine.loss_grad(desc, X, V12[1:2].contiguous(), want_grad=False, want_trace=True)
    _, _, T3 = engine.loss_grad(desc, X, V3.contiguous(), want_grad=False, want_trace=True)
    assert (T3 - (np.conj(a) * T1 + np.conj(b) * T2)).abs().max().item() < 1e-12
    assert T1.abs().max().item() <= 4.0 + 1e-12

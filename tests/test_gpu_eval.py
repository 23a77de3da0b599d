"""K1/K2 parity: CUDA template evaluation, loss and analytic gradient vs the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from helpers import BASES, make_pair
from slam_decomposition_b200 import engine

pytestmark = pytest.mark.gpu
TOL_U = 1e-10  # north_star: template unitaries within 1e-10 absolute
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))

CASES = [
    ("riswap", (0.5,), {}),
    ("riswap", (0.5,), {"no_exterior_1q": True}),
    ("cg", (0.0, 0.0, np.pi / 4, np.pi / 4, 0.5), {}),          # sqCNOT (GM_SYM)
    ("cg", (0.3, -0.7, 3 * np.pi / 8, np.pi / 8, 1.0), {}),     # B with phases (GM_BLOCK const)
    ("riswap", ("Q",), {}),                                    # continuous 2Q search, V2 default
    ("cg", ("Q", "Q", "Q", "Q", "Q"), {}),
    ("cg", ("Q", 0.2, np.pi / 4, "Q", 0.5), {}),
    ("riswap", (0.5,), {"vz_only": True}),
    ("fixed", (), {"fixed": O.berkeley()}),
]


def _dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=dtype)


@pytest.mark.parametrize("kind,slots,kw", CASES)
@pytest.mark.parametrize("k", [1, 2, 3, 6])
def test_template_eval_matches_oracle(kind, slots, kw, k):
    if kw.get("no_exterior_1q") and k == 1:
        pytest.skip("no parameters")
    desc, orc = make_pair(kind, slots, k=k, **kw)
    rng = np.random.default_rng(100 + k)
    X = rng.uniform(-4 * np.pi, 4 * np.pi, (257, orc.n_params))
    U = engine.template_eval(desc, _dev(X)).cpu().numpy()
    Uo = orc.eval_batch(X)
    assert np.abs(U - Uo).max() < TOL_U
    assert np.abs(U - Uo).max() < 1e-13  # in practice a few ulp


def test_b1_golden_vector_on_device():
    desc, orc = make_pair("riswap", (0.5,), k=3)
    vals = {}
    p = 0
    for tri in KATS["B1"]["u3_triples"]:
        for v in tri:
            vals[f"P{p}"] = v
            p += 1
    x = np.array([[vals[n] for n in orc.names_sorted]])
    U = engine.template_eval(desc, _dev(x)).cpu().numpy()[0]
    assert O.c1c2c3(U) == tuple(KATS["B1"]["c1c2c3"])
    swap = _dev(O.SWAP[None])
    loss, _, _ = engine.loss_grad(desc, _dev(x), swap, cost_kind=1, want_grad=False)
    assert loss.item() == pytest.approx(KATS["B1"]["square_cost_vs_swap"], rel=1e-6)


@pytest.mark.parametrize("kind,slots,kw", CASES)
@pytest.mark.parametrize("k", [1, 3, 6])
@pytest.mark.parametrize("cost_kind", [0, 1, 2])
def test_loss_and_gradient_match_oracle(kind, slots, kw, k, cost_kind):
    if kw.get("no_exterior_1q") and k == 1:
        pytest.skip("no parameters")
    desc, orc = make_pair(kind, slots, k=k, **kw)
    rng = np.random.default_rng(7 * k + cost_kind)
    B, Nt = 70, 9
    X = rng.uniform(0, 2 * np.pi, (B, orc.n_params))
    V = O.haar_unitary(rng, Nt)
    tgt = rng.integers(0, Nt, B).astype(np.int32)
    name = ("basic", "square", "basic_inverse")[cost_kind]
    ref = [O.loss_and_grad(orc, X[b], V[tgt[b]], name) for b in range(B)]
    for lpp in (4, 2, 1, 0):  # every team width, and the automatic choice
        loss, grad, trace = engine.loss_grad(desc, _dev(X), _dev(V), _dev(tgt), cost_kind=cost_kind, want_trace=True, lanes=lpp)
        loss, grad, trace = loss.cpu().numpy(), grad.cpu().numpy(), trace.cpu().numpy()
        for b in range(B):
            assert abs(loss[b] - ref[b][0]) < 1e-13
            assert abs(trace[b] - ref[b][2]) < 1e-12
            # gate-parameter derivatives in the oracle are central differences of the gate matrix (O(h^2))
            assert np.abs(grad[b] - ref[b][1]).max() < (1e-8 if "Q" in slots else 1e-12), (lpp, b)
        # loss-only path agrees bit-for-bit with the loss of the gradient path
        loss2, _, _ = engine.loss_grad(desc, _dev(X), _dev(V), _dev(tgt), cost_kind=cost_kind, want_grad=False, lanes=lpp)
        assert np.array_equal(loss2.cpu().numpy(), loss)


def test_default_target_mapping_and_ragged_batches():
    desc, orc = make_pair("cg", (0.0, 0.0, *BASES["sqCNOT"][:2], BASES["sqCNOT"][2]), k=2)
    rng = np.random.default_rng(11)
    V = O.haar_unitary(rng, 5)
    for B in (1, 31, 32, 33, 1000):
        X = rng.uniform(0, 2 * np.pi, (B, orc.n_params))
        loss, grad, _ = engine.loss_grad(desc, _dev(X), _dev(V))
        lo = np.array([O.loss_and_grad(orc, X[b], V[b % 5])[0] for b in range(min(B, 40))])
        assert np.abs(loss.cpu().numpy()[: lo.size] - lo).max() < 1e-13
        assert grad.shape == (B, orc.n_params) and torch.isfinite(grad).all()
    # empty batch
    loss, grad, _ = engine.loss_grad(desc, _dev(np.zeros((0, orc.n_params))), _dev(V))
    assert loss.numel() == 0 and grad.numel() == 0


def test_gradient_is_consistent_at_scale():
    """Size-independent property at a large batch: directional derivative == finite difference of the loss."""
    desc, orc = make_pair("riswap", (0.5,), k=3)
    g = torch.Generator(device="cuda").manual_seed(1)
    B = 1 << 18
    X = torch.rand((B, orc.n_params), device="cuda", dtype=torch.float64, generator=g) * (2 * np.pi)
    D = torch.randn((B, orc.n_params), device="cuda", dtype=torch.float64, generator=g)
    V = _dev(O.haar_unitary(np.random.default_rng(0), 64))
    loss, grad, _ = engine.loss_grad(desc, X, V)
    h = 1e-6
    lp, _, _ = engine.loss_grad(desc, X + h * D, V, want_grad=False)
    lm, _, _ = engine.loss_grad(desc, X - h * D, V, want_grad=False)
    fd = (lp - lm) / (2 * h)
    an = (grad * D).sum(1)
    assert (fd - an).abs().max().item() < 1e-7
    assert loss.min().item() >= 0.0 and loss.max().item() <= 1.0


def test_invalid_descriptors_raise_value_error():
    desc, orc = make_pair("riswap", (0.5,), k=2)
    desc.k = 0
    with pytest.raises(ValueError):
        engine.template_eval(desc, _dev(np.zeros((1, orc.n_params))))
    desc, orc = make_pair("riswap", (0.5,), k=2)
    with pytest.raises(ValueError):
        engine.loss_grad(desc, _dev(np.zeros((1, orc.n_params))), _dev(O.SWAP[None]), cost_kind=9)
    with pytest.raises(Exception):
        engine.template_eval(desc, torch.zeros((1, orc.n_params), dtype=torch.float64))  # CPU tensor: no fallback

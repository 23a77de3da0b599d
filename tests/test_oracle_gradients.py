"""The oracle gradient that the GPU gradient tests are checked against (oracle.loss_and_grad: closed-form U3/RZ derivative
matrices, Richardson-extrapolated central differences of the scipy-expm gate matrix for 2Q parameters) is itself pinned
here, on the CPU, against plain central differences of the oracle LOSS -- a third, independent route.  Templates: the
parameter-bound smush cases of tests/test_gpu_smush_adjoint.py and the closed-form gates."""
import numpy as np
import pytest

import oracle as O

SMUSH = {
    "smush_sqiSwap_k2": dict(gate_kind="smush", slots=("Q", "Q", np.pi / 2, 0.0, "Q", "Q", "Q", "Q", 0.5), k=2, T=2, no_exterior_1q=True),
    "smush_all_free_k1": dict(gate_kind="smush", slots=("Q",) * 9, k=1, T=2),
    "smush1q_all_free_k1_T3": dict(gate_kind="smush1q", slots=("Q",) * 15, k=1, T=3),
    "smush_vz_k2": dict(gate_kind="smush", slots=("Q", "Q", np.pi / 2, 0.0, "Q", "Q", "Q", "Q", 0.5), k=2, T=2, vz_only=True),
    "cg_bound_k2": dict(gate_kind="cg", slots=("Q", 0.2, np.pi / 4, "Q", 0.5), k=2),
    "riswap_bound_k2": dict(gate_kind="riswap", slots=("Q",), k=2),
}


@pytest.mark.parametrize("name", sorted(SMUSH))
@pytest.mark.parametrize("kind", ["basic", "square", "basic_inverse"])
def test_oracle_gradient_against_central_differences_of_the_loss(name, kind):
    tmpl = O.OracleTemplate(**SMUSH[name])
    rng = np.random.default_rng(sorted(SMUSH).index(name))
    x = rng.uniform(-2.0, 2.0, tmpl.n_params)
    V = O.haar_unitary(rng)
    loss, grad, T = O.loss_and_grad(tmpl, x, V, kind, h_gate=2e-3, richardson=True)

    def f(y):
        return O.cost(tmpl.eval(y), V, kind)

    assert abs(loss - f(x)) < 1e-14
    assert abs(T - np.trace(V.conj().T @ tmpl.eval(x))) < 1e-13
    h = 1e-5
    fd = np.array([(f(x + h * e) - f(x - h * e)) / (2 * h) for e in np.eye(tmpl.n_params)])
    assert np.abs(grad - fd).max() < 2e-9, np.abs(grad - fd).max()


def test_richardson_is_more_accurate_than_the_plain_central_difference():
    """Order check on one smush gate parameter: halving h divides the plain central-difference error by ~4 and the
    Richardson-combined one by ~16 (until round-off), so the extrapolation is doing what its docstring says."""
    tmpl = O.OracleTemplate(**SMUSH["smush_all_free_k1"])
    rng = np.random.default_rng(3)
    x = rng.uniform(-1.5, 1.5, tmpl.n_params)
    V = O.haar_unitary(rng)
    ref = O.loss_and_grad(tmpl, x, V, "basic", h_gate=5e-4, richardson=True)[1]
    e_plain = [np.abs(O.loss_and_grad(tmpl, x, V, "basic", h_gate=h)[1] - ref).max() for h in (4e-2, 2e-2)]
    e_rich = [np.abs(O.loss_and_grad(tmpl, x, V, "basic", h_gate=h, richardson=True)[1] - ref).max() for h in (4e-2, 2e-2)]
    assert 3.0 < e_plain[0] / e_plain[1] < 5.0
    assert e_rich[0] / e_rich[1] > 10.0
    assert e_rich[1] < e_plain[1] * 1e-2

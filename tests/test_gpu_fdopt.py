"""K5c: batched L-BFGS with finite-difference gradients (parameter-bound smush templates, BasicCostInverse) -- the
reference's own algorithm class, scipy BFGS with jac=None (src/slam/optimizer.py:270-278), run per (target, restart)."""
import numpy as np
import pytest
import torch

import oracle as O
from helpers import BASES
from slam_decomposition_b200 import _lib, engine
from slam_decomposition_b200.cost_function import BasicCost, SquareCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

pytestmark = pytest.mark.gpu


def _smush_pair(name, k):
    gc, gg, t = BASES[name]
    T = round(t / 0.25)
    basis = pdv.smush_template(gc, gg, t, k)
    basis.spanning_range = range(k, k + 1)
    orc = O.OracleTemplate("smush", ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,), k=k, T=T, no_exterior_1q=True)
    assert orc.n_params == basis.desc.n_params == 6 * (k - 1) + k * (2 + 2 * T)
    return basis, orc


def test_solver_selection():
    basis, _ = _smush_pair("sqiSwap", 2)
    opt = TemplateOptimizer(basis, BasicCost(), override_fail=True)
    assert opt._solver(basis.desc, _lib.COST_BASIC) == "fd"
    assert opt._solver(basis.desc, _lib.COST_BASIC_INVERSE) == "fd"
    assert opt._solver(basis.desc, _lib.COST_MAKHLIN_FUNCTIONAL) == "nm"
    opt.override_method = "Nelder-Mead"
    assert opt._solver(basis.desc, _lib.COST_BASIC) == "nm"


@pytest.mark.parametrize("central", [False, True])
def test_smush_template_reaches_its_own_unitaries(central):
    """Targets drawn from the template itself (no exterior 1Q gates, so only those are reachable).  From the reference's
    x0 distribution U(-4 pi, 4 pi) this 18-parameter landscape is full of local minima: the oracle's literal scipy BFGS
    gets best-of-8 losses of 8e-4 .. 1e-2 on these very targets (recorded 2026-10, seeds 50+i), so the statistical claim
    is only "no worse than that"; started inside the basin (x_true + 0.2 noise) the FD-gradient L-BFGS must reach 1e-9.
    In both cases the returned Xk must reproduce the reported loss in the oracle."""
    basis, orc = _smush_pair("sqiSwap", 2)  # 18 parameters: 6 interior U3 angles + 2 x (phi_c, phi_g, gx[2], gy[2])
    rng = np.random.default_rng(3)
    Nt = 12
    P = orc.n_params
    X_true = rng.uniform(-1.5, 1.5, (Nt, P))
    V = np.stack([orc.eval(x) for x in X_true])
    np.random.seed(5)
    opt = TemplateOptimizer(basis, BasicCost(), override_fail=True, training_restarts=24)
    opt.fd_central = central
    out = opt.approximate_targets(V, range(2, 3))
    for i in range(Nt):
        U = orc.eval(out["Xk"][i, :P])
        assert abs(O.cost(U, V[i], "basic") - out["loss"][i]) < 1e-10
    assert np.median(out["loss"]) < 5e-3 and out["loss"].max() < 0.1, out["loss"]
    # inside the basin of attraction
    R = 4
    x0 = X_true[:, None, :] + 0.2 * rng.standard_normal((Nt, R, P))
    opts = engine.opt_defaults()
    opts.f_far = 1e-4  # as TemplateOptimizer sets it for this solver
    loss, x, iters = engine.fd_lbfgs_solve(basis.desc, torch.as_tensor(V, device="cuda"), R, opts,
                                           x0=torch.as_tensor(x0, device="cuda"), central=central)
    best = loss.min(dim=1).values.cpu().numpy()
    # forward differences carry 7e-8 of gradient noise (1e-15 / 1.49e-8): the last decade needs the central ones
    assert (best <= (1e-9 if central else 1e-7)).mean() >= 0.75, best
    xs, ls = x.cpu().numpy(), loss.cpu().numpy()
    for i in range(Nt):
        r = int(ls[i].argmin())
        assert abs(O.cost(orc.eval(xs[i, r]), V[i], "basic") - ls[i, r]) < 1e-10


def test_fd_lbfgs_converges_where_the_literal_scipy_bfgs_does():
    """Same template, same targets, the oracle's literal loop (scipy BFGS, jac=None): wherever that reaches 1e-8 within
    its restarts, the device solver (more restarts, same algorithm class) must be <= 1e-9."""
    basis, orc = _smush_pair("sqiSwap", 1)  # 6 parameters (phi_c, phi_g, gx[2], gy[2]); a single parallel-driven gate
    rng = np.random.default_rng(17)
    Nt = 6
    V = np.stack([orc.eval(rng.uniform(-1.0, 1.0, orc.n_params)) for _ in range(Nt)])
    np.random.seed(6)
    opt = TemplateOptimizer(basis, SquareCost(), override_fail=True, training_restarts=32)
    out = opt.approximate_targets(V, range(1, 2))
    for i in range(Nt):
        ref = O.literal_run(lambda k: orc, V[i], range(1, 2), restarts=6, kind="square", rng=np.random.default_rng(100 + i), v2=True)
        if ref.best_result <= 1e-8:
            assert out["loss"][i] <= 1e-9, (i, ref.best_result, out["loss"][i])
        U = orc.eval(out["Xk"][i, : orc.n_params])
        assert abs(O.cost(U, V[i], "square") - out["loss"][i]) < 1e-10


def test_fd_lbfgs_respects_box_bounds_and_explicit_x0():
    basis, orc = _smush_pair("sqiSwap", 1)
    rng = np.random.default_rng(2)
    x_true = rng.uniform(-0.8, 0.8, orc.n_params)
    V = torch.as_tensor(orc.eval(x_true)[None], device="cuda")
    opts = engine.opt_defaults()
    lo = torch.full((orc.n_params,), -1.0, dtype=torch.float64, device="cuda")
    hi = torch.full((orc.n_params,), 1.0, dtype=torch.float64, device="cuda")
    opts.lower, opts.upper = lo.data_ptr(), hi.data_ptr()
    x0 = torch.as_tensor(rng.uniform(-3, 3, (1, 16, orc.n_params)), device="cuda")
    ev = torch.zeros(1, dtype=torch.int64, device="cuda")
    loss, x, iters = engine.fd_lbfgs_solve(basis.desc, V, 16, opts, x0=x0, evals=ev)
    xs = x.cpu().numpy()[0]
    assert np.all(xs >= -1.0) and np.all(xs <= 1.0)
    best = int(loss[0].argmin())
    assert float(loss[0, best]) <= 1e-9
    assert abs(O.cost(orc.eval(xs[best]), orc.eval(x_true), "basic") - float(loss[0, best])) < 1e-10
    # forward differences: at least P + 1 evaluations per accepted iteration
    assert int(ev.item()) >= int(iters.sum().item()) * (orc.n_params + 1)
    # deterministic for a given x0 (without the cross-restart early exit, whose timing is not)
    opts.early_exit = 0
    loss2, x2, _ = engine.fd_lbfgs_solve(basis.desc, V, 16, opts, x0=x0)
    loss3, x3, _ = engine.fd_lbfgs_solve(basis.desc, V, 16, opts, x0=x0)
    assert torch.equal(loss2, loss3) and torch.equal(x2, x3)


def test_fd_lbfgs_rejects_coordinate_based_costs():
    basis, _ = _smush_pair("sqiSwap", 1)
    V = torch.as_tensor(O.ISWAP[None].astype(np.complex128), device="cuda")
    opts = engine.opt_defaults()
    opts.cost_kind = _lib.COST_MAKHLIN_FUNCTIONAL
    with pytest.raises(NotImplementedError):  # SLAM_ERR_UNSUPPORTED
        engine.fd_lbfgs_solve(basis.desc, V, 2, opts)

"""In-basin K5c adjoint problem of tests/test_gpu_smush_adjoint.py per history length (how many targets reach 1e-10)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from slam_decomposition_b200 import engine
from test_gpu_smush_adjoint import _smush_pair
basis, orc = _smush_pair("sqiSwap", 2)
rng = np.random.default_rng(3)
Nt, R, P = 12, 4, orc.n_params
X_true = rng.uniform(-1.5, 1.5, (Nt, P))
V = torch.as_tensor(np.stack([orc.eval(x) for x in X_true]), device="cuda")
x0 = torch.as_tensor(X_true[:, None, :] + 0.2 * rng.standard_normal((Nt, R, P)), device="cuda")
for m in (8, 7, 6, 5):
    opts = engine.opt_defaults(); opts.f_far = 1e-4; opts.history = m; opts.diag = 1
    ev = torch.zeros(1, dtype=torch.int64, device="cuda")
    loss, x, it = engine.fd_lbfgs_solve(basis.desc, V, R, opts, x0=x0, central="adjoint", evals=ev)
    best = loss.min(dim=1).values.cpu().numpy()
    reason = (it.cpu().numpy() >> 24).ravel()
    print(f"m={m}: <=1e-10 {(best <= 1e-10).sum()}/12, worst {best.max():.2e}, evals {ev.item()}, reasons {np.bincount(reason)}", flush=True)

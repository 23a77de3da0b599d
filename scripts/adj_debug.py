"""Stop reasons / iteration counts of the K5c adjoint run on the in-basin smush problem (SLAM_B200_FD_DEBUG=1)."""
import os, sys
os.environ["SLAM_B200_FD_DEBUG"] = "1"
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
for mode in ("adjoint", True):
    opts = engine.opt_defaults(); opts.f_far = 1e-4; opts.early_exit = 0
    ev = torch.zeros(1, dtype=torch.int64, device="cuda")
    loss, x, iters = engine.fd_lbfgs_solve(basis.desc, V, R, opts, x0=x0, central=mode, evals=ev)
    it = iters.cpu().numpy(); l = loss.cpu().numpy()
    print("mode", mode, "evals", ev.item())
    for i in range(Nt):
        print(i, " ".join(f"{l[i,r]:.1e}/it{it[i,r] & 0xffffff}/r{it[i,r] >> 24}" for r in range(R)))

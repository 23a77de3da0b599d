"""Stop-reason histogram of the FD-gradient L-BFGS (K5c) on the in-basin smush case of tests/test_gpu_fdopt.py."""
import os, sys
os.environ["SLAM_B200_FD_DEBUG"] = "1"
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from helpers import BASES
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
gc, gg, t = BASES["sqiSwap"]; k = 2; T = 2
basis = pdv.smush_template(gc, gg, t, k)
orc = O.OracleTemplate("smush", ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,), k=k, T=T, no_exterior_1q=True)
rng = np.random.default_rng(3)
Nt, P, R = 12, orc.n_params, 4
X_true = rng.uniform(-1.5, 1.5, (Nt, P))
V = np.stack([orc.eval(x) for x in X_true])
x0 = X_true[:, None, :] + 0.2 * rng.standard_normal((Nt, R, P))
for central, gfar in ((False, 1e-5), (False, 3e-7), (True, 3e-7), (True, 1e-8)):
    opts = engine.opt_defaults(); opts.gtol_far = gfar
    ev = torch.zeros(1, dtype=torch.int64, device="cuda")
    loss, x, iters = engine.fd_lbfgs_solve(basis.desc, torch.as_tensor(V, device="cuda"), R, opts,
                                           x0=torch.as_tensor(x0, device="cuda"), central=central, evals=ev)
    it = iters.cpu().numpy(); ls = loss.cpu().numpy()
    best = ls.min(axis=1)
    reasons = np.bincount((it >> 24).ravel(), minlength=9)
    print(f"central {central} gtol_far {gfar:g}: evals {int(ev)}  best<=1e-9: {(best <= 1e-9).mean():.2f}  best<=1e-10: {(best <= 1e-10).mean():.2f} "
          f"median best {np.median(best):.2e}  mean iters {(it & 0xffffff).mean():.0f}  reasons {reasons.tolist()}")

import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from helpers import make_pair, BASES
from slam_decomposition_b200 import engine
Nt, R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 16
KS = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 2, 3, 4, 5, 6)
V = torch.as_tensor(O.haar_unitary(np.random.default_rng(42), min(Nt, 4096)), device="cuda")
if Nt > V.shape[0]: V = V.repeat((Nt + V.shape[0] - 1) // V.shape[0], 1, 1)[:Nt].contiguous()
for name, kind, slots in (("sqiswap", "riswap", (0.5,)), ("sqCNOT", "cg", (0.0, 0.0, *BASES["sqCNOT"]))):
    for k in KS:
        desc, orc = make_pair(kind, slots, k=k)
        for ee in (0, 1):
            opts = engine.opt_defaults(); opts.early_exit = ee
            ev = torch.zeros(1, dtype=torch.int64, device="cuda")
            if ee == 0 and k == KS[0]: engine.lbfgs_solve(desc, V[:64], R, opts, seed=1)  # warm-up (module load)
            torch.cuda.synchronize(); t0 = time.time()
            loss, x, iters = engine.lbfgs_solve(desc, V, R, opts, seed=43, evals=ev)
            torch.cuda.synchronize(); dt = time.time() - t0
            best = loss.min(1).values
            ok = (best < 1e-10).float().mean().item()
            okr = (loss < 1e-10).float().mean().item()
            n = ev.item()
            print(f"{name} k={k} ee={ee}: {dt*1e3:8.1f} ms  evals={n:.3e} ({n/dt/1e9:.3f} Gev/s, {n/dt*O.F_lossgrad(k)/1e12:.1f} TF alg)  "
                  f"solved targets={ok:.4f} restarts ok={okr:.3f}  iters mean={iters.float().mean().item():.1f} max={iters.max().item()}  "
                  f"median best={best.median().item():.2e} it>=1000: {(iters>=1000).sum().item()}")

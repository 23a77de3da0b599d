"""K5c adjoint-mode rates on the parallel-drive Monte-Carlo templates (no bounds), per history length:
usage: python scripts/k5c_bench.py [targets] [history ...]   (default 131072 targets x 8 restarts, history 8)"""
import os, sys, math, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
hists = [int(a) for a in sys.argv[2:]] or [8]
for (gc, gg, t, k, name) in ((math.pi/2, 0.0, 0.5, 2, "sqiSwap k2 P18"), (math.pi/2, 0.0, 0.5, 3, "sqiSwap k3 P30"), (math.pi/4, math.pi/4, 0.5, 4, "sqCNOT k4 P42")):
    b = pdv.smush_template(gc, gg, t, k)
    rng = np.random.default_rng(3)
    n = Nt if b.desc.n_params <= 32 else Nt // 4
    Vt = engine.template_eval(b.desc, torch.as_tensor(rng.uniform(-1.5, 1.5, (n, b.desc.n_params)), device=dev))
    for m in hists:
        o = engine.opt_defaults(); o.f_far = 1e-4; o.x0_lo, o.x0_hi = -4*math.pi, 4*math.pi; o.history = m
        ev = torch.zeros(1, dtype=torch.int64, device=dev)
        for rep in range(2):
            ev.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
            loss, x, it = engine.fd_lbfgs_solve(b.desc, Vt, 8, o, seed=11, central="adjoint", evals=ev)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = loss.min(dim=1).values
        print(f"{name} m={m}: {n} targets {dt*1e3:8.1f} ms  {int(ev.item())/dt/1e6:7.2f} M evals/s  evals {int(ev.item())}  "
              f"solved<=1e-9 {(best<=1e-9).float().mean().item():.4f}  mean iters {it.float().mean().item():.1f}", flush=True)

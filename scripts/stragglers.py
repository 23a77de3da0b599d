import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import bench
from helpers import make_pair, BASES
from slam_decomposition_b200 import engine
Nt, R = int(sys.argv[1]), 16
V = torch.as_tensor(bench.haar_targets(Nt, 42), device="cuda")
for k in (1, 2, 3, 4, 5):
    desc, orc = make_pair("cg", (0.0, 0.0, *BASES["sqCNOT"]), k=k)
    for seed in (1, 2):
        opts = engine.opt_defaults(); opts.early_exit = 1
        torch.cuda.synchronize(); t0 = time.time()
        loss, x, iters = engine.lbfgs_solve(desc, V, R, opts, seed=seed)
        torch.cuda.synchronize(); dt = time.time() - t0
        it = iters.flatten().cpu().numpy(); ls = loss.flatten().cpu().numpy()
        hist = np.histogram(it, bins=[0, 50, 100, 200, 400, 800, 1600, 2499, 2501])[0]
        idx = np.argsort(-it)[:6]
        loss_g, grad, _ = engine.loss_grad(desc, x.reshape(-1, desc.n_params)[torch.as_tensor(idx, device="cuda")].contiguous(),
                                           V[torch.as_tensor(idx // R, device="cuda")].contiguous(), tgt_idx=torch.arange(len(idx), dtype=torch.int32, device="cuda"))
        print(f"k={k} seed={seed} {dt*1e3:.0f} ms hist={hist.tolist()} top: " + ", ".join(f"({it[i]}, f={ls[i]:.2e}, |g|={grad[j].abs().max().item():.1e})" for j, i in enumerate(idx)))

"""One K5c adjoint-mode launch at bench size (sqrt(iSWAP) k=2 smush template, 131072 targets x 8 restarts) for ncu."""
import math, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
basis = pdv.smush_template(math.pi / 2, 0.0, 0.5, 2)
rng = np.random.default_rng(3)
Nt, R, P = 131072, 8, basis.desc.n_params
Vt = engine.template_eval(basis.desc, torch.as_tensor(rng.uniform(-1.5, 1.5, (Nt, P)), device=dev))
opts = engine.opt_defaults(); opts.f_far = 1e-4; opts.x0_lo, opts.x0_hi = -4 * math.pi, 4 * math.pi
engine.fd_lbfgs_solve(basis.desc, Vt, R, opts, seed=11, central="adjoint")
torch.cuda.synchronize()
print("done")

"""Timing of the thread-per-problem smush kernels: K1/K2 streaming (eval, loss, loss + adjoint gradient), K5c (adjoint and
central-difference L-BFGS) and K5b (Nelder-Mead) on the sqrt(iSWAP) k=2 smush template.  A/B builds: SLAM_B200_LIB=..."""
import math, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from slam_decomposition_b200 import _lib, engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
print("lib:", _lib.LIB_PATH)
dev = engine.require_cuda()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
basis3 = pdv.smush_template(math.pi / 2, 0.0, 0.5, 3)
B = 1 << 20
g = torch.Generator(device=dev).manual_seed(77)
X = (torch.rand((B, basis3.desc.n_params), device=dev, dtype=torch.float64, generator=g) - 0.5) * (8 * math.pi)
rng = np.random.default_rng(0)
V = torch.as_tensor(np.stack([O.haar_unitary(rng) for _ in range(64)]), device=dev)
print(f"K1 smush eval k3      : {B / timed(lambda: engine.template_eval(basis3.desc, X)) / 1e3:8.1f} Mevals/s")
print(f"K2 smush loss k3      : {B / timed(lambda: engine.loss_grad(basis3.desc, X, V, want_grad=False)) / 1e3:8.1f} Mevals/s")
print(f"K2 smush loss+grad k3 : {B / timed(lambda: engine.loss_grad(basis3.desc, X, V)) / 1e3:8.1f} Mevals/s")
basis = pdv.smush_template(math.pi / 2, 0.0, 0.5, 2)
orc = O.OracleTemplate("smush", ("Q", "Q", math.pi / 2, 0.0, "Q", "Q", "Q", "Q", 0.5), k=2, T=2, no_exterior_1q=True)
Nt, R, P = 4096, 8, orc.n_params
rng = np.random.default_rng(3)
Xt = rng.uniform(-1.5, 1.5, (Nt, P))
Vt = engine.template_eval(basis.desc, torch.as_tensor(Xt, device=dev))
x0 = torch.as_tensor(Xt[:, None, :] + 0.2 * rng.standard_normal((Nt, R, P)), device=dev)
for mode in ("adjoint", True):
    opts = engine.opt_defaults(); opts.f_far = 1e-4
    ev = torch.zeros(1, dtype=torch.int64, device=dev)
    ms = timed(lambda: engine.fd_lbfgs_solve(basis.desc, Vt, R, opts, x0=x0, central=mode, evals=ev), reps=1)
    loss, _, _ = engine.fd_lbfgs_solve(basis.desc, Vt, R, opts, x0=x0, central=mode)
    print(f"K5c {str(mode):8s}: {ms:8.1f} ms for {Nt} targets x {R} restarts, solved<=1e-9: {(loss.min(dim=1).values <= 1e-9).float().mean().item():.3f}, evals/run {ev.item() // 2}")
nm = engine.nm_defaults(); nm.cost_kind = _lib.COST_BASIC; nm.max_iter = 2500
ms = timed(lambda: engine.nm_solve(basis.desc, Vt[:1024], 4, nm, x0=x0[:1024, :4].contiguous()), reps=1)
print(f"K5b Nelder-Mead: {ms:8.1f} ms for 1024 targets x 4 restarts")
# throughput at scale (several waves of problems per thread): random starts, the reference's U(-4pi, 4pi) box
for NtL in (32768, 131072):
    VL = Vt[torch.arange(NtL, device=dev) % Nt].contiguous()
    opts = engine.opt_defaults(); opts.f_far = 1e-4; opts.x0_lo, opts.x0_hi = -4 * math.pi, 4 * math.pi
    ev = torch.zeros(1, dtype=torch.int64, device=dev)
    ms = timed(lambda: engine.fd_lbfgs_solve(basis.desc, VL, R, opts, seed=11, central="adjoint", evals=ev), reps=1)
    print(f"K5c adjoint at scale: {NtL} targets x {R} restarts: {ms:8.1f} ms, {ev.item() / 2 / ms / 1e3:6.1f} M loss+grad/s")

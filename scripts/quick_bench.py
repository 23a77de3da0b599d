import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from helpers import make_pair
from slam_decomposition_b200 import engine
pk, ms = engine.fp64_peak(8192)
print(f"fp64 peak {pk/1e12:.2f} TFLOP/s ({ms:.2f} ms)")
V = torch.as_tensor(O.haar_unitary(np.random.default_rng(0), 4096), device="cuda")
for k in (1, 3, 6):
    desc, orc = make_pair("riswap", (0.5,), k=k)
    B = 1 << 22
    X = torch.rand((B, orc.n_params), device="cuda", dtype=torch.float64) * 6.28
    loss = torch.empty(B, device="cuda", dtype=torch.float64); grad = torch.empty_like(X)
    for lpp in ("4", "2", "1"):
        os.environ["SLAM_B200_LPP"] = lpp
        for want_grad in (True, False):
            for _ in range(2): engine.loss_grad(desc, X, V, want_grad=want_grad, out_loss=loss, out_grad=grad)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): engine.loss_grad(desc, X, V, want_grad=want_grad, out_loss=loss, out_grad=grad)
            e1.record(); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 5 * 1e-3
            F = O.F_lossgrad(k) if want_grad else O.F_eval(k)
            print(f"k={k} lpp={lpp} grad={want_grad}: {B/t/1e9:.3f} Gevals/s  {B/t*F/1e12:.2f} TFLOP/s alg ({B/t*F/pk*100:.1f}% of measured DFMA peak)")

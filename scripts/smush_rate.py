"""Streaming rates of the smush K2 kernels (loss, loss + adjoint gradient) for sqrt(iSWAP) k = 2, 3 (T = 2) and the K5c bench case."""
import math, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
V = torch.as_tensor(bench.haar_targets(4096, 5678), device=dev)
for k in (2, 3):
    basis = pdv.smush_template(math.pi / 2, 0.0, 0.5, k)
    B = 1 << 20
    g = torch.Generator(device=dev).manual_seed(77)
    X = (torch.rand((B, basis.desc.n_params), device=dev, dtype=torch.float64, generator=g) - 0.5) * (8 * math.pi)
    for want in (True, False):
        for _ in range(2):
            engine.loss_grad(basis.desc, X, V, want_grad=want)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            engine.loss_grad(basis.desc, X, V, want_grad=want)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"k={k} P={basis.desc.n_params} {'loss+grad' if want else 'loss only'}: {B / dt / 1e6:8.1f} M rows/s", flush=True)

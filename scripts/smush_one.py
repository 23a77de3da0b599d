"""One launch each of the smush K2 kernels (loss + adjoint gradient, loss only) for ncu: sqrt(iSWAP) k=3 T=2, P=30."""
import math, os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
basis = pdv.smush_template(math.pi / 2, 0.0, 0.5, 3)
B = 1 << 20
g = torch.Generator(device=dev).manual_seed(77)
X = (torch.rand((B, basis.desc.n_params), device=dev, dtype=torch.float64, generator=g) - 0.5) * (8 * math.pi)
V = torch.as_tensor(bench.haar_targets(4096, 5678), device=dev)
engine.loss_grad(basis.desc, X, V)
engine.loss_grad(basis.desc, X, V, want_grad=False)
torch.cuda.synchronize()
print("done")

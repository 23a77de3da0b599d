"""One launch each of the K3 Weyl kernel (2^21 Haar matrices) and the K4b trajectory kernel (2^18 trajectories, N=10, R=5) for ncu."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from slam_decomposition_b200 import engine
dev = engine.require_cuda()
U = torch.as_tensor(bench.haar_targets(1 << 21, 99), device=dev)
engine.weyl(U)
B, N = 1 << 18, 10
g = torch.Generator(device=dev).manual_seed(5)
gate = (torch.rand((B, 8), device=dev, dtype=torch.float64, generator=g) - 0.5) * 4
gx = (torch.rand((B, N), device=dev, dtype=torch.float64, generator=g) - 0.5) * 12
gy = (torch.rand((B, N), device=dev, dtype=torch.float64, generator=g) - 0.5) * 12
engine.pd_trajectory(gate, gx, gy, 0.1, R=5, want_final=False)
torch.cuda.synchronize()
print("done")

"""K4b parallel-drive trajectory (pd_playground.py:169-208): N = 10 slices, R = 5 sub-times, 2^18 trajectories; A/B of the
phase-locked kernel (SLAM_B200_TRAJ_SYNC)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from slam_decomposition_b200 import engine
dev = engine.require_cuda()
B, N, R = 1 << 18, 10, 5
g = torch.Generator(device=dev).manual_seed(5)
gate = (torch.rand((B, 8), device=dev, dtype=torch.float64, generator=g) - 0.5) * 4
gx = (torch.rand((B, N), device=dev, dtype=torch.float64, generator=g) - 0.5) * 12
gy = (torch.rand((B, N), device=dev, dtype=torch.float64, generator=g) - 0.5) * 12
ref = None
for rnd in range(2):
    for sync in ("0", "1"):
        os.environ["SLAM_B200_TRAJ_SYNC"] = sync
        out = engine.pd_trajectory(gate, gx, gy, 0.1, R=R)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): out = engine.pd_trajectory(gate, gx, gy, 0.1, R=R)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        c = out[0] if isinstance(out, tuple) else out
        if ref is None: ref = c.clone()
        print(f"TRAJ_SYNC={sync}: {ms:7.2f} ms, {B * N * R / ms / 1e3:7.1f} M trajectory points/s, identical to first run: {bool(torch.equal(ref, c))}", flush=True)

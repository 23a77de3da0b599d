"""A/B of the coverage / Weyl kernels' occupancy variants (SLAM_B200_COV_MINB, SLAM_B200_WEYL_MINB), CUDA-event timed."""
import math, os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
n = 10_000_000
cases = [("sqiSwap k3 plain", pdv.plain_template(math.pi / 2, 0.0, 0.5, 3)),
         ("CNOT    k3 plain", pdv.plain_template(math.pi / 4, math.pi / 4, 1.0, 3)),
         ("sqiSwap k3 smush", pdv.smush_template(math.pi / 2, 0.0, 0.5, 3)),
         ("CNOT    k2 smush", pdv.smush_template(math.pi / 4, math.pi / 4, 1.0, 2))]
hist = torch.zeros(128 ** 3, dtype=torch.int64, device=dev)
def timed(fn, reps=3):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
best = {}
for rnd in range(3):  # interleaved rounds, best-of: clock ramp and order effects cancel
    for sync in ("0", "1"):
        os.environ["SLAM_B200_COV_SYNC"] = sync
        for mb in ("2", "3", "4"):
            os.environ["SLAM_B200_COV_MINB"] = mb
            for name, basis in cases:
                ms = timed(lambda: pdv.coverage_histogram(basis, n, seed=2, hist=hist))
                best[(name, sync, mb)] = min(ms, best.get((name, sync, mb), 1e9))
for (name, sync, mb), ms in sorted(best.items()):
    print(f"{name} SYNC={sync} MINB={mb}: {n / ms / 1e3:8.1f} Msamples/s", flush=True)
U = torch.as_tensor(bench.haar_targets(1 << 21, 99), device=dev)
for mb in ("4", "3"):
    os.environ["SLAM_B200_WEYL_MINB"] = mb
    ms = timed(lambda: engine.weyl(U), 5)
    print(f"WEYL_MINB={mb}: {U.shape[0] / ms / 1e3:8.1f} Mmatrices/s", flush=True)

"""One coverage launch per variant (for ncu): sqCNOT k=3 plain and smush, 4e6 samples."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
n = 4_000_000
for kind in ("smush", "plain"):
    basis = pdv.smush_template(np.pi / 4, np.pi / 4, 0.5, 3) if kind == "smush" else pdv.plain_template(np.pi / 4, np.pi / 4, 0.5, 3)
    hist = torch.zeros(128 ** 3, dtype=torch.int64, device=dev)
    pdv.coverage_histogram(basis, n, seed=2, hist=hist)
    torch.cuda.synchronize()
print("done")

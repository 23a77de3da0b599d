import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
dev = engine.require_cuda()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pk, _ = engine.fp64_peak(4096)
for (gc, gg, t, name, iters) in pdv.GATE_LIST:
    for k in (1, 2, 3):
        for kind in ("plain", "smush"):
            if kind == "plain" and k == 1: continue
            basis = pdv.plain_template(gc, gg, t, k) if kind == "plain" else pdv.smush_template(gc, gg, t, k)
            hist = torch.zeros(128 ** 3, dtype=torch.int64, device=dev)
            pdv.coverage_histogram(basis, 100000, seed=1, hist=hist)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            pdv.coverage_histogram(basis, n, seed=2, hist=hist)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            T = round(t / pdv.duration_1q)
            print(f"{name:8s} k={k} {kind:5s} P={basis.desc.n_params:3d} T={T}: {n/dt/1e6:9.1f} Msamples/s  ({dt*1e3:7.1f} ms for {n:.0e})  occupied bins={int((hist>0).sum())}", flush=True)

"""BASELINE configs[4]: coverage-set sweep of the sqCNOT parallel-drive (smush) templates k = 1..6, `--samples` samples per k
sharded over the ranks (contiguous slices of one Philox stream), one int64 histogram all-reduce per k.
  1 GPU : python scripts/coverage_sweep.py --samples 1e9
  N GPUs: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/coverage_sweep.py --samples 1e9
Prints per k: wall time (device-synchronised, max over ranks), aggregate samples/s, occupied voxels and the Haar-weighted
chamber fraction they cover (voxel analogue of the reference's hull volumes, parallel_drive_volume.py:343-378)."""
import argparse, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from slam_decomposition_b200 import distributed as D, engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=float, default=1e9)
ap.add_argument("--kmax", type=int, default=6)
ap.add_argument("--nbins", type=int, default=128)
args = ap.parse_args()
rank, world, local = D.init_from_env()
dev = engine.require_cuda()
n = int(args.samples)
gc, gg, t = np.pi / 4, np.pi / 4, 0.5  # sqCNOT (parallel_drive_volume.py:94)
total_t = 0.0
for k in range(1, args.kmax + 1):
    basis = pdv.smush_template(gc, gg, t, k)
    pdv.coverage_sweep(basis, 100000, seed=1, nbins=args.nbins)  # warm-up
    torch.cuda.synchronize(); D.barrier(); t0 = time.perf_counter()
    hist = pdv.coverage_sweep(basis, n, seed=2023 + k, nbins=args.nbins)
    torch.cuda.synchronize(); dt = D.max_over_ranks(time.perf_counter() - t0, dev)
    total_t += dt
    if rank == 0:
        assert int(hist.sum()) == n
        print(f"sqCNOT smush k={k} P={basis.desc.n_params:2d}: {n:.0e} samples on {world} GPU(s) in {dt * 1e3:8.1f} ms = {n / dt / 1e9:6.3f} G samples/s; "
              f"occupied voxels {int((hist > 0).sum()):7d}, Haar-weighted chamber fraction {pdv.haar_volume_fraction(hist, args.nbins):.4f}", flush=True)
if rank == 0:
    print(f"total {total_t:.2f} s for k = 1..{args.kmax}")
D.shutdown()

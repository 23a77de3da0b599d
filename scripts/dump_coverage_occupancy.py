"""Coverage histograms for every basis gate of parallel_drive_volume.py:91-97 and every k below the gate's full-coverage
size, smush (parallel-drive) and plain templates, reduced to packed occupancy bitmaps + counts and written to
gpurun_out/coverage_occupancy.npz: the input of the hull / Haar-volume post-processing (pdv.hull_coverage), which runs
on the host.  Usage: python scripts/dump_coverage_occupancy.py [samples_per_cloud]"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    nbins = 128
    out = {}
    for gc, gg, t, name, iters in pdv.GATE_LIST:
        for k in range(1, iters):
            for kind, make in (("smush", pdv.smush_template), ("plain", pdv.plain_template)):
                basis = make(gc, gg, t, k)
                t0 = time.perf_counter()
                hist = pdv.coverage_histogram(basis, n, seed=2023, nbins=nbins)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                occ = (hist > 0).cpu().numpy()
                out[f"{name}_k{k}_{kind}"] = np.packbits(occ)
                print(f"{name:8s} k={k} {kind:5s} P={basis.desc.n_params:3d} {n / dt / 1e6:8.1f} M samples/s "
                      f"occupied voxels {int(occ.sum())}", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/coverage_occupancy.npz", nbins=nbins, samples=n, **out)


if __name__ == "__main__":
    main()

"""Coverage volumes in the reference's protocol (parallel_drive_volume.py:88-410) next to its recorded table
(tests/golden/extended_results.json = src/slam/data/extended_results.json): N = 3000 samples per (gate, k) for several seeds
(the reference's own sample size; its hulls are far from converged, so the comparison must use the same N), and one large-N
row showing where the hull converges to.  Usage: python scripts/coverage_volumes.py [seeds] [large_n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np

from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    big = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
    ref = json.load(open("tests/golden/extended_results.json"))
    out = {}
    for gc, gg, t, name, iters in pdv.GATE_LIST:
        for k in range(1, iters):
            t0 = time.perf_counter()
            rows = [pdv.coverage_study(gc, gg, t, k, seed=100 + 7 * s, exact_flags=(s == 0)) for s in range(seeds)]
            ext = np.array([r[1] for r in rows])
            conv = pdv.coverage_study(gc, gg, t, k, n_samples=big, seed=5, exact_flags=False)
            r = ref[name][str(k)]
            print(f"{name:8s} k={k} ref [{r[0]:.4f} {r[1]:.4f} {[bool(x) for x in r[2:]]}]  N=3000: base {rows[0][0]:.4f} "
                  f"ext {ext.mean():.4f} +- {ext.std():.4f} (min {ext.min():.4f} max {ext.max():.4f}) flags {rows[0][2:]}  "
                  f"N={big:.0e}: ext {conv[1]:.4f}  ({time.perf_counter() - t0:.1f} s)", flush=True)
            out[f"{name}_{k}"] = {"ref": r, "n3000": rows, "large_n": conv, "large_n_samples": big}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/coverage_volumes.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

"""A/B harness for the K5 launch configuration (lanes per problem, exact-length kernels, history storage, history
length): runs the headline sweep (Haar targets x 16 restarts onto sqCNOT templates k = 1..6) under each setting of the
SLAM_B200_LBFGS_* environment overrides and prints per-k kernel time, evaluations and the solved fraction.
Usage: python scripts/lbfgs_config_sweep.py [targets] [config-name ...]"""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate

CONFIGS = {
    "x384": {},
    "x384_f32": {"SLAM_B200_LBFGS_HIST": "0"},
    "generic": {"SLAM_B200_LBFGS_EXACT": "0"},
    "x512_m6": {"SLAM_B200_LBFGS_MAXT": "512", "SLAM_B200_LBFGS_MMIN": "6"},
    "x512_m5": {"SLAM_B200_LBFGS_MAXT": "512", "SLAM_B200_LBFGS_MMIN": "5"},
    "x512_m4": {"SLAM_B200_LBFGS_MAXT": "512", "SLAM_B200_LBFGS_MMIN": "4"},
    "x512_m3": {"SLAM_B200_LBFGS_MAXT": "512", "SLAM_B200_LBFGS_MMIN": "3"},
    "lpp2_m3": {"SLAM_B200_LBFGS_LPP": "2", "SLAM_B200_LBFGS_MMIN": "3"},
}
KEYS = ("SLAM_B200_LBFGS_LPP", "SLAM_B200_LBFGS_EXACT", "SLAM_B200_LBFGS_HIST", "SLAM_B200_LBFGS_MMIN", "SLAM_B200_LBFGS_TEAMS",
        "SLAM_B200_LBFGS_MAXT")


def main():
    Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    names = sys.argv[2:] or list(CONFIGS)
    dev = engine.require_cuda()
    basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
    opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
    V = torch.as_tensor(bench.haar_targets(Nt, 42), device=dev)
    out = {}
    for name in names:
        for key in KEYS:
            os.environ.pop(key, None)
        os.environ.update(CONFIGS[name])
        np.random.seed(7)
        opt._run_batch(V, range(1, 7))  # warm-up
        torch.cuda.synchronize()
        engine.LBFGS_EVENTS = []
        opt.launch_evals = []
        reps = 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            res = opt._run_batch(V, range(1, 7))
        e1.record()
        torch.cuda.synchronize()
        ev, le = engine.LBFGS_EVENTS, list(opt.launch_evals)
        engine.LBFGS_EVENTS = None
        per_k = {}
        for (k, a, b), (_, n) in zip(ev, le):
            d = per_k.setdefault(k, [0.0, 0])
            d[0] += a.elapsed_time(b) / reps
            d[1] += n / reps
        solved = float((res["best_loss"] <= 1e-10).mean())
        mean_k = float(res["best_k"].astype(np.float64).mean())
        total = e0.elapsed_time(e1) / reps
        kern = sum(v[0] for v in per_k.values())
        line = " ".join(f"k{k}: {v[0]:6.2f} ms {v[1] / 1e6:6.2f} Mev {v[1] / v[0] / 1e6:5.2f} Gev/s |" for k, v in sorted(per_k.items()))
        print(f"{name:20s} sweep {total:7.2f} ms  kernels {kern:7.2f} ms  solved {solved:.5f} mean_k {mean_k:.4f}\n    {line}", flush=True)
        out[name] = {"sweep_ms": total, "kernel_ms": kern, "solved": solved, "mean_k": mean_k,
                     "per_k": {k: {"ms": v[0], "evals": v[1]} for k, v in per_k.items()}}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/lbfgs_config_sweep.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

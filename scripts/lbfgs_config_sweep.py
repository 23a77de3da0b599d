"""A/B harness for the K5 launch configuration: runs the headline sweep (Haar targets x 16 restarts onto sqCNOT templates
k = 1..6, chained launches) under each setting of the explicit SlamOptOpts.tune_* fields and prints the sweep time, the
per-k charged time, evaluations and the solved fraction.
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
    "auto": {},
    "m4": {"tune_hist_min": 4},
    "h3": {"history": 3},
    "h4": {"history": 4},
    "h5": {"history": 5},
    "w12": {"tune_sm_threads": 384},
    "lpp2": {"tune_lanes": 2},
    "teams96": {"tune_max_teams": 96},
    "nopipe": {"_pipeline": False},
}


def main():
    Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    names = sys.argv[2:] or list(CONFIGS)
    dev = engine.require_cuda()
    basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
    V = torch.as_tensor(bench.haar_targets(Nt, 42), device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    out = {}
    for name in names:
        cfg = dict(CONFIGS[name])
        opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
        opt.pipeline = cfg.pop("_pipeline", True)
        opt.tune = cfg
        np.random.seed(7)
        for _ in range(2):
            opt._run_batch(V, range(1, 7))  # warm-up
        torch.cuda.synchronize()
        engine.LBFGS_EVENTS = []
        engine.LBFGS_SPANS = []
        opt.launch_evals = []
        reps = 3
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = opt._run_batch(V, range(1, 7))
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ev, le, spans = engine.LBFGS_EVENTS, list(opt.launch_evals), engine.LBFGS_SPANS
        engine.LBFGS_EVENTS = None
        engine.LBFGS_SPANS = None
        per_k = {}
        per_sweep = len(ev) // reps
        for idx, ((k, a, b), (_, n)) in enumerate(zip(ev, le)):
            d = per_k.setdefault(k, [0.0, 0])
            if spans:
                prev_end = spans[idx // per_sweep][0] if idx % per_sweep == 0 else ev[idx - 1][2]
                d[0] += max(prev_end.elapsed_time(b), 0.0) / reps
            else:
                d[0] += a.elapsed_time(b) / reps
            d[1] += n / reps
        solved = float((res["best_loss"] <= 1e-10).mean())
        mean_k = float(res["best_k"].astype(np.float64).mean())
        total = tot / reps
        kern = (sum(a.elapsed_time(b) for a, b in spans) / reps) if spans else sum(v[0] for v in per_k.values())
        evals = sum(v[1] for v in per_k.values())
        line = " ".join(f"k{k}: {v[0]:6.2f} ms {v[1] / 1e6:6.2f} Mev |" for k, v in sorted(per_k.items()))
        print(f"{name:12s} sweep {total:7.2f} ms  kernels {kern:7.2f} ms  {evals / total / 1e6:7.3f} Gev/s  solved {solved:.5f} "
              f"mean_k {mean_k:.4f}\n    {line}", flush=True)
        out[name] = {"sweep_ms": total, "kernel_ms": kern, "solved": solved, "mean_k": mean_k, "evals": evals,
                     "per_k": {k: {"ms": v[0], "evals": v[1]} for k, v in per_k.items()}}
    os.makedirs("gpurun_out", exist_ok=True)
    tag = os.environ.get("SWEEP_TAG", "")
    with open(f"gpurun_out/lbfgs_config_sweep{tag}.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

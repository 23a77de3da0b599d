"""One chained sqCNOT sweep (k = 1..6) for ncu captures of lbfgs_kernel.  Usage: python scripts/k5_one.py [targets] [cta_warps]"""
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

Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = engine.require_cuda()
basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
if len(sys.argv) > 2:
    opt.tune = {"tune_cta_warps": int(sys.argv[2])}
opt.pipeline = False  # one launch at a time: per-kernel counters are not mixed with a co-running launch
np.random.seed(7)
V = torch.as_tensor(bench.haar_targets(Nt, 42), device=dev)
res = opt._run_batch(V, range(1, 7))
torch.cuda.synchronize()
print("solved", float((res["best_loss"] <= 1e-10).mean()), "evals", opt.last_stats["evals"])

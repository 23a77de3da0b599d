import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate
dev = engine.require_cuda()
Nt = int(sys.argv[1])
basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
V = torch.as_tensor(bench.haar_targets(Nt, 42), device=dev)
for _ in range(2): opt._run_batch(V, range(1, 7))
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); opt._run_batch(V, range(1, 7)); torch.cuda.synchronize(); print("wall", time.perf_counter() - t0)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

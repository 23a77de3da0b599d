import os, sys, time, math
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate
dev = engine.require_cuda()
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
V_host = torch.as_tensor(bench.haar_targets(Nt, 42)).pin_memory()
V_dev = V_host.to(dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for name, fn in (("resident", lambda: opt._run_batch(V_dev, range(1, 7))), ("host_api", lambda: opt.approximate_targets(V_host, range(1, 7)))):
    for rep in range(3):
        t0 = T(); fn(); t1 = T()
        print(name, rep, f"{(t1-t0)*1e3:.1f} ms", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); opt.approximate_targets(V_host, range(1, 7)); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

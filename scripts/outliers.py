import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate
dev = engine.require_cuda()
Nt = int(sys.argv[1]); steps = int(sys.argv[2])
basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
V = torch.as_tensor(bench.haar_targets(Nt, 42), device=dev)
np.random.seed(7)
for s in range(steps):
    engine.LBFGS_EVENTS = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = opt._run_batch(V, range(1, 7))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ms = [a.elapsed_time(b) for _, a, b in engine.LBFGS_EVENTS]
    mx = [int(r["iters"].max().item()) for r in res["per_k"]]
    print(f"step {s}: {dt*1e3:.1f} ms  kernels={['%.1f' % m for m in ms]} max_iters={mx}", flush=True)

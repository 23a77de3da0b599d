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
orig = engine.lbfgs_solve
log = []
def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(*a, **k)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    log.append((a[0].k, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    return r
engine.lbfgs_solve = timed
import slam_decomposition_b200.optimizer as om
for s in range(steps):
    log.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt._run_batch(V, range(1, 7))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"step {s}: {dt:.0f} ms; per k (launch-call ms, wait ms): " + " ".join(f"k{k}:{a:.0f}/{b:.0f}" for k, a, b in log), flush=True)

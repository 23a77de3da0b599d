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
Nt = int(sys.argv[1])
V = torch.as_tensor(bench.haar_targets(Nt, 42), device=dev)
for m in (0, 3, 4, 5, 6, 8):
    basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6)
    opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=16)
    opts = engine.opt_defaults(); opts.history = m
    np.random.seed(3)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = opt._run_batch(V, range(1, 7), opts)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"history={m}: {min(ts)*1e3:.0f} ms  evals={opt.last_stats['evals']:.3e}  solved={(res['best_loss']<=1e-10).mean():.4f} k-hist={np.bincount(res['best_k'], minlength=7)[1:].tolist()}", flush=True)

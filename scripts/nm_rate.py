"""K5b (batched Nelder-Mead) rate: 131072 x 4 Makhlin-functional problems on the sqrt(iSWAP) k = 3 template."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import _lib, engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.utils.gates.custom_gates import RiSwapGate
dev = engine.require_cuda()
basis = CircuitTemplate(base_gates=[RiSwapGate(1 / 2)], maximum_span_guess=3, preseed=False)
basis.build(3)
V = torch.as_tensor(bench.haar_targets(131072, 9), device=dev)
nm = engine.nm_defaults()
nm.cost_kind = _lib.COST_MAKHLIN_FUNCTIONAL
ev = torch.zeros(1, dtype=torch.int64, device=dev)
for rep in range(3):
    ev.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
    loss, x, it = engine.nm_solve(basis.desc, V, 4, nm, seed=3, evals=ev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{dt * 1e3:8.1f} ms  {int(ev.item()) / dt / 1e6:7.1f} M objective evaluations/s  solved<=1e-8 {(loss.min(dim=1).values <= 1e-8).float().mean().item():.4f}", flush=True)

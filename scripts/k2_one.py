"""One launch each of the streaming K2 kernel (loss + gradient) at k = 3 and k = 6, sqCNOT templates, B = 2^22 (for ncu)."""
import os, sys
sys.path.insert(0, os.getcwd())
import math
import torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate
dev = engine.require_cuda()
V = torch.as_tensor(bench.haar_targets(4096, 5678), device=dev)
for k in (3, 6):
    basis = CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=k, preseed=False)
    basis.build(k)
    B = 1 << 22
    X = torch.rand((B, basis.desc.n_params), device=dev, dtype=torch.float64) * (2 * math.pi)
    engine.loss_grad(basis.desc, X, V)
torch.cuda.synchronize()
print("done")

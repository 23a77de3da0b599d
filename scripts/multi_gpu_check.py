"""Run under torchrun on N GPUs: the sharded coverage sweep must reproduce the single-rank histogram bit for bit
(counter-based Philox stream + int64 all-reduce over NCCL), and the gathered decomposition table must hold every
rank's shard.  Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import distributed as D, engine
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate

rank, world, local = D.init_from_env()
dev = engine.require_cuda()
n = 20_000_000
basis = pdv.smush_template(np.pi / 4, np.pi / 4, 0.5, 3)   # sqCNOT, k = 3 (configs[4] shape)
torch.cuda.synchronize(); D.barrier(); t0 = time.perf_counter()
hist = pdv.coverage_sweep(basis, n, seed=2023)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
ok_cov = True
if rank == 0:
    ref = pdv.coverage_histogram(basis, n, seed=2023)
    ok_cov = bool(torch.equal(ref, hist)) and int(hist.sum()) == n
# decomposition table gather
Nt = 4096
V = torch.as_tensor(bench.haar_targets(Nt, 100 + rank), device=dev)
opt = TemplateOptimizer(CircuitTemplate(base_gates=[ConversionGainGate(*bench.SQCNOT)], maximum_span_guess=6), BasicCost(),
                        override_fail=True, training_restarts=16)
np.random.seed(rank)
res = opt._run_batch(V, range(1, 7))
tab = D.allgather_table({"loss": res["best_loss_dev"], "k": res["best_k_dev"], "x": res["best_x"]})
ok_tab = tab["loss"].shape[0] == world * Nt and bool((tab["loss"] <= 1e-10).all())
mine = tab["loss"][rank * Nt:(rank + 1) * Nt]
ok_tab = ok_tab and bool(torch.equal(mine, res["best_loss_dev"]))
if rank == 0:
    print(f"world={world}: coverage {n:.0e} samples in {dt*1e3:.1f} ms ({n/dt/1e6:.0f} Msamples/s), sharded == single-rank: {ok_cov}; "
          f"gathered table rows={tab['loss'].shape[0]} all solved & shard-consistent: {ok_tab}", flush=True)
assert ok_cov and ok_tab
D.shutdown()

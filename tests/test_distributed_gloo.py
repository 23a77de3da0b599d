"""N>1 path on CPU: world_size-2 gloo process group exercising the sharding + the two collectives of the hot path
(histogram all-reduce for the coverage sweep, result-table all-gather for the decomposition sweep).  The per-rank
"compute" here is the oracle's coverage stream so that the sharded result can be checked against the single-rank one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import oracle as O
    from slam_decomposition_b200 import distributed as D

    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and D.world() == (rank, world)
    tmpl = O.OracleTemplate("cg", (0.0, 0.0, np.pi / 2, 0.0, 0.5), k=2, no_exterior_1q=True)
    lo, hi = D.shard_range(n, r, w)
    hist = torch.as_tensor(O.coverage_histogram(tmpl, 2023, lo, hi - lo, 0.0, 2 * np.pi, nbins=8))
    D.allreduce_histogram(hist)
    # result-table gather: every rank contributes (hi-lo) rows padded to the common shard size
    rows = n // w
    tab = {"loss": torch.full((rows,), float(rank), dtype=torch.float64), "k": torch.full((rows,), rank + 1, dtype=torch.int32),
           "x": torch.arange(rows * 3, dtype=torch.float64).reshape(rows, 3) + 1000 * rank}
    full = D.allgather_table(tab)
    # asynchronous form (bench.py overlaps the gather with the next sweep) and ragged shards padded to a common length
    full_async, handles = D.allgather_table(tab, async_op=True)
    D.wait_all(handles)
    assert all(torch.equal(full_async[k], full[k]) for k in tab)
    ragged = {"loss": torch.full((rows - rank,), float(rank), dtype=torch.float64)}
    padded = D.allgather_table(ragged, pad_to=rows)
    assert padded["loss"].shape == (w * rows,) and padded["loss"][rows - 1] == 0.0 and padded["loss"][rows] == 1.0
    assert padded["loss"][2 * rows - 1] == 0.0  # rank 1 contributed rows - 1 rows: its last row is padding
    t = D.max_over_ranks(float(rank + 1), device="cpu")
    s = D.sum_over_ranks(float(rank + 1), device="cpu")
    D.barrier()
    if rank == 0:
        q.put((hist.numpy(), {k: v.numpy() for k, v in full.items()}, t, s))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_sharded_coverage_and_table_gather():
    import oracle as O

    world, n = 2, 600
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    hist, full, t, s = q.get(timeout=150)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    tmpl = O.OracleTemplate("cg", (0.0, 0.0, np.pi / 2, 0.0, 0.5), k=2, no_exterior_1q=True)
    ref = O.coverage_histogram(tmpl, 2023, 0, n, 0.0, 2 * np.pi, nbins=8)
    assert np.array_equal(hist, ref) and hist.sum() == n
    rows = n // world
    assert full["loss"].shape == (n,) and np.array_equal(full["loss"][:rows], np.zeros(rows)) and np.all(full["loss"][rows:] == 1)
    assert np.array_equal(full["k"], np.repeat([1, 2], rows))
    assert full["x"].shape == (n, 3) and full["x"][rows, 0] == 1000
    assert (t, s) == (2.0, 3.0)

"""Multi-GPU plumbing: one process per GPU, work sharded by contiguous index range, one collective at the end.

The reference has no parallelism at all.  The hot path shards trivially -- (target, restart)
problems and Monte-Carlo samples are independent and the RNG is counter-based -- so the only inter-GPU
step is the gather of the per-target result table / the sum of coverage histograms, done
with ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank).  Initialises the default process group when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of range(n) for `rank`: [r*n/G, (r+1)*n/G)."""
    return (rank * n) // world_size, ((rank + 1) * n) // world_size


def allreduce_histogram(hist: torch.Tensor) -> torch.Tensor:
    """Sum of per-rank coverage histograms (int64), in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def allgather_table(table: Dict[str, torch.Tensor], async_op: bool = False, pad_to: int | None = None):
    """Concatenate per-rank result tables along dim 0.  Every tensor must have the same dim-0 length on all ranks (equal
    shards in the weak-scaling sweep); ragged shards pass `pad_to` = the largest shard length and get zero-padded rows.
    With `async_op` the collectives are only issued: returns (tables, handles) and the caller calls `wait_all(handles)` before
    it reads the tables (the gather then overlaps whatever is enqueued next)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return (table, []) if async_op else table
    out = {}
    handles = []
    ws = dist.get_world_size()
    for key, t in table.items():
        t = t.contiguous()
        if pad_to is not None and t.shape[0] < pad_to:
            pad = torch.zeros((pad_to - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad])
        full = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        h = dist.all_gather_into_tensor(full, t, async_op=async_op)
        if async_op:
            handles.append(h)
        out[key] = full
    return (out, handles) if async_op else out


def wait_all(handles) -> None:
    for h in handles:
        h.wait()


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shutdown():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()

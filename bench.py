#!/usr/bin/env python
"""bench.py -- headline benchmark of the SLAM hot path on B200 (contract: see the task statement / DESIGN.md).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
  batched Haar-target decomposition sweep -- `--targets` Haar-random 2Q unitaries per GPU x 16 restarts onto
  sqCNOT = ConversionGainGate(0, 0, pi/4, pi/4, 1/2) templates, k = 1..6 ascending with early exit at loss < 1e-10,
  BasicCost, analytic-gradient L-BFGS on the device.  One "step" = one full sweep over the batch.

metric  = template evaluations (loss + gradient, fp64) per second, whole job (all ranks)
value   = with the targets already resident in HBM
e2e     = the same through TemplateOptimizer.approximate_targets() with HOST buffers (H2D of the targets and
          D2H of the result table inside the timed region)
extras  = haar_decompositions_per_sec, roofline (FP64 DFMA peak measured live; algorithmic AND executed-FLOP fractions),
          cpu_baseline (oracle port on the host cores: evals/s and whole decompositions/s), kernel micro-benchmarks, the other
          BASELINE configs ([1] coverage Monte-Carlo table, [3] smush training grid, [4] 1e9-sample coverage sweep sharded
          over the GPUs with its histogram all-reduce), and at N > 1 a strong-scaling sweep of the same 1e5 targets.

`--impl reference` times the reference's CPU path (the numpy/scipy oracle port of optimizer.py:188-313: scipy BFGS
with finite-difference gradients, one Python call per template evaluation) on all host cores, same metric/unit.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SQCNOT = (0.0, 0.0, math.pi / 4, math.pi / 4, 0.5)
K_MAX = 6
RESTARTS = 16

# ---- numbers taken from committed ncu captures (profiles/r02_executed_flops.txt lists the reports and the arithmetic) --------
# dram bytes (read + write) of one lbfgs_kernel launch (k = 3, 1e5 targets x 16 restarts)
NCU_DRAM_BYTES_K3_LAUNCH = 61_125_376 + 302_178_304
# FP64 pipe activity of that launch (sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active)
NCU_PIPE_FP64_ACTIVE_K5_K3 = 0.624
# EXECUTED FP64 FLOP per unit = (2 dfma + dmul + dadd thread instructions, smsp__sass_thread_inst_executed_op_*_pred_on) of
# one launch / the units that launch processed (scripts/executed_flops.py -> profiles/r02_executed_flops.txt).  The
# algorithmic model below credits dense 4x4 products the kernels do not execute (Kronecker / block structure, factored U3),
# so `frac` (algorithmic) overstates pipe use; `frac_executed` does not.
NCU_EXEC_FLOP = {
    "k5_eval_k3": 9300.5,          # lbfgs_kernel, k = 3 launch of the sweep: per loss+grad evaluation incl. the L-BFGS bookkeeping
    "k2_lossgrad_k3": 7834.0,      # loss_grad_kernel<2 lanes>, sqCNOT k = 3, per row
    "k2_lossgrad_k6": 13966.0,     # loss_grad_kernel<4 lanes>, sqCNOT k = 6, per row
    "k3_weyl": 3597.5,             # weyl_kernel, per Haar matrix
    "k4b_traj_point": 6260.1,      # trajectory_kernel, per trajectory point (one slice exponential + prefix product + c1c2c3)
    "k6_plain_sqcnot_k3": 5262.2,  # coverage_kernel, plain template, per sample
    "k6_smush_sqcnot_k3": 27853.6, # coverage_kernel, parallel-drive template (6 slice exponentials), per sample
    "k2_smush_lossgrad": 65115.8,  # smush_loss_grad_kernel<grad, eigen-form>, sqrt(iSWAP) k = 3 T = 2 (P = 30), per row
    "k2_smush_loss": 24587.4,      # smush_loss_grad_kernel<loss only>, per row
}


# ----------------------------------------------------------------------------------------------------
# ALGORITHMIC FLOP model: real add/mul = 1, FMA = 2, complex FMA = 8, dense 4x4 complex matmul = 512
# ----------------------------------------------------------------------------------------------------
def F_eval(k: int) -> int:
    """loss only: 2k matmuls + (k+1) layers (U3 build + kron: 124) + trace (128)."""
    return 512 * 2 * k + 124 * (k + 1) + 128


def F_lossgrad(k: int) -> int:
    """reversible adjoint: per layer 3 matmuls + 6 partial-trace contractions of 128; per 2Q gate 2 matmuls."""
    return F_eval(k) + 512 * (3 * (k + 1) + 2 * k) + 768 * (k + 1)


# ----------------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------------
def haar_targets(n: int, seed: int) -> np.ndarray:
    """n Haar-random U(4) (Ginibre -> QR -> phase fix; the algorithm of scipy.stats.unitary_group, vectorised)."""
    rng = np.random.default_rng(seed)
    z = (rng.standard_normal((n, 4, 4)) + 1j * rng.standard_normal((n, 4, 4))) / math.sqrt(2.0)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r, axis1=-2, axis2=-1)
    return np.ascontiguousarray(q * (d / np.abs(d))[:, None, :])


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference loop, bounded sample, all cores
# ----------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, budget_s = args
    import oracle as O

    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    lossgrad = 0.0
    nfev = 0
    restarts = 0
    k = 1
    while time.perf_counter() - t0 < budget_s:
        V = O.haar_unitary(rng)
        tmpl = O.OracleTemplate("cg", SQCNOT, k=k)
        res = O.literal_run(lambda kk: tmpl, V, range(k, k + 1), restarts=1, kind="basic", rng=rng)
        nfev += res.nfev
        lossgrad += res.nfev / (tmpl.n_params + 1)  # one FD gradient = P+1 template evaluations (scipy 2-point)
        restarts += 1
        k = k % K_MAX + 1
    return lossgrad, nfev, restarts, time.perf_counter() - t0


def cpu_sample(budget_s: float, cores: int):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        out = pool.map(_cpu_worker, [(1000 + i, budget_s) for i in range(cores)])
    wall = time.perf_counter() - t0
    lossgrad = sum(o[0] for o in out)
    nfev = sum(o[1] for o in out)
    restarts = sum(o[2] for o in out)
    busy = max(o[3] for o in out)
    return {"lossgrad_per_s": lossgrad / busy, "template_evals_per_s": nfev / busy, "restarts": restarts,
            "busy_s": busy, "wall_s": wall}


def _cpu_decomp_worker(args):
    """Full reference flow for a few targets: k = 1..6 x 16 restarts x scipy BFGS(jac=None) with the reference's early
    exits (optimizer.py:188-313, restated in oracle.literal_run)."""
    idx, targets, seed = args
    import oracle as O

    out = []
    for i, V in zip(idx, targets):
        rng = np.random.default_rng(seed + int(i))
        t0 = time.perf_counter()
        res = O.literal_run(lambda k: O.OracleTemplate("cg", SQCNOT, k=k), V, range(1, K_MAX + 1), restarts=RESTARTS,
                            kind="basic", rng=rng)
        out.append((int(i), float(res.best_result), int(res.best_cycles), int(res.nfev), time.perf_counter() - t0))
    return out


def cpu_decompositions(n_targets: int, cores: int):
    """Haar decompositions/s of the CPU port: the first `n_targets` targets of the GPU arm's own stream (seed 42), each run to
    completion, spread over all cores."""
    import multiprocessing as mp

    V = haar_targets(n_targets, 42)
    chunks = [(list(range(c, n_targets, cores)), V[c::cores], 4242) for c in range(min(cores, n_targets))]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(len(chunks)) as pool:
        res = [r for part in pool.map(_cpu_decomp_worker, chunks) for r in part]
    wall = time.perf_counter() - t0
    loss = np.array([r[1] for r in res])
    return {"haar_decompositions_per_sec": n_targets / wall, "targets": n_targets, "wall_s": wall, "cores": len(chunks),
            "solved_fraction_1e-10": float((loss <= 1e-10).mean()), "solved_fraction_1e-9": float((loss <= 1e-9).mean()),
            "mean_nfev_per_target": float(np.mean([r[3] for r in res])), "mean_cycles": float(np.mean([r[2] for r in res])),
            "mean_seconds_per_target_per_core": float(np.mean([r[4] for r in res])),
            "what": (f"oracle.literal_run to completion (k = 1..{K_MAX}, {RESTARTS} restarts, scipy BFGS with finite-difference "
                     "gradients, early exit at 1e-10) on the first targets of the GPU arm's Haar stream (seed 42)")}


def vectorised_numpy_rate(budget_s: float = 2.0, batch: int = 4096):
    """the oracle vectorised over the batch axis (one core) -- an upper bound on what numpy can do for the
    forward evaluation + BasicCost; template evaluations per second, cycling k = 1..6 like the literal loop."""
    import oracle as O

    rng = np.random.default_rng(5)
    V = haar_targets(batch, 9)
    t0 = time.perf_counter()
    n = 0
    k = 1
    while time.perf_counter() - t0 < budget_s:
        tmpl = O.OracleTemplate("cg", SQCNOT, k=k)
        X = rng.uniform(0, 2 * math.pi, (batch, tmpl.n_params))
        U = tmpl.eval_batch(X)
        T = np.einsum("bij,bij->b", np.conj(V), U)
        _ = 1.0 - np.abs(T) / 4.0
        n += batch
        k = k % K_MAX + 1
    return n / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    budget = args.cpu_seconds
    vals = []
    for _ in range(args.warmup):
        cpu_sample(min(2.0, budget), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_sample(budget, cores))
    total = time.perf_counter() - t0
    v = float(np.mean([s["lossgrad_per_s"] for s in vals]))
    decomp = cpu_decompositions(args.cpu_targets, cores) if args.cpu_targets > 0 else None
    sample = (f"{cores} processes x {budget:.0f} s of scipy BFGS restarts (finite-difference gradients, one numpy "
              f"template evaluation per call) on sqCNOT templates cycling k=1..{K_MAX}, Haar targets; "
              "loss+grad evaluation = (P+1) function evaluations")
    line = {
        "impl": "reference", "metric": "template_evals_per_sec", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "raw_template_evals_per_s": float(np.mean([s["template_evals_per_s"] for s in vals])),
                         "vectorised_numpy_template_evals_per_s_1core": vectorised_numpy_rate(),
                         "decompositions": decomp},
        "haar_decompositions_per_sec": decomp["haar_decompositions_per_sec"] if decomp else None,
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.fh.close()
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                       "samples": len(sm)}
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


def workload_config(args):
    return {"workload": f"haar_sweep_sqCNOT: {args.targets} Haar targets/GPU x {RESTARTS} restarts, "
                        f"ConversionGainGate(0,0,pi/4,pi/4,1/2) templates k=1..{K_MAX}, BasicCost, early exit at 1e-10",
            "targets_per_gpu": args.targets, "restarts": RESTARTS, "k_max": K_MAX,
            "baseline_config": "configs[2] (batched Haar-target decomposition sweep)",
            "l2": "L2 flushed (256 MiB write) between timed steps"}


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from slam_decomposition_b200 import distributed as D
    from slam_decomposition_b200 import engine
    from slam_decomposition_b200.basis import CircuitTemplate
    from slam_decomposition_b200.cost_function import BasicCost
    from slam_decomposition_b200.optimizer import TemplateOptimizer
    from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate

    rank, world, local = D.init_from_env()
    dev = engine.require_cuda()
    Nt = args.targets
    np.random.seed(1234 + rank)

    def make_opt():
        basis = CircuitTemplate(base_gates=[ConversionGainGate(*SQCNOT)], maximum_span_guess=K_MAX, preseed=False)
        return TemplateOptimizer(basis=basis, objective=BasicCost(), use_callback=False, override_fail=True,
                                 training_restarts=RESTARTS)

    opt = make_opt()
    V_host = torch.as_tensor(haar_targets(Nt, seed=42 + rank)).pin_memory()
    V_dev = V_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    k_range = range(1, K_MAX + 1)

    # The only inter-GPU step of the sweep: gather of the per-target result table.  It is issued asynchronously
    # (NCCL runs it on its own stream) and waited for one step later, so it overlaps the next sweep; the result tensors of a
    # sweep are fresh allocations, so nothing the gather reads is overwritten meanwhile.
    pending = []

    def finish_gather():
        while pending:
            handles, keep = pending.pop(0)
            D.wait_all(handles)

    def step_resident():
        res = opt._run_batch(V_dev, k_range)
        finish_gather()  # the previous step's gather has had a whole sweep to complete
        src = {"loss": res["best_loss_dev"], "k": res["best_k_dev"], "x": res["best_x"]}
        tab, handles = D.allgather_table(src, async_op=True)
        pending.append((handles, (tab, src)))  # (the sources stay referenced until the gather has completed)
        return res

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
        flush.fill_(1.0)
    finish_gather()
    torch.cuda.synchronize()

    # ---- timed: resident inputs ------------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    engine.LBFGS_EVENTS = []
    engine.LBFGS_SPANS = []
    opt.launch_evals = []
    launches0 = engine.LAUNCHES
    evals_total = 0
    solved = 0
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_resident()
        evals_total += opt.last_stats["evals"]
        solved += int((res["best_loss"] <= opt.success_threshold).sum())
        flush.fill_(1.0)
    finish_gather()
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    t_local = e0.elapsed_time(e1) * 1e-3
    launches = engine.LAUNCHES - launches0
    events = engine.LBFGS_EVENTS
    spans = engine.LBFGS_SPANS
    engine.LBFGS_EVENTS = None
    engine.LBFGS_SPANS = None
    launch_evals = list(opt.launch_evals)
    clock_info = clocks.stop() if rank == 0 else None
    t = D.max_over_ranks(t_local, dev)
    evals_all = D.sum_over_ranks(float(evals_total), dev)
    solved_all = D.sum_over_ranks(float(solved), dev)

    # the collective alone (it is overlapped in the timed region above): three synchronous gathers of the last result table
    coll_ms = 0.0
    if world > 1:
        src = {"loss": res["best_loss_dev"], "k": res["best_k_dev"], "x": res["best_x"]}
        D.allgather_table(src)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            D.allgather_table(src)
        c1.record()
        torch.cuda.synchronize()
        coll_ms = D.max_over_ranks(c0.elapsed_time(c1) / 3, dev)

    # ---- roofline of the dominant kernel (lbfgs_kernel), per-launch CUDA events from the timed region ---
    # The launches of consecutive template sizes are chained on two streams and overlap (the next size fills the SMs the
    # draining one frees), so the kernel time of a sweep is the span from the first launch to the end of the last one, not
    # the sum of the per-launch durations (which count the overlap twice; they are reported per k for the shares only).
    if spans:
        kern_ms = sum(a.elapsed_time(b) for a, b in spans)
    else:
        kern_ms = sum(a.elapsed_time(b) for _, a, b in events)
    alg_flops = sum(n * F_lossgrad(k) for k, n in launch_evals)
    per_k = {}
    per_sweep = len(events) // len(spans) if spans else 0
    for idx, ((k, a, b), (_, n)) in enumerate(zip(events, launch_evals)):
        d = per_k.setdefault(k, {"ms": 0.0, "evals": 0, "launches": 0})
        if spans:
            # chained launches: a launch is charged the time between the previous launch's end (the span start for the first
            # of a sweep) and its own end -- its start event fires while it still waits for SMs
            prev_end = spans[idx // per_sweep][0] if idx % per_sweep == 0 else events[idx - 1][2]
            d["ms"] += max(prev_end.elapsed_time(b), 0.0)
        else:
            d["ms"] += a.elapsed_time(b)
        d["evals"] += n
        d["launches"] += 1

    # ---- timed: end to end through the public host-buffer API ------------------------------------------
    h2d = V_host.numel() * V_host.element_size()
    d2h = 0
    e2e_evals = 0
    for _ in range(1):
        opt.approximate_targets(V_host, k_range)  # warm the pinned path
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = opt.approximate_targets(V_host, k_range)
        e2e_evals += opt.last_stats["evals"]
        d2h = sum(v.nbytes for v in out.values())
    torch.cuda.synchronize()
    D.barrier()
    t_e2e = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_all = D.sum_over_ranks(float(e2e_evals), dev)

    # ---- strong scaling of the sweep: the SAME 1e5 targets split over the ranks (tail effects at 1e5 / N per GPU) ------
    strong = None
    if world > 1:
        lo_t, hi_t = D.shard_range(args.targets, rank, world)
        Vs = torch.as_tensor(haar_targets(args.targets, seed=42)[lo_t:hi_t]).to(dev)
        opt_s = make_opt()
        np.random.seed(99)
        opt_s._run_batch(Vs, k_range)
        D.barrier()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        ev_s = 0
        for _ in range(args.steps):
            r_s = opt_s._run_batch(Vs, k_range)
            D.allgather_table({"loss": r_s["best_loss_dev"], "k": r_s["best_k_dev"]}, pad_to=(args.targets + world - 1) // world)
            ev_s += opt_s.last_stats["evals"]
        s1.record()
        torch.cuda.synchronize()
        D.barrier()
        t_s = D.max_over_ranks(s0.elapsed_time(s1) * 1e-3, dev)
        ev_s_all = D.sum_over_ranks(float(ev_s), dev)
        strong = {"targets_total": args.targets, "targets_per_gpu": hi_t - lo_t, "ms_per_sweep": 1e3 * t_s / args.steps,
                  "evals_per_s": ev_s_all / t_s, "haar_decompositions_per_sec": args.targets * args.steps / t_s,
                  "scaling": "strong"}
        del Vs, opt_s

    # ---- BASELINE configs[4]: sqCNOT parallel-drive coverage sweep, k = 1..6, `--coverage-samples` samples per k in total,
    #      sharded over the ranks (contiguous slices of one Philox stream), one int64 128^3 all-reduce per k -----------------
    coverage = None
    if args.coverage_samples > 0:
        coverage = coverage_sweep_block(engine, D, args.coverage_samples, rank, world, dev)

    if rank != 0:
        D.shutdown()
        return 0

    # ---- rank 0: denominators, micro-benchmarks, CPU baseline, the JSON line ---------------------------
    peak_flops, _ = engine.fp64_peak(8192)
    frac = (alg_flops / (kern_ms * 1e-3)) / peak_flops if kern_ms else None
    k3 = per_k.get(3)
    line = {
        "metric": "template_evals_per_sec", "value": evals_all / t, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "clocks": clock_info,
        "e2e": {"value": e2e_all / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "haar_decompositions_per_sec": world * Nt * args.steps / t_e2e},
        "gpu_launches": launches,
        "collective_ms_per_step": coll_ms,
        "collective_note": ("all_gather of the per-target result table (36 MB per rank), timed ALONE after the timed region; inside "
                            "it the gather is issued asynchronously and waited for one sweep later, i.e. overlapped with the next sweep"),
        "haar_decompositions_per_sec": world * Nt * args.steps / t,
        "solved_fraction": solved_all / (world * Nt * args.steps),
        "evals_per_step": evals_all / args.steps,
        "roofline": {
            "bound": "fp64", "kernel": "slam::lbfgs_kernel (K5: loss+grad adjoint + L-BFGS, state in shared memory)",
            "achieved": alg_flops / (kern_ms * 1e-3) / 1e12 if kern_ms else None, "peak": peak_flops / 1e12,
            "unit": "TFLOP/s", "frac": frac,
            "frac_note": ("ALGORITHMIC flops (SURVEY 8d model, dense 4x4 products credited) over the measured DFMA peak; it is not "
                          "pipe utilisation -- see frac_executed / pipe_fp64_active"),
            "frac_executed_k3": ((k3["evals"] * NCU_EXEC_FLOP["k5_eval_k3"] / (k3["ms"] * 1e-3)) / peak_flops) if k3 and k3["ms"] > 0 else None,
            "executed_flop_per_eval_k3": NCU_EXEC_FLOP["k5_eval_k3"],
            "pipe_fp64_active": NCU_PIPE_FP64_ACTIVE_K5_K3,
            "pipe_fp64_active_note": "sm__pipe_fp64_cycles_active of the k = 3 launch, ncu --set full (profiles/r02_lbfgs_k3_ncu_full.txt)",
            "traffic": NCU_DRAM_BYTES_K3_LAUNCH, "traffic_unit": "bytes/launch",
            "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of the k=3 lbfgs_kernel launch of this workload "
                             "(1e5 targets x 16 restarts) from one ncu --set full capture, profiles/r02_lbfgs_k3_ncu_full.txt; "
                             "algorithmic bytes of that launch = result table 1.6e6 x (24+1) x 8 B + iters 6.4 MB = 326 MB "
                             "+ targets 25.6 MB: the kernel is compute bound, DRAM throughput 0.1 % of peak"),
            "peak_source": "slam_fp64_peak: register-resident DFMA loop measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
            "flop_model": "SURVEY 8(d): F_lossgrad(k) = 1024k + 124(k+1) + 128 + 512(5k+3) + 768(k+1) per evaluation",
            "kernel_share_of_step": kern_ms * 1e-3 / t_local if t_local else None,
            "timing": ("span of the chained lbfgs_kernel launches (k = 1..6 on two streams), CUDA events on the launching stream"
                       if spans else "sum of per-launch CUDA-event durations"),
            "per_k": {str(k): {"ms_per_launch": v["ms"] / v["launches"], "evals_per_launch": v["evals"] / v["launches"],
                               "tflops": (v["evals"] * F_lossgrad(k) / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else None}
                      for k, v in sorted(per_k.items())},
        },
    }
    if strong is not None:
        line["strong_scaling"] = strong
    if coverage is not None:
        line["coverage_1e9"] = coverage
    if world == 1 and not args.no_micro:
        line["kernels"] = micro_benchmarks(engine, peak_flops)
        line["configs"] = config_workloads(engine)
    if world == 1 and args.cpu_seconds > 0:
        cores = os.cpu_count() or 1
        s = cpu_sample(args.cpu_seconds, cores)
        line["cpu_baseline"] = {
            "value": s["lossgrad_per_s"], "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": (f"{cores} processes x {args.cpu_seconds:.0f} s of scipy BFGS restarts (finite-difference gradients) on "
                       f"sqCNOT templates cycling k=1..{K_MAX}; one loss+grad evaluation = (P+1) numpy template evaluations"),
            "raw_template_evals_per_s": s["template_evals_per_s"], "restarts_completed": s["restarts"],
            # forward evaluation + BasicCost vectorised over the batch axis, one core (loss only)
            "vectorised_numpy_template_evals_per_s_1core": vectorised_numpy_rate(),
            # the second half of the BASELINE metric: whole decompositions per second of the same CPU port
            "decompositions": cpu_decompositions(args.cpu_targets, cores) if args.cpu_targets > 0 else None}
    print(json.dumps(line), flush=True)
    D.shutdown()
    return 0


def coverage_sweep_block(engine, D, n_total, rank, world, dev):
    """BASELINE configs[4] (8xB200 coverage-set sweep sqCNOT k = 1..6, 1e9 samples sharded, histogram all-reduce over NVLink):
    every rank bins its contiguous slice of the SAME Philox stream, one all-reduce(sum) of the int64 128^3 histogram per k.
    The histogram is a pure function of (seed, sample index), so its checksum must be identical at every GPU count -- a free
    multi-GPU parity check (bench callers compare `hist_checksum` across N)."""
    import torch

    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

    nbins = 128
    weights = (torch.arange(nbins ** 3, dtype=torch.int64, device=dev) * 2654435761 + 12345) % 1000003
    lo, hi = D.shard_range(n_total, rank, world)
    per_k = {}
    t_kernel = t_coll = 0.0
    checksum = 0
    for k in range(1, K_MAX + 1):
        basis = pdv.smush_template(math.pi / 4, math.pi / 4, 0.5, k)
        if k == 1:
            pdv.coverage_histogram(basis, 100_000, seed=1)  # warm-up
        D.barrier()
        torch.cuda.synchronize()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        hist = pdv.coverage_histogram(basis, hi - lo, seed=2023, nbins=nbins, first_sample=lo)
        b.record()
        hist = D.allreduce_histogram(hist)
        c.record()
        torch.cuda.synchronize()
        tk = D.max_over_ranks(a.elapsed_time(b) * 1e-3, dev)
        tc = D.max_over_ranks(b.elapsed_time(c) * 1e-3, dev)
        t_kernel += tk
        t_coll += tc
        cs = int((hist * weights).sum().item())
        total = int(hist.sum().item())
        assert total == n_total, (k, total, n_total)
        checksum = (checksum * 1000003 + cs) % (1 << 61)
        per_k[str(k)] = {"params": basis.desc.n_params, "samples_per_s": n_total / (tk + tc), "kernel_ms": 1e3 * tk,
                         "allreduce_ms": 1e3 * tc, "hist_checksum": cs,
                         "haar_volume_voxels": pdv.haar_volume_fraction(hist, nbins) if rank == 0 else None}
    return {"workload": f"sqCNOT parallel-drive templates k=1..{K_MAX}, {n_total} samples per k in total, sharded over {world} GPU(s), "
                        "128^3 int64 histogram, one all-reduce per k", "samples_per_k": n_total,
            "samples_per_s": K_MAX * n_total / (t_kernel + t_coll), "kernel_s": t_kernel, "allreduce_ms": 1e3 * t_coll,
            "hist_checksum": checksum, "scaling": "strong", "per_k": per_k}


def _timed(fn, reps=3, warm=2):
    import torch

    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


def config_workloads(engine):
    """The BASELINE configs that are not the bench workload, as specified, driver-visible:
    configs[1]  coverage-set Monte-Carlo for the sqiSwap / CNOT / B bases, k = 1..3, 1e7 samples each, plain and parallel-drive
                templates -> binned Weyl-chamber points (samples/s each);
    configs[3]  parallel-drive speed-limit sweep: the iSwap-smush training grid (durations t = 0.25 .. 1.5, T = 1 .. 6 time
                slices; flow of the reference's scripts/local_smush_test.ipynb, batched): targets/s and solved fraction."""
    import torch

    from slam_decomposition_b200.basisv2 import CircuitTemplateV2
    from slam_decomposition_b200.cost_function import BasicCost
    from slam_decomposition_b200.optimizer import TemplateOptimizer
    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
    from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainSmushGate

    dev = engine.require_cuda()
    out = {}
    n = 10_000_000
    cov = {}
    for gc, gg, t, name, _ in pdv.GATE_LIST:
        if name not in ("sqiSwap", "CNOT", "B"):
            continue
        for k in (1, 2, 3):
            for kind, make in (("plain", pdv.plain_template), ("smush", pdv.smush_template)):
                basis = make(gc, gg, t, k)
                hist = torch.zeros(128 ** 3, dtype=torch.int64, device=dev)
                dt = _timed(lambda: pdv.coverage_histogram(basis, n, seed=2023, hist=hist), reps=1, warm=1)
                cov[f"{name}_k{k}_{kind}"] = {"samples_per_s": n / dt, "params": basis.desc.n_params,
                                              "occupied_voxels": int((hist > 0).sum().item())}
    out["configs[1]_coverage_1e7"] = {"samples": n, "cases": cov,
                                      "total_samples_per_s": len(cov) * n / sum(n / c["samples_per_s"] for c in cov.values())}

    # configs[3]
    rng = np.random.default_rng(0)
    np.random.seed(0)
    n_t, restarts = 4096, 8
    grid = {}
    total_s = 0.0
    solved_inst = []
    for t_el in (0.25, 0.5, 0.75, 1.0, 1.25, 1.5):
        T = round(t_el / 0.25)

        def pp2(*vargs, T=T, t_el=t_el):
            return ConversionGainSmushGate(0, 0, math.pi / 2, 0, vargs[:T], vargs[T:], t_el=t_el)

        basis = CircuitTemplateV2(n_qubits=2, base_gates=[pp2], edge_params=[[(0, 1)]], vz_only=False, param_vec_expand=[0, T, T])
        basis.build(1)
        basis.spanning_range = range(1, 2)
        for el in basis.circuit.parameters:
            if "Q" in str(el):
                basis.add_bound(str(el), 2 * math.pi, -2 * math.pi)
        P = basis.desc.n_params
        lo, hi = basis.x0_bound_arrays()
        half = n_t // 2
        own = engine.template_eval(basis.desc, torch.as_tensor(rng.uniform(lo, hi, (half, P)), device=dev)).cpu().numpy()
        V = np.concatenate([own, haar_targets(n_t - half, 7)])
        opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=restarts)
        opt.approximate_targets(V[:64], range(1, 2))  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = opt.approximate_targets(V, range(1, 2))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        total_s += dt
        ok = res["loss"] <= 1e-9
        solved_inst.append(float(ok[:half].mean()))
        grid[f"t={t_el}"] = {"slices": T, "params": P, "ms": 1e3 * dt, "loss_grad_evals_per_s": opt.last_stats["evals"] / dt,
                             "solved_template_instances": float(ok[:half].mean()), "solved_haar": float(ok[half:].mean())}
    out["configs[3]_smush_training_grid"] = {
        "targets_per_duration": n_t, "restarts": restarts, "targets_per_s": 6 * n_t / total_s, "grid_seconds": total_s,
        "solved_fraction_template_instances": float(np.mean(solved_inst)), "per_duration": grid,
        "what": ("iSwap-smush template (1Q layer, ConversionGainSmushGate(0,0,pi/2,0,gx[T],gy[T],t), 1Q layer), amplitudes bounded to "
                 "+-2pi, half of the targets are instances of the template (reachable), half Haar; K5c with the adjoint gradient "
                 "through TemplateOptimizer.approximate_targets (host buffers in and out)")}
    return out


def micro_benchmarks(engine, peak_flops):
    """Streaming kernels timed alone (CUDA events, warm-ups, inputs > L2 or generated on chip).  For every kernel:
    `frac_of_fp64_peak` = ALGORITHMIC flops where the model defines them, and `frac_executed` = EXECUTED FP64 flops
    per unit from the committed ncu captures (NCU_EXEC_FLOP) x the measured rate / the DFMA peak measured in this run."""
    import torch

    from slam_decomposition_b200.circuit import Parameter, TemplateCircuit, lower
    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
    from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainGate

    out = {}
    dev = engine.require_cuda()
    V = torch.as_tensor(haar_targets(4096, 5678), device=dev)
    for k in (1, 3, 6):
        qc = TemplateCircuit(2)
        p = 0
        for i in range(k + 1):
            for q in (0, 1):
                qc.u(*[Parameter(f"P{p + j}") for j in range(3)], q)
                p += 3
            if i < k:
                qc.append(ConversionGainGate(*SQCNOT), (0, 1))
        desc, names, _ = lower(qc)
        B = 1 << 22
        g = torch.Generator(device=dev).manual_seed(1234)
        X = torch.rand((B, desc.n_params), device=dev, dtype=torch.float64, generator=g) * (2 * math.pi)
        loss = torch.empty(B, device=dev, dtype=torch.float64)
        grad = torch.empty_like(X)
        t = _timed(lambda: engine.loss_grad(desc, X, V, out_loss=loss, out_grad=grad), reps=5, warm=3)
        ex = NCU_EXEC_FLOP.get(f"k2_lossgrad_k{k}")
        out[f"loss_grad_k{k}"] = {"evals_per_s": B / t, "tflops_alg": B / t * F_lossgrad(k) / 1e12,
                                  "frac_of_fp64_peak": B / t * F_lossgrad(k) / peak_flops,
                                  "frac_executed": (B / t * ex / peak_flops) if ex else None,
                                  "hbm_gbs": B * (16 * desc.n_params + 8) / t / 1e9, "batch": B}
        del X, grad, loss
    U = torch.as_tensor(haar_targets(1 << 21, 99), device=dev)
    t = _timed(lambda: engine.weyl(U), reps=5, warm=3)
    out["weyl_c1c2c3"] = {"matrices_per_s": U.shape[0] / t, "hbm_gbs": U.shape[0] * (256 + 24) / t / 1e9,
                          "frac_executed": U.shape[0] / t * NCU_EXEC_FLOP["k3_weyl"] / peak_flops,
                          "executed_flop_per_matrix": NCU_EXEC_FLOP["k3_weyl"]}
    # K6 coverage Monte-Carlo on the template the executed-FLOP capture was taken on (sqCNOT k=3), plain and parallel-drive
    n = 10_000_000
    for label, basis, key in (("coverage_sqCNOT_k3_plain", pdv.plain_template(math.pi / 4, math.pi / 4, 0.5, 3), "k6_plain_sqcnot_k3"),
                              ("coverage_sqCNOT_k3_smush", pdv.smush_template(math.pi / 4, math.pi / 4, 0.5, 3), "k6_smush_sqcnot_k3")):
        hist = torch.zeros(128 ** 3, dtype=torch.int64, device=dev)
        t = _timed(lambda: pdv.coverage_histogram(basis, n, seed=2023, hist=hist), reps=1, warm=2)
        out[label] = {"samples_per_s": n / t, "samples": n, "params": basis.desc.n_params,
                      "frac_executed": n / t * NCU_EXEC_FLOP[key] / peak_flops, "executed_flop_per_sample": NCU_EXEC_FLOP[key]}
    # K2 on a parameter-bound smush template (parallel_drive_volume.py:175-199, sqrt(iSWAP) k=3, T=2, P=30): loss + analytic
    # adjoint gradient through the slice exponentials vs the forward evaluation a finite-difference gradient repeats P+1 times
    basis = pdv.smush_template(math.pi / 2, 0.0, 0.5, 3)
    Bs = 1 << 20
    g = torch.Generator(device=dev).manual_seed(77)
    Xs = (torch.rand((Bs, basis.desc.n_params), device=dev, dtype=torch.float64, generator=g) - 0.5) * (8 * math.pi)
    ls = torch.empty(Bs, device=dev, dtype=torch.float64)
    gs = torch.empty_like(Xs)
    for label, want, key in (("smush_k3_loss_grad_adjoint", True, "k2_smush_lossgrad"), ("smush_k3_loss_only", False, "k2_smush_loss")):
        t = _timed(lambda: engine.loss_grad(basis.desc, Xs, V, out_loss=ls, out_grad=gs if want else None, want_grad=want))
        out[label] = {"evals_per_s": Bs / t, "batch": Bs, "params": basis.desc.n_params,
                      "frac_executed": Bs / t * NCU_EXEC_FLOP[key] / peak_flops, "executed_flop_per_row": NCU_EXEC_FLOP[key]}
    # K4b parallel-drive Weyl trajectories (pd_playground.py:169-208): N = 10 slices of the 1Q-phase smush Hamiltonian, R = 5
    # sub-times each -> one expm + prefix product + c1c2c3 per trajectory point
    Bt, Nn, Rr = 1 << 18, 10, 5
    gate = (torch.rand((Bt, 8), device=dev, dtype=torch.float64, generator=g) - 0.5) * 4
    ax = (torch.rand((Bt, Nn), device=dev, dtype=torch.float64, generator=g) - 0.5) * (4 * math.pi)
    ay = (torch.rand((Bt, Nn), device=dev, dtype=torch.float64, generator=g) - 0.5) * (4 * math.pi)
    t = _timed(lambda: engine.pd_trajectory(gate, ax, ay, 0.1, R=Rr, want_final=False))
    out["pd_trajectory_N10_R5"] = {"points_per_s": Bt * Nn * Rr / t, "trajectories": Bt,
                                   "frac_executed": Bt * Nn * Rr / t * NCU_EXEC_FLOP["k4b_traj_point"] / peak_flops,
                                   "executed_flop_per_point": NCU_EXEC_FLOP["k4b_traj_point"]}
    out["smush_k3_adjoint_vs_fd_gradient"] = (out["smush_k3_loss_grad_adjoint"]["evals_per_s"] * (basis.desc.n_params + 1)
                                              / out["smush_k3_loss_only"]["evals_per_s"])
    # K5c: batched L-BFGS with the adjoint gradient on a parameter-bound smush template (sqrt(iSWAP) k=2, T=2, P=18;
    # 131072 template-instance targets x 8 restarts from the reference's U(-4pi, 4pi) start box: enough problems per thread
    # (28) that the drain of the persistent grid -- restarts that run to maxiter = 2500 -- does not dominate the launch)
    b2 = pdv.smush_template(math.pi / 2, 0.0, 0.5, 2)
    rng = np.random.default_rng(3)
    n_k5c = 131072
    Vt = engine.template_eval(b2.desc, torch.as_tensor(rng.uniform(-1.5, 1.5, (n_k5c, b2.desc.n_params)), device=dev))
    o = engine.opt_defaults()
    o.f_far = 1e-4
    o.x0_lo, o.x0_hi = -4 * math.pi, 4 * math.pi
    ev = torch.zeros(1, dtype=torch.int64, device=dev)

    def k5c():
        ev.zero_()
        engine.fd_lbfgs_solve(b2.desc, Vt, 8, o, seed=11, central="adjoint", evals=ev)

    t = _timed(k5c, reps=1, warm=1)
    out["k5c_smush_adjoint_lbfgs"] = {"loss_grad_evals_per_s": int(ev.item()) / t, "ms": 1e3 * t, "params": b2.desc.n_params,
                                      "problems": n_k5c * 8}

    # K5b: batched Nelder-Mead on the Makhlin functional (the coordinate-based costs and the pulse searches run on it)
    from slam_decomposition_b200 import _lib
    from slam_decomposition_b200.basis import CircuitTemplate as _CT
    from slam_decomposition_b200.utils.gates.custom_gates import RiSwapGate

    b3 = _CT(base_gates=[RiSwapGate(1 / 2)], maximum_span_guess=3, preseed=False)
    b3.build(3)
    Vn = torch.as_tensor(haar_targets(131072, 9), device=dev)
    nm = engine.nm_defaults()
    nm.cost_kind = _lib.COST_MAKHLIN_FUNCTIONAL

    def k5b():
        ev.zero_()
        engine.nm_solve(b3.desc, Vn, 4, nm, seed=3, evals=ev)

    t = _timed(k5b, reps=1, warm=1)
    out["k5b_nelder_mead_makhlin"] = {"objective_evals_per_s": int(ev.item()) / t, "ms": 1e3 * t, "params": b3.desc.n_params,
                                      "problems": 131072 * 4}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--targets", type=int, default=100000, help="Haar targets per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample length per process")
    ap.add_argument("--cpu-targets", type=int, default=64,
                    help="CPU-baseline Haar targets decomposed to completion over all cores (0 = skip; ~20-40 s per target per core)")
    ap.add_argument("--coverage-samples", type=float, default=1e9,
                    help="samples per k of the configs[4] coverage sweep, in total over all GPUs (0 = skip)")
    ap.add_argument("--no-micro", action="store_true")
    args = ap.parse_args()
    args.coverage_samples = int(args.coverage_samples)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

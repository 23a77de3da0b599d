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
extras  = haar_decompositions_per_sec, roofline (FP64 DFMA peak measured live), cpu_baseline (oracle port on the
          host cores), kernel micro-benchmarks.

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
# dram bytes (read + write) of one lbfgs_kernel launch (k = 3, 1e5 targets x 16 restarts) from the committed ncu capture
NCU_DRAM_BYTES_K3_LAUNCH = 60_492_032 + 300_494_080
K_MAX = 6
RESTARTS = 16


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
                         "vectorised_numpy_template_evals_per_s_1core": vectorised_numpy_rate()},
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

    import oracle as O
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

    basis = CircuitTemplate(base_gates=[ConversionGainGate(*SQCNOT)], maximum_span_guess=K_MAX, preseed=False)
    opt = TemplateOptimizer(basis=basis, objective=BasicCost(), use_callback=False, override_fail=True,
                            training_restarts=RESTARTS)
    V_host = torch.as_tensor(haar_targets(Nt, seed=42 + rank)).pin_memory()
    V_dev = V_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    k_range = range(1, K_MAX + 1)

    def gather(res_loss, res_k, res_x):
        # the only inter-GPU step: gather of the per-target result table
        return D.allgather_table({"loss": res_loss, "k": res_k, "x": res_x})

    gather_ms = []

    def step_resident():
        res = opt._run_batch(V_dev, k_range)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        tab = gather(res["best_loss_dev"], res["best_k_dev"], res["best_x"])
        g1.record()
        gather_ms.append((g0, g1))
        return res, tab

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
        flush.fill_(1.0)
    torch.cuda.synchronize()

    # ---- timed: resident inputs ------------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    engine.LBFGS_EVENTS = []
    engine.LBFGS_SPANS = []
    opt.launch_evals = []
    gather_ms.clear()
    launches0 = engine.LAUNCHES
    evals_total = 0
    solved = 0
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res, tab = step_resident()
        evals_total += opt.last_stats["evals"]
        solved += int((res["best_loss"] <= opt.success_threshold).sum())
        flush.fill_(1.0)
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

    # ---- roofline of the dominant kernel (lbfgs_kernel), per-launch CUDA events from the timed region ---
    # The launches of consecutive template sizes are chained on two streams and overlap (the next size fills the SMs the
    # draining one frees), so the kernel time of a sweep is the span from the first launch to the end of the last one, not
    # the sum of the per-launch durations (which count the overlap twice; they are reported per k for the shares only).
    if spans:
        kern_ms = sum(a.elapsed_time(b) for a, b in spans)
    else:
        kern_ms = sum(a.elapsed_time(b) for _, a, b in events)
    alg_flops = sum(n * O.F_lossgrad(k) for k, n in launch_evals)
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

    if rank != 0:
        D.shutdown()
        return 0

    # ---- rank 0: denominators, micro-benchmarks, CPU baseline, the JSON line ---------------------------
    peak_flops, _ = engine.fp64_peak(8192)
    line = {
        "metric": "template_evals_per_sec", "value": evals_all / t, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "clocks": clock_info,
        "e2e": {"value": e2e_all / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "haar_decompositions_per_sec": world * Nt * args.steps / t_e2e},
        "gpu_launches": launches,
        "collective_ms_per_step": sum(a.elapsed_time(b) for a, b in gather_ms[: args.steps]) / args.steps,
        "haar_decompositions_per_sec": world * Nt * args.steps / t,
        "solved_fraction": solved_all / (world * Nt * args.steps),
        "evals_per_step": evals_all / args.steps,
        "roofline": {
            "bound": "fp64", "kernel": "slam::lbfgs_kernel (K5: loss+grad adjoint + L-BFGS, state in shared memory)",
            "achieved": alg_flops / (kern_ms * 1e-3) / 1e12 if kern_ms else None, "peak": peak_flops / 1e12,
            "unit": "TFLOP/s", "frac": (alg_flops / (kern_ms * 1e-3)) / peak_flops if kern_ms else None,
            "traffic": NCU_DRAM_BYTES_K3_LAUNCH, "traffic_unit": "bytes/launch",
            "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of the k=3 lbfgs_kernel launch of this workload "
                             "(1e5 targets x 16 restarts) from one ncu --set full capture, profiles/r01_lbfgs_kernel_ncu_full.txt; "
                             "algorithmic bytes of that launch = result table 1.6e6 x (24+1) x 8 B + iters 6.4 MB = 326 MB "
                             "+ targets 25.6 MB: the kernel is compute bound, DRAM throughput 0.1 % of peak"),
            "peak_source": "slam_fp64_peak: register-resident DFMA loop measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
            "flop_model": "SURVEY 8(d): F_lossgrad(k) = 1024k + 124(k+1) + 128 + 512(5k+3) + 768(k+1) per evaluation",
            "kernel_share_of_step": kern_ms * 1e-3 / t_local if t_local else None,
            "timing": ("span of the chained lbfgs_kernel launches (k = 1..6 on two streams), CUDA events on the launching stream"
                       if spans else "sum of per-launch CUDA-event durations"),
            "per_k": {str(k): {"ms_per_launch": v["ms"] / v["launches"], "evals_per_launch": v["evals"] / v["launches"],
                               "tflops": (v["evals"] * O.F_lossgrad(k) / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else None}
                      for k, v in sorted(per_k.items())},
        },
    }
    if world == 1 and not args.no_micro:
        line["kernels"] = micro_benchmarks(engine, peak_flops)
    if world == 1 and args.cpu_seconds > 0:
        cores = os.cpu_count() or 1
        s = cpu_sample(args.cpu_seconds, cores)
        line["cpu_baseline"] = {
            "value": s["lossgrad_per_s"], "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": (f"{cores} processes x {args.cpu_seconds:.0f} s of scipy BFGS restarts (finite-difference gradients) on "
                       f"sqCNOT templates cycling k=1..{K_MAX}; one loss+grad evaluation = (P+1) numpy template evaluations"),
            "raw_template_evals_per_s": s["template_evals_per_s"], "restarts_completed": s["restarts"],
            # forward evaluation + BasicCost vectorised over the batch axis, one core (loss only)
            "vectorised_numpy_template_evals_per_s_1core": vectorised_numpy_rate()}
    print(json.dumps(line), flush=True)
    D.shutdown()
    return 0


def micro_benchmarks(engine, peak_flops):
    """Streaming kernels timed alone (CUDA events, 3 warm-ups, inputs > L2)."""
    import torch

    import oracle as O
    from slam_decomposition_b200.circuit import Parameter, TemplateCircuit, lower
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
        for _ in range(3):
            engine.loss_grad(desc, X, V, out_loss=loss, out_grad=grad)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 5
        for _ in range(reps):
            engine.loss_grad(desc, X, V, out_loss=loss, out_grad=grad)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) * 1e-3 / reps
        out[f"loss_grad_k{k}"] = {"evals_per_s": B / t, "tflops_alg": B / t * O.F_lossgrad(k) / 1e12,
                                  "frac_of_fp64_peak": B / t * O.F_lossgrad(k) / peak_flops,
                                  "hbm_gbs": B * (16 * desc.n_params + 8) / t / 1e9, "batch": B}
        del X, grad, loss
    U = torch.as_tensor(haar_targets(1 << 21, 99), device=dev)
    for _ in range(3):
        engine.weyl(U)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        engine.weyl(U)
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3 / 5
    out["weyl_c1c2c3"] = {"matrices_per_s": U.shape[0] / t, "hbm_gbs": U.shape[0] * (256 + 24) / t / 1e9}
    # K6 coverage Monte-Carlo (BASELINE configs[1]): 1e7 samples, sqrt(iSWAP) k=3, plain and smush (parallel-drive) templates
    from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv

    n = 10_000_000
    for label, basis in (("coverage_sqiSwap_k3_plain", pdv.plain_template(math.pi / 2, 0.0, 0.5, 3)),
                         ("coverage_sqiSwap_k3_smush", pdv.smush_template(math.pi / 2, 0.0, 0.5, 3))):
        hist = torch.zeros(128 ** 3, dtype=torch.int64, device=dev)
        for _ in range(3):
            pdv.coverage_histogram(basis, 1_000_000, seed=1, hist=hist)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pdv.coverage_histogram(basis, n, seed=2023, hist=hist)
        b.record()
        torch.cuda.synchronize()
        out[label] = {"samples_per_s": n / (a.elapsed_time(b) * 1e-3), "samples": n, "params": basis.desc.n_params}
    # K2 on a parameter-bound smush template (parallel_drive_volume.py:175-199, sqrt(iSWAP) k=3, T=2, P=30): loss + analytic
    # adjoint gradient through the slice exponentials vs the forward evaluation a finite-difference gradient repeats P+1 times
    basis = pdv.smush_template(math.pi / 2, 0.0, 0.5, 3)
    Bs = 1 << 20
    g = torch.Generator(device=dev).manual_seed(77)
    Xs = (torch.rand((Bs, basis.desc.n_params), device=dev, dtype=torch.float64, generator=g) - 0.5) * (8 * math.pi)
    ls = torch.empty(Bs, device=dev, dtype=torch.float64)
    gs = torch.empty_like(Xs)
    for label, want in (("smush_k3_loss_grad_adjoint", True), ("smush_k3_loss_only", False)):
        for _ in range(3):
            engine.loss_grad(basis.desc, Xs, V, out_loss=ls, out_grad=gs if want else None, want_grad=want)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            engine.loss_grad(basis.desc, Xs, V, out_loss=ls, out_grad=gs if want else None, want_grad=want)
        b.record()
        torch.cuda.synchronize()
        out[label] = {"evals_per_s": 3 * Bs / (a.elapsed_time(b) * 1e-3), "batch": Bs, "params": basis.desc.n_params}
    # K4b parallel-drive Weyl trajectories (BASELINE configs[3]; pd_playground.py:169-208): N = 10 slices of the 1Q-phase smush
    # Hamiltonian, R = 5 sub-times each -> one expm + prefix product + c1c2c3 per trajectory point
    Bt, Nn, Rr = 1 << 18, 10, 5
    gate = (torch.rand((Bt, 8), device=dev, dtype=torch.float64, generator=g) - 0.5) * 4
    ax = (torch.rand((Bt, Nn), device=dev, dtype=torch.float64, generator=g) - 0.5) * (4 * math.pi)
    ay = (torch.rand((Bt, Nn), device=dev, dtype=torch.float64, generator=g) - 0.5) * (4 * math.pi)
    for _ in range(3):
        engine.pd_trajectory(gate, ax, ay, 0.1, R=Rr, want_final=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        engine.pd_trajectory(gate, ax, ay, 0.1, R=Rr, want_final=False)
    b.record()
    torch.cuda.synchronize()
    out["pd_trajectory_N10_R5"] = {"points_per_s": 3 * Bt * Nn * Rr / (a.elapsed_time(b) * 1e-3), "trajectories": Bt}
    out["smush_k3_adjoint_vs_fd_gradient"] = (out["smush_k3_loss_grad_adjoint"]["evals_per_s"] * (basis.desc.n_params + 1)
                                              / out["smush_k3_loss_only"]["evals_per_s"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--targets", type=int, default=100000, help="Haar targets per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample length per process")
    ap.add_argument("--no-micro", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

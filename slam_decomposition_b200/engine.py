"""Torch-tensor level wrappers over the C ABI (device memory + streams are torch's; the math is ours).

Every function takes CUDA tensors, launches on ``torch.cuda.current_stream()`` and returns CUDA
tensors.  Nothing here computes on the CPU; a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import SlamNmOpts, SlamOptOpts, SlamTemplateDesc, check, load

__all__ = [
    "template_eval", "loss_grad", "weyl", "lbfgs_solve", "best_gather", "fd_lbfgs_solve", "nm_solve", "nm_defaults", "coverage_mc", "pd_trajectory", "fp64_peak", "opt_defaults",
    "require_cuda",
]


# number of kernel launches issued through this module (bench.py reports it as `gpu_launches`)
LAUNCHES = 0
# when set to a list, lbfgs_solve appends (k, start_event, end_event) for per-launch device timing
LBFGS_EVENTS = None
# bench.py instrumentation for the chained (two-stream) sweep: (start, end) CUDA events around the whole launch chain
LBFGS_SPANS = None


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.SlamError("slam_decomposition_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    load()
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.SlamError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _enter(t: torch.Tensor):
    lib = load()
    check(lib.slam_set_device(t.device.index if t.device.index is not None else torch.cuda.current_device()), "set_device")
    return lib


def template_eval(desc: SlamTemplateDesc, x: torch.Tensor) -> torch.Tensor:
    """K1: x [B,P] float64 -> U [B,4,4] complex128."""
    x = _dev(x, torch.float64, "x")
    if x.dim() != 2 or x.shape[1] != desc.n_params:
        raise ValueError(f"x must be [B,{desc.n_params}]")
    B = x.shape[0]
    U = torch.empty((B, 4, 4), dtype=torch.complex128, device=x.device)
    with torch.cuda.device(x.device):
        lib = _enter(x)
        check(lib.slam_template_eval(C.byref(desc), _ptr(x), max(x.stride(0), desc.n_params), _ptr(U), B, _stream()),
              "slam_template_eval")
    _count()
    return U


def loss_grad(desc: SlamTemplateDesc, x: torch.Tensor, V: torch.Tensor, tgt_idx: Optional[torch.Tensor] = None,
              cost_kind: int = 0, want_grad: bool = True, want_trace: bool = False,
              out_loss: Optional[torch.Tensor] = None, out_grad: Optional[torch.Tensor] = None, lanes: int = 0):
    """K2: (loss [B], grad [B,P] | None, trace [B] complex | None).  `lanes`: team width 1 / 2 / 4 (0 = automatic)."""
    x = _dev(x, torch.float64, "x")
    V = _dev(V, torch.complex128, "V")
    if x.dim() != 2 or x.shape[1] != desc.n_params:
        raise ValueError(f"x must be [B,{desc.n_params}]")
    if V.dim() != 3 or V.shape[1:] != (4, 4):
        raise ValueError("V must be [Nt,4,4]")
    B, P = x.shape
    if tgt_idx is not None:
        tgt_idx = _dev(tgt_idx, torch.int32, "tgt_idx")
        if tgt_idx.numel() != B:
            raise ValueError("tgt_idx must have B entries")
    if out_loss is not None:
        _dev(out_loss, torch.float64, "out_loss")
        if out_loss.shape != (B,) or not out_loss.is_contiguous():
            raise ValueError("loss_grad: out_loss must be a contiguous [B] tensor")
    loss = out_loss if out_loss is not None else torch.empty(B, dtype=torch.float64, device=x.device)
    grad = None
    if want_grad:
        if out_grad is not None:
            _dev(out_grad, torch.float64, "out_grad")
            if out_grad.shape != (B, P) or not out_grad.is_contiguous():
                raise ValueError("loss_grad: out_grad must be a contiguous [B,P] tensor")
        grad = out_grad if out_grad is not None else torch.empty((B, P), dtype=torch.float64, device=x.device)
    trace = torch.empty(B, dtype=torch.complex128, device=x.device) if want_trace else None
    with torch.cuda.device(x.device):
        lib = _enter(x)
        check(lib.slam_loss_grad_lanes(C.byref(desc), _ptr(x), max(x.stride(0), P), _ptr(V), V.shape[0], _ptr(tgt_idx),
                                       int(cost_kind), _ptr(loss), _ptr(grad), P, _ptr(trace), B, int(lanes), _stream()),
              "slam_loss_grad")
    _count()
    return loss, grad, trace


def weyl(U: torch.Tensor, fold: bool = False, round8: bool = False, want_c: bool = True, want_g: bool = False):
    """K3: U [B,4,4] complex128 -> (c [B,3] | None, g [B,3] | None)."""
    U = _dev(U, torch.complex128, "U")
    if U.dim() != 3 or U.shape[1:] != (4, 4):
        raise ValueError("U must be [B,4,4]")
    B = U.shape[0]
    c = torch.empty((B, 3), dtype=torch.float64, device=U.device) if want_c else None
    g = torch.empty((B, 3), dtype=torch.float64, device=U.device) if want_g else None
    flags = (_lib.WEYL_FOLD if fold else 0) | (_lib.WEYL_ROUND8 if round8 else 0)
    with torch.cuda.device(U.device):
        lib = _enter(U)
        check(lib.slam_weyl(_ptr(U), B, _ptr(c), _ptr(g), flags, _stream()), "slam_weyl")
    _count()
    return c, g


def opt_defaults() -> SlamOptOpts:
    o = SlamOptOpts()
    load().slam_opt_defaults(C.byref(o))
    return o


def _check_tables(name: str, out: tuple, Nt: int, restarts: int, P: int) -> tuple:
    """Shape / dtype / contiguity of caller-provided (loss [Nt,R], x [Nt,R,P], iters [Nt,R]) result tables: the kernels
    write them through raw pointers."""
    loss, x, iters = out
    if (loss.shape != (Nt, restarts) or x.shape != (Nt, restarts, P) or iters.shape != (Nt, restarts)
            or not (loss.is_contiguous() and x.is_contiguous() and iters.is_contiguous())):
        raise ValueError(f"{name}: preallocated outputs have the wrong shape")
    _dev(loss, torch.float64, "out loss"), _dev(x, torch.float64, "out x"), _dev(iters, torch.int32, "out iters")
    return loss, x, iters


def lbfgs_solve(desc: SlamTemplateDesc, V: torch.Tensor, restarts: int, opts: SlamOptOpts,
                x0: Optional[torch.Tensor] = None, seed: int = 0, active: Optional[torch.Tensor] = None,
                evals: Optional[torch.Tensor] = None, out: Optional[tuple] = None, best_key: Optional[torch.Tensor] = None):
    """K5: returns (loss [Nt,R], x [Nt,R,P], iters [Nt,R]).  `out` = preallocated (loss, x, iters) to reuse.
    `best_key` = int64 [Nt] tensor holding the packed per-target reduction keys (SlamOptOpts.best_key; initialise to -1 =
    all ones); `best_gather` reads the winners back."""
    V = _dev(V, torch.complex128, "V")
    Nt = V.shape[0]
    P = desc.n_params
    ld = P
    if x0 is not None:
        x0 = _dev(x0, torch.float64, "x0")
        if x0.shape != (Nt, restarts, P):
            raise ValueError(f"x0 must be [{Nt},{restarts},{P}]")
    if active is not None:
        active = _dev(active, torch.int32, "active")
        if active.numel() != Nt:
            raise ValueError("active must have one entry per target")
    o = opts
    if best_key is not None:
        _dev(best_key, torch.int64, "best_key")
        if best_key.shape != (Nt,) or not best_key.is_contiguous():
            raise ValueError("lbfgs_solve: best_key must be a contiguous int64 [Nt] tensor")
        o = SlamOptOpts.from_buffer_copy(opts)
        o.best_key = best_key.data_ptr()
    if out is not None:
        loss, x, iters = _check_tables("lbfgs_solve", out, Nt, restarts, P)
    else:
        loss = torch.empty((Nt, restarts), dtype=torch.float64, device=V.device)
        x = torch.empty((Nt, restarts, P), dtype=torch.float64, device=V.device)
        iters = torch.empty((Nt, restarts), dtype=torch.int32, device=V.device)
    with torch.cuda.device(V.device):
        lib = _enter(V)
        ev = None
        if LBFGS_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        check(lib.slam_lbfgs_solve(C.byref(desc), _ptr(V), Nt, int(restarts), _ptr(x0), ld, C.c_uint64(seed), _ptr(active),
                                   C.byref(o), _ptr(loss), _ptr(x), _ptr(iters), _ptr(evals), _stream()),
              "slam_lbfgs_solve")
        if ev is not None:
            ev[1].record()
            LBFGS_EVENTS.append((desc.k, ev[0], ev[1]))
    _count()
    return loss, x, iters


def best_gather(best_key: torch.Tensor, tables, restarts: int, ldx: int):
    """Winners of the packed per-target reduction: `tables` = [(k, loss [Nt,R], x [Nt,R,P]), ...] of the chained launches ->
    (best_loss [Nt] f64, best_k [Nt] i32, best_P [Nt] i32, best_x [Nt, ldx] f64), one small kernel."""
    _dev(best_key, torch.int64, "best_key")
    Nt = best_key.shape[0]
    n = len(tables)
    ks = (C.c_int32 * n)(*[int(k) for k, _, _ in tables])
    Ps = (C.c_int32 * n)(*[int(x.shape[2]) for _, _, x in tables])
    for _, loss, x in tables:
        _check_tables("best_gather", (loss, x, torch.empty((Nt, restarts), dtype=torch.int32, device=loss.device)), Nt, restarts,
                      x.shape[2])
    lp = (C.c_void_p * n)(*[loss.data_ptr() for _, loss, _ in tables])
    xp = (C.c_void_p * n)(*[x.data_ptr() for _, _, x in tables])
    dev = best_key.device
    bl = torch.empty(Nt, dtype=torch.float64, device=dev)
    bk = torch.empty(Nt, dtype=torch.int32, device=dev)
    bp = torch.empty(Nt, dtype=torch.int32, device=dev)
    bx = torch.empty((Nt, int(ldx)), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        lib = _enter(best_key)
        check(lib.slam_best_gather(_ptr(best_key), Nt, int(restarts), n, ks, Ps, lp, xp, _ptr(bl), _ptr(bk), _ptr(bp), _ptr(bx),
                                   int(ldx), _stream()), "slam_best_gather")
    _count()
    return bl, bk, bp, bx


def fd_lbfgs_solve(desc: SlamTemplateDesc, V: torch.Tensor, restarts: int, opts: SlamOptOpts,
                   x0: Optional[torch.Tensor] = None, seed: int = 0, active: Optional[torch.Tensor] = None,
                   evals: Optional[torch.Tensor] = None, out: Optional[tuple] = None, central=False):
    """K5c: batched L-BFGS over the generic forward objective (parameter-bound smush gates, BasicCostInverse).
    `central`: False / 0 = forward differences (scipy's jac=None), True / 1 = central differences, 2 or "adjoint" =
    analytic adjoint gradient through the smush slices (smush templates only).  Returns (loss [Nt,R], x [Nt,R,P],
    iters [Nt,R]); `evals` counts forward evaluations (one per gradient in adjoint mode)."""
    mode = 2 if central == "adjoint" else int(central)
    V = _dev(V, torch.complex128, "V")
    Nt = V.shape[0]
    P = desc.n_params
    if x0 is not None:
        x0 = _dev(x0, torch.float64, "x0")
        if x0.shape != (Nt, restarts, P):
            raise ValueError(f"x0 must be [{Nt},{restarts},{P}]")
    if active is not None:
        active = _dev(active, torch.int32, "active")
    if out is not None:
        loss, x, iters = _check_tables("fd_lbfgs_solve", out, Nt, restarts, P)
    else:
        loss = torch.empty((Nt, restarts), dtype=torch.float64, device=V.device)
        x = torch.empty((Nt, restarts, P), dtype=torch.float64, device=V.device)
        iters = torch.empty((Nt, restarts), dtype=torch.int32, device=V.device)
    with torch.cuda.device(V.device):
        lib = _enter(V)
        check(lib.slam_fd_lbfgs_solve(C.byref(desc), _ptr(V), Nt, int(restarts), _ptr(x0), P, C.c_uint64(seed), _ptr(active),
                                      C.byref(opts), mode, _ptr(loss), _ptr(x), _ptr(iters), _ptr(evals),
                                      _stream()), "slam_fd_lbfgs_solve")
    _count()
    return loss, x, iters


def nm_defaults() -> SlamNmOpts:
    o = SlamNmOpts()
    load().slam_nm_defaults(C.byref(o))
    return o


def nm_solve(desc: SlamTemplateDesc, V: torch.Tensor, restarts: int, opts: SlamNmOpts, x0: Optional[torch.Tensor] = None,
             seed: int = 0, active: Optional[torch.Tensor] = None, evals: Optional[torch.Tensor] = None,
             out: Optional[tuple] = None):
    """K5b: batched Nelder-Mead, generic objective.  Returns (loss [Nt,R], x [Nt,R,P], iters [Nt,R])."""
    V = _dev(V, torch.complex128, "V")
    Nt = V.shape[0]
    P = desc.n_params
    if x0 is not None:
        x0 = _dev(x0, torch.float64, "x0")
        if x0.shape != (Nt, restarts, P):
            raise ValueError(f"x0 must be [{Nt},{restarts},{P}]")
    if active is not None:
        active = _dev(active, torch.int32, "active")
    if out is not None:
        loss, x, iters = _check_tables("nm_solve", out, Nt, restarts, P)
    else:
        loss = torch.empty((Nt, restarts), dtype=torch.float64, device=V.device)
        x = torch.empty((Nt, restarts, P), dtype=torch.float64, device=V.device)
        iters = torch.empty((Nt, restarts), dtype=torch.int32, device=V.device)
    with torch.cuda.device(V.device):
        lib = _enter(V)
        check(lib.slam_nm_solve(C.byref(desc), _ptr(V), Nt, int(restarts), _ptr(x0), P, C.c_uint64(seed), _ptr(active),
                                C.byref(opts), _ptr(loss), _ptr(x), _ptr(iters), _ptr(evals), _stream()), "slam_nm_solve")
    _count()
    return loss, x, iters


def coverage_mc(desc: SlamTemplateDesc, seed: int, first_sample: int, n_samples: int, lo: float, hi: float,
                nbins: int = 128, hist: Optional[torch.Tensor] = None, want_coords: bool = False,
                device: Optional[torch.device] = None):
    """K6: returns (hist int64[nbins^3] (accumulated into `hist` if given), coords [n,3] | None)."""
    device = device or (hist.device if hist is not None else require_cuda())
    if hist is None:
        hist = torch.zeros(nbins ** 3, dtype=torch.int64, device=device)
    else:
        if hist.dtype != torch.int64 or not hist.is_cuda:
            raise TypeError("hist must be a CUDA int64 tensor")
        if not hist.is_contiguous() or hist.numel() != int(nbins) ** 3:
            # the kernel indexes hist[(b0 * nbins + b1) * nbins + b2] through a raw pointer
            raise ValueError(f"hist must be contiguous with nbins^3 = {int(nbins) ** 3} entries, got {hist.numel()}")
    coords = torch.empty((n_samples, 3), dtype=torch.float64, device=device) if want_coords else None
    with torch.cuda.device(device):
        lib = _enter(hist)
        check(lib.slam_coverage_mc(C.byref(desc), C.c_uint64(seed), int(first_sample), int(n_samples), float(lo), float(hi),
                                   int(nbins), _ptr(hist), _ptr(coords), _stream()), "slam_coverage_mc")
    _count()
    return hist, coords


def pd_trajectory(gate: torch.Tensor, gx: torch.Tensor, gy: torch.Tensor, dt: float, R: int = 5, fold: bool = True,
                  round8: bool = True, want_coords: bool = True, want_final: bool = True):
    """K4b: gate [B,8] (one gate row per trajectory) or [B,N,8] (one per slice: multi-segment pulses), gx/gy [B,N] ->
    (coords [B,N,R,3] | None, Ufinal [B,4,4] | None)."""
    gate = _dev(gate, torch.float64, "gate")
    gx = _dev(gx, torch.float64, "gx")
    gy = _dev(gy, torch.float64, "gy")
    B, N = gx.shape
    per_slice = gate.dim() == 3
    if gate.shape != ((B, N, 8) if per_slice else (B, 8)) or gy.shape != (B, N):
        raise ValueError("gate must be [B,8] or [B,N,8], gx/gy [B,N]")
    coords = torch.empty((B, N, R, 3), dtype=torch.float64, device=gx.device) if want_coords else None
    Uf = torch.empty((B, 4, 4), dtype=torch.complex128, device=gx.device) if want_final else None
    flags = (_lib.WEYL_FOLD if fold else 0) | (_lib.WEYL_ROUND8 if round8 else 0)
    with torch.cuda.device(gx.device):
        lib = _enter(gx)
        fn = lib.slam_pd_trajectory_slices if per_slice else lib.slam_pd_trajectory
        check(fn(_ptr(gate), _ptr(gx), _ptr(gy), N, R, float(dt), flags, _ptr(coords), _ptr(Uf), B, _stream()),
              "slam_pd_trajectory")
    _count()
    return coords, Uf


def fp64_peak(iters: int = 4096) -> Tuple[float, float]:
    """Measured register-resident DFMA throughput of the current device: (FLOP/s, ms)."""
    require_cuda()
    lib = load()
    check(lib.slam_set_device(torch.cuda.current_device()), "set_device")
    f = C.c_double(0.0)
    ms = C.c_double(0.0)
    check(lib.slam_fp64_peak(int(iters), C.byref(f), C.byref(ms)), "slam_fp64_peak")
    return f.value, ms.value

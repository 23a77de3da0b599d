"""ctypes binding of libslam_b200.so (the C ABI declared in include/slam_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

SLAM_MAX_K = 16
SLAM_MAX_SLOTS = 40
SLAM_MAX_PARAMS = 256

GATE_RISWAP, GATE_CG, GATE_SMUSH, GATE_SMUSH_1QPHASE, GATE_FIXED = range(5)
(COST_BASIC, COST_SQUARE, COST_BASIC_INVERSE, COST_MAKHLIN_FUNCTIONAL, COST_MAKHLIN_EUCLIDEAN, COST_WEYL_EUCLIDEAN,
 COST_BASIC_REDUCED, COST_SQUARE_REDUCED) = range(8)
WEYL_FOLD, WEYL_ROUND8 = 1, 2
ABI_VERSION = 4

import os

LIB_PATH = Path(os.environ.get("SLAM_B200_LIB") or (Path(__file__).resolve().parent / "libslam_b200.so"))


class SlamTemplateDesc(C.Structure):
    _fields_ = [
        ("gate_kind", C.c_int32),
        ("k", C.c_int32),
        ("T", C.c_int32),
        ("n_slots", C.c_int32),
        ("n_params", C.c_int32),
        ("no_exterior_1q", C.c_int32),
        ("vz_only", C.c_int32),
        ("reserved", C.c_int32),
        ("p1q", (C.c_int32 * 6) * (SLAM_MAX_K + 1)),
        ("slot_param", (C.c_int32 * SLAM_MAX_SLOTS) * SLAM_MAX_K),
        ("slot_const", (C.c_double * SLAM_MAX_SLOTS) * SLAM_MAX_K),
        ("fixed_gate", C.c_double * 32),
    ]


class SlamOptOpts(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32),
        ("history", C.c_int32),
        ("cost_kind", C.c_int32),
        ("early_exit", C.c_int32),
        ("success_threshold", C.c_double),
        ("f_stop", C.c_double),
        ("gtol", C.c_double),
        ("gtol_far", C.c_double),
        ("f_far", C.c_double),
        ("x0_lo", C.c_double),
        ("x0_hi", C.c_double),
        ("trace_cap", C.c_int32),
        ("diag", C.c_int32),
        ("trace_loss", C.c_void_p),
        ("trace_x", C.c_void_p),
        ("lower", C.c_void_p),
        ("upper", C.c_void_p),
        ("solved_in", C.c_void_p),
        ("solved_out", C.c_void_p),
        ("con_max", C.c_double),
        ("con_mu", C.c_double),
        ("con_lambda", C.c_void_p),
        ("best_key", C.c_void_p),
        ("tune_lanes", C.c_int32),
        ("tune_sm_threads", C.c_int32),
        ("tune_hist_min", C.c_int32),
        ("tune_max_teams", C.c_int32),
    ]


class SlamNmOpts(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32),
        ("cost_kind", C.c_int32),
        ("early_exit", C.c_int32),
        ("reserved", C.c_int32),
        ("success_threshold", C.c_double),
        ("xatol", C.c_double),
        ("fatol", C.c_double),
        ("x0_lo", C.c_double),
        ("x0_hi", C.c_double),
    ]


class SlamError(RuntimeError):
    pass


_P = C.c_void_p
_PROTOS = {
    "slam_abi_version": (C.c_int, []),
    "slam_status_string": (C.c_char_p, [C.c_int]),
    "slam_last_cuda_error": (C.c_char_p, []),
    "slam_device_count": (C.c_int, []),
    "slam_set_device": (C.c_int, [C.c_int]),
    "slam_template_eval": (C.c_int, [C.POINTER(SlamTemplateDesc), _P, C.c_int64, _P, C.c_int64, _P]),
    "slam_loss_grad": (C.c_int, [C.POINTER(SlamTemplateDesc), _P, C.c_int64, _P, C.c_int64, _P, C.c_int32, _P, _P,
                                 C.c_int64, _P, C.c_int64, _P]),
    "slam_loss_grad_lanes": (C.c_int, [C.POINTER(SlamTemplateDesc), _P, C.c_int64, _P, C.c_int64, _P, C.c_int32, _P, _P,
                                       C.c_int64, _P, C.c_int64, C.c_int32, _P]),
    "slam_weyl": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int32, _P]),
    "slam_opt_defaults": (None, [C.POINTER(SlamOptOpts)]),
    "slam_lbfgs_solve": (C.c_int, [C.POINTER(SlamTemplateDesc), _P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_uint64, _P,
                                   C.POINTER(SlamOptOpts), _P, _P, _P, _P, _P]),
    "slam_best_gather": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "slam_fd_lbfgs_solve": (C.c_int, [C.POINTER(SlamTemplateDesc), _P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_uint64, _P,
                                      C.POINTER(SlamOptOpts), C.c_int32, _P, _P, _P, _P, _P]),
    "slam_nm_defaults": (None, [C.POINTER(SlamNmOpts)]),
    "slam_nm_solve": (C.c_int, [C.POINTER(SlamTemplateDesc), _P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_uint64, _P,
                                C.POINTER(SlamNmOpts), _P, _P, _P, _P, _P]),
    "slam_coverage_mc": (C.c_int, [C.POINTER(SlamTemplateDesc), C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                   C.c_int32, _P, _P, _P]),
    "slam_pd_trajectory": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_double, C.c_int32, _P, _P, C.c_int64, _P]),
    "slam_pd_trajectory_slices": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_double, C.c_int32, _P, _P, C.c_int64, _P]),
    "slam_fp64_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "slam_selftest_sincos": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SlamError(
            f"{LIB_PATH} not found: the sm_100a CUDA extension is not built. "
            "Run `python -m slam_decomposition_b200.build` (there is no CPU fallback)."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.slam_abi_version() != ABI_VERSION:
        raise SlamError("libslam_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status == 0:
        return
    lib = load()
    msg = lib.slam_status_string(status).decode()
    if status == -3:
        msg += ": " + lib.slam_last_cuda_error().decode()
    if status == -1:
        raise ValueError(f"{what}: {msg}")  # the reference raises ValueError for bad templates (basis.py:127-128)
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    raise SlamError(f"{what}: {msg}")

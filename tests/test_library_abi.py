"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/slam_b200.h declares.
No compute calls are made here (no GPU in the CPU test tier)."""
import ctypes
import os
import re

import numpy as np

import pytest

from slam_decomposition_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "slam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slam_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)


def test_abi_version_and_status_strings(lib):
    assert lib.slam_abi_version() == 4
    assert lib.slam_status_string(0) == b"ok"
    assert b"invalid" in lib.slam_status_string(-1)
    assert b"unsupported" in lib.slam_status_string(-2)


def test_struct_layouts_match_the_header():
    # SlamTemplateDesc: 8 int32 + int32[K+1][6] + int32[K][40] + double[K][40] + double[32], K = SLAM_MAX_K = 16
    assert ctypes.sizeof(_lib.SlamTemplateDesc) == 32 + 17 * 6 * 4 + 16 * 40 * 4 + 16 * 40 * 8 + 32 * 8
    assert _lib.SlamTemplateDesc.slot_const.offset % 8 == 0
    # 4 int32, 7 double, 2 int32, 9 pointers/doubles (v3), then v4: best_key pointer + 4 int32 (tune_*)
    assert ctypes.sizeof(_lib.SlamOptOpts) == 4 * 4 + 7 * 8 + 2 * 4 + 9 * 8 + 8 + 4 * 4


def test_opt_defaults_follow_the_reference_constants(lib):
    o = _lib.SlamOptOpts()
    lib.slam_opt_defaults(ctypes.byref(o))
    assert o.max_iter == 2500            # optimizer.py:274
    assert o.success_threshold == 1e-10  # optimizer.py:18
    assert o.gtol_far == 1e-5            # scipy BFGS default gtol
    assert (o.x0_lo, round(o.x0_hi, 12)) == (0.0, round(6.283185307179586, 12))


def test_argument_validation_happens_before_any_cuda_call(lib):
    """Bad descriptors are rejected on the host (ValueError semantics of basis.py:127-128) even without a device."""
    d = _lib.SlamTemplateDesc()
    d.k = 0
    assert lib.slam_template_eval(ctypes.byref(d), None, 0, None, 1, None) == -1
    d.k = 2
    d.gate_kind = 99
    assert lib.slam_loss_grad(ctypes.byref(d), None, 0, ctypes.c_void_p(8), 1, None, 0, ctypes.c_void_p(8), None, 0, None, 1,
                              None) == -1
    assert lib.slam_weyl(None, 1, None, None, 0, None) == -1
    assert lib.slam_coverage_mc(None, 0, 0, 1, 0.0, 1.0, 8, None, None, None) == -1


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "slam_decomposition_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_fast_sincos_matches_libm():
    """The kernels' (cos, sin) routine (slam_core.cuh fast_sincos), evaluated on the host through the diagnostic
    entry point: within 1 ulp-ish of libm over the fast-path range, exact hand-off to libm outside it."""
    import ctypes as C

    from slam_decomposition_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(11)
    xs = np.concatenate([
        rng.uniform(-8 * np.pi, 8 * np.pi, 2_000_000), rng.uniform(-105615.0, 105615.0, 2_000_000),
        rng.uniform(-1e-3, 1e-3, 100_000), np.arange(-4000, 4001) * (np.pi / 4), np.arange(-4000, 4001) * (np.pi / 2),
        np.array([0.0, -0.0, 1e-300, 105615.0, -105615.0, 105616.0, 1e9, -3e15]),
    ])
    s = np.empty_like(xs)
    c = np.empty_like(xs)
    rc = lib.slam_selftest_sincos(xs.ctypes.data_as(C.c_void_p), xs.size, s.ctypes.data_as(C.c_void_p),
                                  c.ctypes.data_as(C.c_void_p))
    assert rc == 0
    rs, rc_ = np.sin(xs), np.cos(xs)
    # absolute error bound: 1 ulp of a value in [0.5, 1) is 1.1e-16; allow 2.3e-16 (reduction + polynomial)
    assert np.max(np.abs(s - rs)) < 2.3e-16
    assert np.max(np.abs(c - rc_)) < 2.3e-16
    # relative accuracy near the zeros of sin / cos (no cancellation blow-up from the reduction)
    big = np.abs(rs) > 1e-12
    assert np.max(np.abs(s[big] - rs[big]) / np.abs(rs[big])) < 1e-15
    bigc = np.abs(rc_) > 1e-12
    assert np.max(np.abs(c[bigc] - rc_[bigc]) / np.abs(rc_[bigc])) < 1e-15
    assert np.all(s * s + c * c - 1.0 < 5e-16)
    nan_s = np.empty(2)
    nan_c = np.empty(2)
    bad = np.array([np.nan, np.inf])
    lib.slam_selftest_sincos(bad.ctypes.data_as(C.c_void_p), 2, nan_s.ctypes.data_as(C.c_void_p), nan_c.ctypes.data_as(C.c_void_p))
    assert np.all(np.isnan(nan_s)) and np.all(np.isnan(nan_c))


def test_ctypes_field_offsets_match_the_c_compiler(tmp_path):
    """Every field of the three POD structs sits where gcc puts it: a C program including include/slam_b200.h prints
    offsetof() for each field and the struct sizes; the ctypes mirrors in _lib.py must agree byte for byte."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    structs = {"SlamTemplateDesc": _lib.SlamTemplateDesc, "SlamOptOpts": _lib.SlamOptOpts, "SlamNmOpts": _lib.SlamNmOpts}
    lines = ["#include <stddef.h>", "#include <stdio.h>", '#include "slam_b200.h"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{name} {fname} %zu\\n", offsetof({name}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    subprocess.run([gcc, "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        sname, field, val = ln.split()
        cls = structs[sname]
        if field == "size":
            assert ctypes.sizeof(cls) == int(val), sname
        else:
            assert getattr(cls, field).offset == int(val), (sname, field)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


def test_optimiser_option_validation_needs_no_device(lib):
    """The ABI v2/v3 option fields are validated on the host, before any CUDA call: chained launches need early_exit and a
    flag array to propagate into; the circuit-cost constraint belongs to slam_fd_lbfgs_solve's finite-difference modes."""
    from helpers import make_pair

    desc, _ = make_pair("riswap", (0.5,), k=2)
    fake = ctypes.c_void_p(8)  # never dereferenced: every call below must fail in the argument checks

    def opts():
        o = _lib.SlamOptOpts()
        lib.slam_opt_defaults(ctypes.byref(o))
        return o

    def lbfgs(o):
        return lib.slam_lbfgs_solve(ctypes.byref(desc), fake, 4, 2, None, desc.n_params, 0, None, ctypes.byref(o), fake, fake,
                                    fake, None, None)

    def fd(o, mode, d=desc):
        return lib.slam_fd_lbfgs_solve(ctypes.byref(d), fake, 4, 2, None, d.n_params, 0, None, ctypes.byref(o), mode, fake,
                                       fake, fake, None, None)

    o = opts()
    assert (o.solved_in, o.solved_out, o.con_mu, o.con_lambda) == (None, None, 0.0, None)
    o.solved_in = 8  # without solved_out
    assert lbfgs(o) == -1
    o = opts()
    o.solved_out, o.early_exit = 8, 0
    assert lbfgs(o) == -1
    o = opts()
    o.con_mu = 10.0
    assert lbfgs(o) == -2   # constrained runs go through slam_fd_lbfgs_solve
    o = opts()
    assert fd(o, 3) == -1   # gradient mode out of range
    assert fd(o, 2) == -2   # the adjoint mode is for parameter-bound smush templates
    o.con_mu = -1.0
    assert fd(o, 1) == -1
    o = opts()
    o.cost_kind = _lib.COST_MAKHLIN_FUNCTIONAL
    assert fd(o, 0) == -2   # coordinate-based functionals have no gradient (8-dp rounding)
    # ABI v4: the packed per-target key holds 8 bits of restart index; box bounds come as a pair; tuning values are checked
    o = opts()
    o.best_key = 8
    assert lib.slam_lbfgs_solve(ctypes.byref(desc), fake, 4, 257, None, desc.n_params, 0, None, ctypes.byref(o), fake, fake,
                                fake, None, None) == -2
    o = opts()
    o.lower = 8  # without upper
    assert lbfgs(o) == -1
    assert fd(o, 1) == -1
    o = opts()
    o.tune_lanes = 3
    assert lbfgs(o) == -1
    assert lib.slam_best_gather(None, 4, 2, 1, None, None, None, None, None, None, None, None, 4, None) == -1
    sdesc, _ = make_pair("smush", ("Q", "Q", np.pi / 2, 0.0, "Q", "Q", "Q", "Q", 0.5), k=1, T=2, no_exterior_1q=True)
    o = opts()
    o.con_mu = 1.0
    assert fd(o, 2, sdesc) == -2  # the constraint term is differenced, not adjoint
    # the history length of the generic-objective solver is a runtime option, range checked; diag defaults to off
    o = opts()
    assert (o.history, o.diag) == (0, 0)
    o.history = 9
    assert fd(o, 2, sdesc) == -1 and lbfgs(o) == -1
    o.history = -1
    assert fd(o, 2, sdesc) == -1

"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/slam_b200.h declares.
No compute calls are made here (no GPU in the CPU test tier)."""
import ctypes
import os
import re

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
    assert lib.slam_abi_version() == 1
    assert lib.slam_status_string(0) == b"ok"
    assert b"invalid" in lib.slam_status_string(-1)
    assert b"unsupported" in lib.slam_status_string(-2)


def test_struct_layouts_match_the_header():
    # SlamTemplateDesc: 8 int32 + int32[9][6] + int32[8][40] + double[8][40] + double[32]
    assert ctypes.sizeof(_lib.SlamTemplateDesc) == 32 + 9 * 6 * 4 + 8 * 40 * 4 + 8 * 40 * 8 + 32 * 8
    assert _lib.SlamTemplateDesc.slot_const.offset % 8 == 0
    assert ctypes.sizeof(_lib.SlamOptOpts) == 4 * 4 + 7 * 8 + 2 * 4 + 4 * 8


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

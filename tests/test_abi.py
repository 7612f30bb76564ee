"""The C-ABI library loads, exports every symbol include/mpcgpu.h declares, and fails LOUDLY without a
GPU (no CPU fallback).  No compute is attempted here."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as entry
from mpcgpu import _capi, shell3x3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        entry.build()
    return _capi.load_library()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "mpcgpu.h")).read()
    declared = set(re.findall(r"\b(mpcgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None


def test_struct_layout_matches_header():
    # 8 int32 + 14 pointers + double + 4 pointers = 32 + 112 + 8 + 32
    assert C.sizeof(_capi.ProblemStruct) == 184
    assert C.sizeof(_capi.Counters) == 6 * 8 + 3 * 8


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ps, keep = _capi.make_problem_struct(shell3x3(2))
    h = C.c_void_p()
    rc = lib.mpcgpu_create(C.byref(ps), 0, C.byref(h))
    assert rc == 2 and not h.value
    msg = lib.mpcgpu_last_error(None).decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    from mpcgpu import Evaluator, MpcGpuError
    with pytest.raises(MpcGpuError):
        Evaluator(shell3x3(2))


def test_product_does_not_import_oracle():
    """The product path never imports, links, loads or executes anything under oracle/ (comments may cite it)."""
    pkg = os.path.join(ROOT, "model-predictive-control-tuning_b200")
    pat = re.compile(r"(^\s*(import|from)\s+oracle\b)|liboracle|oracle\.py|orc_[a-z_]+\s*\(|#include\s+\".*oracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), f


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/mpcgpu.h must compile as C99 (and as C++) on its own, with no CUDA or torch types."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "mpcgpu.h"\nint main(void) { return sizeof(mpcgpu_problem) + sizeof(mpcgpu_nmpc_problem) + sizeof(mpcgpu_dtc_problem) > 0 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)])
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "mpcgpu.h")).read(), flags=re.S)      # code only
    assert "cudaStream_t" not in hdr and "#include <cuda" not in hdr      # streams cross the boundary as void *
    assert "torch" not in hdr and "at::" not in hdr

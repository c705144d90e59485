"""Run-time specialisation of predicate programs (fle_jit.cu), CPU side: the generated CUDA C++
is a function of the program's SHAPE only (not of its literals), and it compiles for sm_100a with
NVRTC on this box (no device needed) -- so syntax errors in the kernel skeleton or the generator
surface here.  The GPU parity of the specialised kernels is in test_gpu_jit.py."""
import shutil
import subprocess

import pytest

from fle_b200 import ALL, AND, EQ, GE, GT, IN, LE, LT, NONE, OR
from fle_b200.capi import E_INVALID, OK, eval_codegen

Q6 = [(GE, 0, 730), (LE, 0, 1094), (AND,), (GE, 1, 5), (LE, 1, 7), (AND,), (AND,), (LT, 2, 23), (AND,)]
SHAPES = {
    "q6": ([12, 4, 6], Q6),
    "or_tree": ([3, 2, 6], [(IN, 0, [1, 4]), (EQ, 1, 2), (OR,), (LT, 2, 23), (AND,)]),
    "odd_widths": ([1, 31, 17], [(EQ, 0, 1), (GT, 1, 5), (AND,), (LE, 2, 99), (AND,)]),
    "in8_in_tree": ([13, 5], [(IN, 0, list(range(8))), (GE, 1, 3), (AND,), (IN, 0, [9, 10, 11, 12, 13, 14, 15, 16]), (OR,)]),
    "long_in": ([12, 32], [(IN, 0, list(range(20))), (LT, 1, 5), (OR,)]),
    "consts": ([12, 4], [(ALL,), (LT, 0, 1000), (AND,), (NONE,), (OR,), (GE, 1, 3), (AND,)]),
    "eight_columns": ([5] * 8, [(LT, 0, 3)] + sum([[(GE, c, 2), (AND,)] for c in range(1, 8)], [])),
}


def nvrtc_available():
    rc, log = eval_codegen([4, 4], [(LT, 0, 1), (LT, 1, 1), (AND,)], compile=True)
    return rc == OK, log


def test_text_depends_on_shape_not_literals():
    rc, a = eval_codegen([12, 4, 6], Q6)
    assert rc == OK and "jit_eval_group" in a and "struct JitArgs" in a
    other = [(LT, 0, 1), (GT, 0, 2000), (AND,), (LE, 1, 9), (GE, 1, 0), (AND,), (AND,), (GE, 2, 50), (AND,)]
    rc, b = eval_codegen([12, 4, 6], other)
    assert rc == OK and a == b  # Lt/Le/Gt/Ge differ only in JitArgs::ini / ::neg
    rc, c = eval_codegen([12, 5, 6], Q6)
    assert rc == OK and c != a  # a bit width is shape
    rc, d = eval_codegen([3, 2], [(IN, 0, [1, 4]), (EQ, 1, 2), (OR,)])
    rc2, e = eval_codegen([3, 2], [(IN, 0, [0, 6]), (EQ, 1, 3), (OR,)])
    assert rc == OK and rc2 == OK and d == e
    rc, f = eval_codegen([3, 2], [(IN, 0, [0, 6, 2]), (EQ, 1, 3), (OR,)])
    assert rc == OK and f != d  # the length of a short IN list is shape (it is unrolled)


def test_generated_work_is_the_plane_algebra():
    rc, t = eval_codegen([12, 4, 6], Q6)
    assert rc == OK
    # 5 chains: 2 x 12 + 2 x 4 + 6 plane steps, every plane loaded once per column
    assert t.count("jit_lt(") == 38 and t.count("jit_eq(") == 0
    assert t.count("load_pair<0, 12>") == 6 and t.count("load_pair<1, 4>") == 2 and t.count("load_pair<2, 6>") == 3
    assert "uint32_t mask[38]" in t


def test_bad_programs_are_rejected():
    rc, _ = eval_codegen([12], [(AND,)])
    assert rc == E_INVALID
    rc, _ = eval_codegen([40], [(LT, 0, 1)])
    assert rc == E_INVALID


@pytest.mark.parametrize("name", list(SHAPES))
def test_shapes_compile_with_nvrtc(name):
    ok, why = nvrtc_available()
    if not ok:
        pytest.skip("NVRTC unavailable here: " + why[:200])
    bws, prog = SHAPES[name]
    rc, log = eval_codegen(bws, prog, compile=True)
    assert rc == OK, log[:4000]


def test_specialised_sass_is_register_resident():
    ok, why = nvrtc_available()
    if not ok or not shutil.which("cuobjdump"):
        pytest.skip("NVRTC / cuobjdump unavailable here")
    rc, cubin = eval_codegen([12, 4, 6], Q6, compile=2)
    assert rc == OK and len(cubin) > 1000
    import os
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin)
    try:
        res = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
        sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
    finally:
        os.unlink(f.name)
    assert "fle_jit_kernel" in res and "STACK:0" in res and "LOCAL:0" in res  # no spills
    assert "UBLKCP" in sass and "SYNCS" in sass  # TMA bulk copies + mbarrier waits


def test_fused_aggregate_shape_compiles():
    ok, why = nvrtc_available()
    if not ok:
        pytest.skip("NVRTC unavailable here: " + why[:200])
    rc, text = eval_codegen([12, 4, 6], Q6, agg_col=1)
    assert rc == OK and "#define FLE_JIT_AGG 1" in text and "jit_agg_code_smem" in text
    assert "kJitAggCol = 1" in text and "kJitAggBw = 4" in text
    rc, log = eval_codegen([12, 4, 6], Q6, compile=True, agg_col=1)
    assert rc == OK, log[:4000]
    rc, log = eval_codegen([12, 5, 7], Q6, compile=True, agg_col=2)  # odd width
    assert rc == OK, log[:4000]
    rc, _ = eval_codegen([12, 4, 6], Q6, agg_col=5)
    assert rc != OK


@pytest.mark.parametrize("name", ["q6", "odd_widths", "long_in", "in8_in_tree"])
def test_skewed_shapes_compile_with_nvrtc(name):
    """The variant for columns at arbitrary byte addresses (page images evaluated in place): another shape --
    whole-chunk loads + funnel shift instead of the aligned plane loads -- that must compile and stay spill-free."""
    ok, why = nvrtc_available()
    if not ok:
        pytest.skip("NVRTC unavailable here: " + why[:200])
    bws, prog = SHAPES[name]
    rc, text = eval_codegen(bws, prog, skew=True)
    rc0, plain = eval_codegen(bws, prog)
    assert rc == OK and rc0 == OK and text != plain and "FLE_JIT_SKEW" in text and "load_all<" in text
    rc, log = eval_codegen(bws, prog, compile=True, skew=True)
    assert rc == OK, log[:4000]
    rc, _ = eval_codegen([12, 4, 6], Q6, agg_col=1, skew=True)
    assert rc != OK  # the fused aggregate is not built for skewed columns (the caller gets the two-kernel form)

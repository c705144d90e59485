"""GPU: the host-side C++ drop-ins (impala-avx2-parquet-scanner_b200/host/) behind the reference's
own class interfaces.

1. The C++ test programs that follow the reference's fle-test.cc / dict-test.cc plus the scanner
   batch-interface test.
2. oracle/ref_driver.cc -- the harness that drives the reference's FleEncoder/FleDecoder -- is
   compiled a second time against the drop-in util/fle-encoding.h; identical call sequences
   (Put/Flush, Get, Get(skip), Eq..In at arbitrary cursors, the 1024-row window + Skip loop, the
   AssembleRows survivor loop, And/Or trees) must give identical results to the reference
   (run live from oracle/_ref when present, and always to the golden vectors it produced)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, valid_row_mask_planes
from oracle.oracle import AND, EQ, GE, GT, IN, LE, LT, OR, Ref

pytestmark = pytest.mark.gpu

PKG = os.path.join(ROOT, "impala-avx2-parquet-scanner_b200")
HOST = os.path.join(PKG, "host")
OPS = (EQ, LT, LE, GT, GE)
PRE_GETS = (0, 13, 64, 100)


@pytest.fixture(scope="module")
def host_tests(lib_path):
    subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "tests")


@pytest.mark.parametrize("name", ["fle-test", "dict-test", "scanner-test", "group-test"])
def test_cpp_suite(host_tests, name):
    p = subprocess.run([os.path.join(host_tests, name)], capture_output=True, text=True, timeout=900)
    print(p.stdout[-3000:], p.stderr[-3000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "0 failure(s)" in p.stdout


@pytest.fixture(scope="module")
def hostdrv(lib_path):
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libfle_hostdrv.so")
    src = os.path.join(ROOT, "oracle", "ref_driver.cc")
    deps = [src, lib_path] + [os.path.join(HOST, "util", f) for f in os.listdir(os.path.join(HOST, "util"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=gnu++17", "-O2", "-fPIC", "-shared", "-pthread", "-w",
                               f"-I{HOST}", f"-I{HOST}/util", f"-I{HOST}/boost_compat", f"-I{ROOT}/include",
                               "-o", so, src, f"-L{PKG}", "-lfle_b200", f"-Wl,-rpath,{PKG}"])
    return Ref(so)


@pytest.mark.parametrize("bw", list(range(1, 33)))
def test_driver_vs_golden(hostdrv, golden, bw):
    n = int(golden["meta_n"][0])
    vals, words = golden[f"vals_{bw}"], golden[f"words_{bw}"]
    mine = hostdrv.pack(vals, bw)                       # FleEncoder::Put x n + Flush
    assert mine.shape == words.shape
    full = (n // 64) * bw
    assert (mine[:full] == words[:full]).all()
    m = valid_row_mask_planes(n % 64)
    assert ((mine[full:] & m) == (words[full:] & m)).all()
    dec, got = hostdrv.unpack(words, bw, n)             # FleDecoder::Get x n
    assert got == n and (dec == vals).all()
    _, got = hostdrv.unpack(words, bw, n + 64)          # Get() == false at the buffer guard
    assert got == ((n + 63) // 64) * 64
    lits, G = golden[f"pred_lits_{bw}"], golden[f"pred_{bw}"]
    for oi, op in enumerate(OPS):
        for li, lit in enumerate(lits):
            for pi, pre in enumerate(PRE_GETS):         # head (scalar) + body (planes) at any cursor
                bm = hostdrv.pred(op, words, bw, pre, n - pre, [lit])
                assert (bm == G[oi, li, pi, : len(bm)]).all(), (bw, op, int(lit), pre)
    for pi, pre in enumerate(PRE_GETS[:3]):
        bm = hostdrv.pred(IN, words, bw, pre, n - pre, golden[f"in_lits_{bw}"])
        assert (bm == golden[f"in_{bw}"][pi, : len(bm)]).all()
    lit = int(golden[f"win_lit_{bw}"][0])
    assert (hostdrv.pred_windowed(GE, words, bw, n, 100, [lit]) == golden[f"win_ge_{bw}"]).all()
    assert (hostdrv.scan_materialize(LT, words, bw, n, 100, [lit]) == golden[f"mat_lt_{bw}"]).all()
    sv, got = hostdrv.get_skip(words, bw, golden[f"skips_{bw}"])
    assert got == len(sv) and (sv == golden[f"skipvals_{bw}"]).all()


def test_driver_trees_vs_golden(hostdrv, golden):
    n = int(golden["tree_n"][0])
    cols = [(golden[f"tree_words_{i}"], bw) for i, bw in enumerate((12, 4, 6))]
    q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
          (AND, None, None), (LT, 2, 23), (AND, None, None)]
    assert (hostdrv.eval_tree(cols, q6, n) == golden["tree_q6"]).all()
    orv = [(IN, 1, [2, 9]), (EQ, 2, 17), (OR, None, None), (LT, 0, 1200), (AND, None, None), (GT, 1, 8),
           (OR, None, None)]
    assert (hostdrv.eval_tree(cols, orv, n) == golden["tree_or"]).all()


@pytest.mark.parametrize("bw", [1, 5, 8, 12, 16, 21, 32])
def test_driver_vs_live_reference(hostdrv, ref, bw):
    """Same harness, same inputs, reference classes vs drop-in classes (needs oracle/_ref)."""
    rng = np.random.default_rng(4242 + bw)
    n = 64 * 300 + 45
    vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
    words = ref.pack(vals, bw)
    for op in OPS:
        lit = int(vals[rng.integers(0, n)])
        for pre in (0, 7, 63, 64, 65, 1000):
            assert (hostdrv.pred(op, words, bw, pre, n - pre, [lit]) == ref.pred(op, words, bw, pre, n - pre, [lit])).all()
        for window in (1024, 1000, 64, 37):
            assert (hostdrv.pred_windowed(op, words, bw, n, window, [lit]) ==
                    ref.pred_windowed(op, words, bw, n, window, [lit])).all()
        assert (hostdrv.scan_materialize(op, words, bw, n, 1024, [lit]) ==
                ref.scan_materialize(op, words, bw, n, 1024, [lit])).all()
    lits = rng.choice(vals, size=33).astype(np.uint64)
    assert (hostdrv.pred_windowed(IN, words, bw, n, 1024, lits) == ref.pred_windowed(IN, words, bw, n, 1024, lits)).all()
    skips = rng.integers(0, 200, size=150).astype(np.int32)
    a, ga = hostdrv.get_skip(words, bw, skips)
    b, gb = ref.get_skip(words, bw, skips)
    assert ga == gb and (a[:ga] == b[:gb]).all()


@pytest.mark.parametrize("bw", [1, 7, 12, 16, 32])
def test_deviations_pinned_against_live_reference(hostdrv, ref, bw):
    """DESIGN.md's deliberate deviations, run on the reference itself and on the drop-in in BOTH modes: a buffer that
    holds a multiple of 64 values, Skip() landing exactly on its end (fle-encoding.h:381-402) and the AssembleRows loop
    that returns when that Skip reports false (hdfs-parquet-scanner.cc:1107-1113)."""
    rng = np.random.default_rng(900 + bw)
    n = 64 * 48
    vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
    words = ref.pack(vals, bw)
    nbytes = (n // 64) * bw * 8                     # exactly the groups that hold values (no Flush spare group)
    lit = int(np.sort(vals)[n // 2])
    try:
        for skips in ([n], [64, n - 64], [n - 1, 1], [1000, 1000, n - 2000], [n + 64], [100, 200, 300]):
            want, want_get = ref.skip_seq(words, nbytes, bw, skips)
            hostdrv.set_compat(True)
            got, got_get = hostdrv.skip_seq(words, nbytes, bw, skips)
            assert got == want, ("reference_compat", skips)
            # (after a Skip() that failed the reference's Get() hands out stale staging values and reports success;
            #  the drop-in reports false -- compared only where every Skip succeeded)
            if all(want):
                assert got_get == want_get
            hostdrv.set_compat(False)
            got, got_get = hostdrv.skip_seq(words, nbytes, bw, skips)
            if sum(skips) == n:                      # the ONE case the default mode answers differently
                assert want[-1] == 0 and got[-1] == 1 and got[:-1] == want[:-1] and got_get == 0
            else:
                assert got == want, skips
        for window in (1024, 64):
            want = ref.scan_materialize_checked(LT, words, nbytes, bw, n, window, [lit])
            full = ref.scan_materialize(LT, words, bw, n, window, [lit])   # the loop without the early return
            hostdrv.set_compat(True)
            assert (hostdrv.scan_materialize_checked(LT, words, nbytes, bw, n, window, [lit]) == want).all()
            hostdrv.set_compat(False)
            got = hostdrv.scan_materialize_checked(LT, words, nbytes, bw, n, window, [lit])
            assert (got == full).all() and (full == vals[vals < lit]).all()
            assert len(want) <= len(full)           # the reference may stop early; the default mode never does
    finally:
        hostdrv.set_compat(False)

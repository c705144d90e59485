"""CPU: the C restatement against the reference's own code run live (oracle/_ref, compiled
from /root/reference/fle-encoding.h unmodified).  Skipped only where _ref was never built."""
import numpy as np
import pytest

from conftest import valid_row_mask_planes
from oracle.oracle import AND, EQ, GE, GT, IN, LE, LT, OR


@pytest.mark.parametrize("bw", [1, 2, 3, 5, 8, 9, 12, 16, 17, 24, 31, 32])
def test_random_roundtrip_and_predicates(oracle, ref, bw):
    rng = np.random.default_rng(1000 + bw)
    n = 64 * 33 + 5
    vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
    rw, ow = ref.pack(vals, bw), oracle.pack(vals, bw)
    full = (n // 64) * bw
    assert rw.nbytes == ow.nbytes == oracle.encoded_bytes(n, bw)
    assert (rw[:full] == ow[:full]).all()
    m = valid_row_mask_planes(n % 64)
    assert ((rw[full:] & m) == (ow[full:] & m)).all()
    dec, got = ref.unpack(ow, bw, n)  # reference decodes the oracle's encoding
    assert got == n and (dec == vals).all()
    for op in (EQ, LT, LE, GT, GE):
        for lit in (int(vals[3]), int(rng.integers(0, 1 << bw))):
            for pre in (0, 1, 63, 64, 65, 200):
                a = ref.pred(op, rw, bw, pre, n - pre, [lit])
                b, _ = oracle.pred(op, rw, bw, pre, n - pre, [lit])
                assert (a == b).all(), (bw, op, lit, pre)
    lits = rng.choice(vals, size=19).astype(np.uint64)
    a = ref.pred(IN, rw, bw, 0, n, lits)
    b, _ = oracle.pred(IN, rw, bw, 0, n, lits)
    assert (a == b).all()
    assert (ref.pred_windowed(LE, rw, bw, n, 1024, [int(vals[9])]) == oracle.pred(LE, rw, bw, 0, n, [int(vals[9])])[0]).all()


def test_fle_test_values_shapes(oracle, ref):
    # Fle.TestValues (fle-test.cc:246-262): 1 value, 1024 x (v % 2^bw), 1024 x 0, 1024 x 1
    for bw in range(1, 33):
        for vals in (np.array([1], dtype=np.uint32), (np.arange(1024, dtype=np.uint64) % (1 << bw)).astype(np.uint32),
                     np.zeros(1024, dtype=np.uint32), np.ones(1024, dtype=np.uint32)):
            w = ref.pack(vals, bw)
            assert w.nbytes == oracle.encoded_bytes(len(vals), bw)
            assert (oracle.unpack(w, bw, 0, len(vals)) == vals).all()
            assert (ref.unpack(oracle.pack(vals, bw), bw, len(vals))[0] == vals).all()


def test_tree_matches_reference(oracle, ref):
    rng = np.random.default_rng(77)
    n = 64 * 21 + 40
    c = [rng.integers(0, 1 << bw, size=n).astype(np.uint32) for bw in (12, 3, 16)]
    cols = [(oracle.pack(v, bw), bw) for v, bw in zip(c, (12, 3, 16))]
    prog = [(GE, 0, 100), (LT, 0, 3000), (AND, None, None), (IN, 1, [1, 5, 6]), (OR, None, None),
            (GT, 2, 40000), (EQ, 2, 12345), (OR, None, None), (AND, None, None)]
    assert (ref.eval_tree(cols, prog, n) == oracle.eval_tree(cols, prog, n)[0]).all()

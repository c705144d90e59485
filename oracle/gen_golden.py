"""TEST INFRASTRUCTURE: generates tests/golden/fle_golden.npz from the reference itself.

Run in the build container (where /root/reference is mounted and oracle/_ref was built):

    python oracle/gen_golden.py

Every array in the file is an OUTPUT OF THE REFERENCE'S OWN CODE (fle-encoding.h compiled
unmodified, oracle/ref_driver.cc) on seeded inputs, plus the known answers the reference's
fle-test.cc holds (encoded lengths, fle-test.cc:219,224,238; byte images of the full words,
fle-test.cc:217-223,235-236).  The GPU box has no /root/reference; tests there compare the
CUDA path and the C restatement against this file.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import EQ, GE, GT, IN, LE, LT, AND, OR, Ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fle_golden.npz")
N = 64 * 5 + 37  # ragged tail on purpose
PRE_GETS = (0, 13, 64, 100)
OPS = (EQ, LT, LE, GT, GE)


def literals_for(vals, bw):
    return [int(vals[5]), 1 << (bw - 1), (1 << bw) - 1, 0]


def main():
    ref = Ref()
    rng = np.random.default_rng(0xF1E0)
    g = {}
    g["meta_n"] = np.array([N])
    for bw in range(1, 33):
        vals = rng.integers(0, 1 << bw, size=N, dtype=np.uint64).astype(np.uint32)
        words = ref.pack(vals, bw)
        g[f"vals_{bw}"] = vals
        g[f"words_{bw}"] = words  # last group's unfilled rows hold stale encoder state
        dec, got = ref.unpack(words, bw, N)
        assert got == N
        g[f"unpacked_{bw}"] = dec
        lits = literals_for(vals, bw)
        preds = np.zeros((len(OPS), len(lits), len(PRE_GETS), (N + 63) // 64), dtype=np.uint64)
        for oi, op in enumerate(OPS):
            for li, lit in enumerate(lits):
                for pi, pre in enumerate(PRE_GETS):
                    bm = ref.pred(op, words, bw, pre, N - pre, [lit])
                    preds[oi, li, pi, : len(bm)] = bm
        g[f"pred_{bw}"] = preds
        g[f"pred_lits_{bw}"] = np.array(lits, dtype=np.uint64)
        in_lits = rng.choice(vals, size=7).astype(np.uint64)
        g[f"in_lits_{bw}"] = in_lits
        nw = (N + 63) // 64
        rows = []
        for pre in PRE_GETS[:3]:
            bm = ref.pred(IN, words, bw, pre, N - pre, in_lits)
            rows.append(np.pad(bm, (0, nw - len(bm))))
        g[f"in_{bw}"] = np.stack(rows)
        lit = int(vals[7])
        g[f"win_ge_{bw}"] = ref.pred_windowed(GE, words, bw, N, 100, [lit])
        g[f"win_lit_{bw}"] = np.array([lit], dtype=np.uint64)
        g[f"mat_lt_{bw}"] = ref.scan_materialize(LT, words, bw, N, 100, [lit])
        skips = rng.integers(0, 9, size=40).astype(np.int32)
        sv, gotk = ref.get_skip(words, bw, skips)
        assert gotk == len(skips)
        g[f"skips_{bw}"] = skips
        g[f"skipvals_{bw}"] = sv

    # compound trees over three columns (Q6 shape + an OR variant), reference-evaluated leaf by leaf
    n = 64 * 9 + 11
    c0 = rng.integers(0, 2526, size=n).astype(np.uint32)
    c1 = rng.integers(0, 11, size=n).astype(np.uint32)
    c2 = rng.integers(0, 50, size=n).astype(np.uint32)
    cols = [(ref.pack(c0, 12), 12), (ref.pack(c1, 4), 4), (ref.pack(c2, 6), 6)]
    for i, (w, _) in enumerate(cols):
        g[f"tree_words_{i}"] = w
    g["tree_vals"] = np.stack([c0, c1, c2])
    g["tree_n"] = np.array([n])
    q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
          (AND, None, None), (LT, 2, 23), (AND, None, None)]
    g["tree_q6"] = ref.eval_tree(cols, q6, n)
    orv = [(IN, 1, [2, 9]), (EQ, 2, 17), (OR, None, None), (LT, 0, 1200), (AND, None, None), (GT, 1, 8),
           (OR, None, None)]
    g["tree_or"] = ref.eval_tree(cols, orv, n)

    # fle-test.cc Fle.SpecificSequences known answers
    v = np.array([0] * 50 + [1] * 50, dtype=np.uint32)
    g["spec_a_bw1"] = ref.pack(v, 1)   # len 16 (:219); word0 == 0x3fff (:217)
    g["spec_a_bw2"] = ref.pack(v, 2)   # len 32 (:224)
    v = np.array([i % 2 for i in range(100)], dtype=np.uint32)
    g["spec_b_bw1"] = ref.pack(v, 1)   # len 16 (:238); word0 == 0x5555555555555555 (:235)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()

"""CPU: the C restatement (oracle/fle_oracle.c) against the golden vectors the reference
itself produced (tests/golden/fle_golden.npz, made by oracle/gen_golden.py), including the
known answers of the reference's fle-test.cc."""
import numpy as np
import pytest

from conftest import valid_row_mask_planes
from oracle.oracle import AND, EQ, GE, GT, IN, LE, LT, OR

OPS = (EQ, LT, LE, GT, GE)
PRE_GETS = (0, 13, 64, 100)
BWS = list(range(1, 33))


def test_fle_test_known_answers(golden, oracle):
    # Fle.SpecificSequences: 50x0 + 50x1 -> 16 bytes @bw1, 32 bytes @bw2 (fle-test.cc:219,224)
    assert golden["spec_a_bw1"].nbytes == 16 and golden["spec_a_bw2"].nbytes == 32
    assert golden["spec_b_bw1"].nbytes == 16  # fle-test.cc:238
    # commented byte images hold for every FULL word (fle-test.cc:217, :235)
    assert int(golden["spec_a_bw1"][0]) == 0x3FFF
    assert int(golden["spec_b_bw1"][0]) == 0x5555555555555555
    # the tail word carries stale staging bits below the 36 valid rows
    m = valid_row_mask_planes(36)
    assert (golden["spec_a_bw1"][1] & m) == np.uint64(0xFFFFFFFFF0000000)
    v = np.array([0] * 50 + [1] * 50, dtype=np.uint32)
    o = oracle.pack(v, 1)
    assert o.nbytes == 16 and int(o[0]) == 0x3FFF and int(o[1]) == 0xFFFFFFFFF0000000
    assert oracle.pack(v, 2).nbytes == 32
    assert oracle.encoded_bytes(0, 7) == 56  # Flush always emits a group


@pytest.mark.parametrize("bw", BWS)
def test_pack_unpack_vs_golden(golden, oracle, bw):
    n = int(golden["meta_n"][0])
    vals, words = golden[f"vals_{bw}"], golden[f"words_{bw}"]
    mine = oracle.pack(vals, bw)
    assert mine.shape == words.shape
    full = (n // 64) * bw
    assert (mine[:full] == words[:full]).all()
    m = valid_row_mask_planes(n % 64)
    assert ((mine[full:] & m) == (words[full:] & m)).all()
    assert (golden[f"unpacked_{bw}"] == vals).all()
    assert (oracle.unpack(words, bw, 0, n) == vals).all()
    assert (oracle.unpack(words, bw, 77, 200) == vals[77:277]).all()


@pytest.mark.parametrize("bw", BWS)
def test_predicates_vs_golden(golden, oracle, bw):
    n = int(golden["meta_n"][0])
    words = golden[f"words_{bw}"]
    lits = golden[f"pred_lits_{bw}"]
    G = golden[f"pred_{bw}"]
    for oi, op in enumerate(OPS):
        for li, lit in enumerate(lits):
            for pi, pre in enumerate(PRE_GETS):
                bm, cnt = oracle.pred(op, words, bw, pre, n - pre, [lit])
                want = G[oi, li, pi, : len(bm)]
                assert (bm == want).all(), (bw, op, int(lit), pre)
                assert cnt == int(np.unpackbits(want.view(np.uint8)).sum())
                if pre % 64 == 0:
                    bm2, _ = oracle.pred(op, words, bw, pre, n - pre, [lit], planes=True)
                    assert (bm2 == want).all()
                # cursor-faithful model: head (scalar) + body (planes)
                d = oracle.decoder(words, words.nbytes, bw)
                for _ in range(pre):
                    assert d.get()[0]
                assert (d.pred(op, n - pre, [lit]) == want).all()
    for pi, pre in enumerate(PRE_GETS[:3]):
        bm, _ = oracle.pred(IN, words, bw, pre, n - pre, golden[f"in_lits_{bw}"])
        assert (bm == golden[f"in_{bw}"][pi, : len(bm)]).all()


@pytest.mark.parametrize("bw", BWS)
def test_window_skip_materialise_vs_golden(golden, oracle, bw):
    n = int(golden["meta_n"][0])
    words, vals = golden[f"words_{bw}"], golden[f"vals_{bw}"]
    lit = int(golden[f"win_lit_{bw}"][0])
    # scanner loop (predicate on <=100 rows, Skip(100)) == whole-column evaluation
    whole, _ = oracle.pred(GE, words, bw, 0, n, [lit])
    assert (whole == golden[f"win_ge_{bw}"]).all()
    d = oracle.decoder(words, words.nbytes, bw)
    out = np.zeros((n + 63) // 64, dtype=np.uint64)
    done = 0
    while done < n:
        w = min(100, n - done)
        bm = d.pred(GE, w, [lit])
        for i in range(w):
            if (int(bm[i >> 6]) >> (i & 63)) & 1:
                out[(done + i) >> 6] |= np.uint64(1 << ((done + i) & 63))
        assert d.skip(w)
        done += w
    assert (out == whole).all()
    # survivor materialisation via Get(val, skip)
    assert (golden[f"mat_lt_{bw}"] == vals[vals < lit]).all()
    d = oracle.decoder(words, words.nbytes, bw)
    pos = -1
    for s, want in zip(golden[f"skips_{bw}"], golden[f"skipvals_{bw}"]):
        ok, v = d.get_skip(int(s))
        pos += int(s) + 1
        assert ok and v == int(want) == int(vals[pos])


def test_trees_vs_golden(golden, oracle):
    n = int(golden["tree_n"][0])
    cols = [(golden[f"tree_words_{i}"], bw) for i, bw in enumerate((12, 4, 6))]
    q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
          (AND, None, None), (LT, 2, 23), (AND, None, None)]
    bm, cnt = oracle.eval_tree(cols, q6, n)
    assert (bm == golden["tree_q6"]).all()
    c0, c1, c2 = golden["tree_vals"]
    truth = (c0 >= 730) & (c0 <= 1094) & (c1 >= 5) & (c1 <= 7) & (c2 < 23)
    assert cnt == int(truth.sum())
    orv = [(IN, 1, [2, 9]), (EQ, 2, 17), (OR, None, None), (LT, 0, 1200), (AND, None, None), (GT, 1, 8),
           (OR, None, None)]
    bm, cnt = oracle.eval_tree(cols, orv, n)
    assert (bm == golden["tree_or"]).all()
    truth = ((np.isin(c1, [2, 9]) | (c2 == 17)) & (c0 < 1200)) | (c1 > 8)
    assert cnt == int(truth.sum())


def test_intersect_bitset(oracle):
    rng = np.random.default_rng(5)
    n = 1000
    root_bits = rng.random(n) < 0.7
    sub_bits = rng.random(int(root_bits.sum())) < 0.4
    root = np.packbits(root_bits, bitorder="little")
    root = np.pad(root, (0, (-len(root)) % 8)).view(np.uint64)
    sub = np.packbits(sub_bits, bitorder="little")
    sub = np.pad(sub, (0, (-len(sub)) % 8)).view(np.uint64)
    got, cnt = oracle.intersect_bitset(root, n, sub)
    want = np.zeros(n, dtype=bool)
    want[np.flatnonzero(root_bits)] = sub_bits  # hdfs-parquet-scanner.cc:326-331
    got_bits = np.unpackbits(got.view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (got_bits == want).all() and cnt == int(want.sum())


def test_dict_translate(oracle):
    # dict-encoding.h:461-541 on a sorted dictionary
    d = np.array([10, 20, 30, 40], dtype=np.int64)
    from oracle.oracle import ALL, NONE
    T = oracle.dict_translate
    assert T(EQ, d, 30) == (EQ, 2) and T(EQ, d, 25)[0] == NONE and T(EQ, d, 50)[0] == NONE
    assert T(GT, d, 40)[0] == NONE and T(GT, d, 5)[0] == ALL and T(GT, d, 20) == (GE, 2) and T(GT, d, 25) == (GE, 2)
    assert T(LT, d, 10)[0] == NONE and T(LT, d, 41)[0] == ALL and T(LT, d, 30) == (LT, 2) and T(LT, d, 25) == (LT, 2)
    assert T(GE, d, 41)[0] == NONE and T(GE, d, 10)[0] == ALL and T(GE, d, 30) == (GE, 2) and T(GE, d, 25) == (GE, 2)
    assert T(LE, d, 9)[0] == NONE and T(LE, d, 40)[0] == ALL and T(LE, d, 30) == (LT, 3) and T(LE, d, 25) == (LT, 2)
    assert list(oracle.dict_translate_in(d, [40, 11, 10, 99])) == [3, 0]
    assert [oracle.dict_bit_width(n) for n in (0, 1, 2, 3, 4, 5, 40000)] == [0, 1, 1, 2, 2, 3, 16]

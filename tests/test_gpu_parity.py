"""GPU parity tests proper: the CUDA path, called through the C ABI (libfle_b200.so), against
(1) the golden vectors the reference produced, (2) the CPU oracle on the same seeded inputs and
(3) size-independent properties at BASELINE.json's full sizes.  Integer/bit work: bit-exact."""
import numpy as np
import pytest

from conftest import tail_mask_words, valid_row_mask_planes
from oracle.oracle import ALL, AND, EQ, GE, GT, IN, LE, LT, NONE, OR

pytestmark = pytest.mark.gpu

OPS = (EQ, LT, LE, GT, GE)
BWS = list(range(1, 33))
Q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
      (AND, None, None), (LT, 2, 23), (AND, None, None)]
ORV = [(IN, 1, [2, 9]), (EQ, 2, 17), (OR, None, None), (LT, 0, 1200), (AND, None, None), (GT, 1, 8),
       (OR, None, None)]


def popcount(words):
    return int(np.unpackbits(np.ascontiguousarray(words).view(np.uint8)).sum())


# ------------------------------------------------------------------------------------------
# golden vectors (outputs of the reference's own code)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("codec_variant", [1, 0, 2, 3, 4])
@pytest.mark.parametrize("bw", BWS)
def test_codec_vs_golden(ctx, golden, bw, codec_variant):
    ctx.set_option("codec_variant", codec_variant)
    n = int(golden["meta_n"][0])
    vals, words = golden[f"vals_{bw}"], golden[f"words_{bw}"]
    mine = ctx.pack(vals, bw)
    assert mine.shape == words.shape
    full = (n // 64) * bw
    assert (mine[:full] == words[:full]).all()
    m = valid_row_mask_planes(n % 64)
    assert ((mine[full:] & m) == (words[full:] & m)).all()
    assert (mine[full:] & ~m == 0).all()  # our padding rows are zero
    assert (ctx.unpack(words, bw, 0, n) == vals).all()  # stale tail bits in `words` must not leak
    for first, cnt in ((0, 1), (1, 1), (13, 100), (63, 2), (64, 64), (100, n - 100), (n - 1, 1)):
        assert (ctx.unpack(words, bw, first, cnt) == vals[first:first + cnt]).all(), (first, cnt)
    ctx.set_option("codec_variant", 1)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("bw", BWS)
def test_predicates_vs_golden(ctx, golden, bw, variant):
    ctx.set_option("eval_variant", variant)
    try:
        n = int(golden["meta_n"][0])
        words, lits, G = golden[f"words_{bw}"], golden[f"pred_lits_{bw}"], golden[f"pred_{bw}"]
        for oi, op in enumerate(OPS):
            for li, lit in enumerate(lits):
                for pi, pre in ((0, 0), (2, 64)):
                    bm, cnt = ctx.eval([(words, bw)], [(op, 0, int(lit))], n - pre, first_row=pre)
                    want = G[oi, li, pi, : len(bm)]
                    assert (bm == want).all(), (bw, op, int(lit), pre)
                    assert cnt == popcount(want)
        for pi, pre in ((0, 0), (2, 64)):
            bm, cnt = ctx.eval([(words, bw)], [(IN, 0, golden[f"in_lits_{bw}"])], n - pre, first_row=pre)
            assert (bm == golden[f"in_{bw}"][pi, : len(bm)]).all()
        bm, _ = ctx.eval([(words, bw)], [(GE, 0, int(golden[f"win_lit_{bw}"][0]))], n)
        assert (bm == golden[f"win_ge_{bw}"]).all()
    finally:
        ctx.set_option("eval_variant", 1)


@pytest.mark.parametrize("variant", [0, 1, 2, "jit"])
def test_trees_vs_golden(ctx, golden, variant):
    ctx.set_option("eval_variant", 1 if variant == "jit" else variant)
    ctx.set_option("eval_jit", 2 if variant == "jit" else 0)
    try:
        n = int(golden["tree_n"][0])
        cols = [(golden[f"tree_words_{i}"], bw) for i, bw in enumerate((12, 4, 6))]
        bm, cnt = ctx.eval(cols, Q6, n)
        assert (bm == golden["tree_q6"]).all() and cnt == popcount(golden["tree_q6"])
        bm, cnt = ctx.eval(cols, ORV, n)
        assert (bm == golden["tree_or"]).all() and cnt == popcount(golden["tree_or"])
    finally:
        ctx.set_option("eval_variant", 1)
        ctx.set_option("eval_jit", 1)


# ------------------------------------------------------------------------------------------
# seeded random inputs vs the oracle
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bw", BWS)
def test_single_column_vs_oracle(ctx, oracle, bw):
    rng = np.random.default_rng(0xF1E00000 + bw)
    n = 64 * 4096 * 3 + 64 * 37 + 29  # several TMA tiles + ragged groups + ragged rows
    vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
    words = oracle.pack(vals, bw)
    assert (ctx.pack(vals, bw) == words).all()
    assert (ctx.unpack(words, bw, 0, n) == vals).all()
    D = 1 << bw
    lo, hi = D // 4, max(D // 4, (3 * D) // 4 - 1)
    progs = [
        [(GE, 0, lo), (LE, 0, hi), (AND, None, None)],
        [(LT, 0, lo), (GT, 0, hi), (OR, None, None)],
        [(EQ, 0, int(vals[11]))],
        [(IN, 0, rng.choice(vals, size=8).astype(np.uint64))],
    ]
    for variant in (1, 2, 0):
        ctx.set_option("eval_variant", variant)
        for prog in progs:
            bm, cnt = ctx.eval([(words, bw)], prog, n)
            want, wcnt = oracle.eval_tree([(words, bw)], prog, n)
            assert (bm == want).all() and cnt == wcnt, (bw, variant, prog[0][0])
    ctx.set_option("eval_variant", 1)


@pytest.mark.parametrize("opts", [dict(), dict(eval_variant=2), dict(eval_variant=2, eval_stages=2, eval_gpt=2),
                                  dict(eval_variant=2, eval_stages=4, eval_ctas_per_sm=1), dict(eval_variant=0),
                                  dict(h2d_chunk_rows=64 * 4096), dict(eval_jit=2), dict(eval_jit=2, eval_gpt=1, eval_stages=2),
                                  dict(eval_jit=0)])
def test_multi_column_programs_vs_oracle(ctx, oracle, opts):
    rng = np.random.default_rng(2024)
    n = 64 * 20000 + 17
    bws = (12, 4, 6, 3, 16)
    doms = (2526, 11, 50, 7, 40000)
    vals = [rng.integers(0, d, size=n).astype(np.uint32) for d in doms]
    cols = [(oracle.pack(v, bw), bw) for v, bw in zip(vals, bws)]
    progs = {
        "q6": (cols[:3], Q6),
        "or": (cols[:3], ORV),
        "q12": ([cols[3], cols[0]], [(IN, 0, [1, 4]), (GE, 1, 400), (AND, None, None), (LT, 1, 765), (AND, None, None)]),
        "in64": ([cols[4]], [(IN, 0, rng.choice(vals[4], size=64).astype(np.uint64))]),
        "in256_and": ([cols[0], cols[2]], [(IN, 0, rng.choice(vals[0], size=256).astype(np.uint64)), (LT, 1, 30),
                                            (AND, None, None)]),
        "consts": (cols[:2], [(ALL, None, None), (LT, 0, 1000), (AND, None, None), (NONE, None, None), (OR, None, None),
                              (GE, 1, 3), (AND, None, None)]),
        "deep": (cols[:3], [(LT, 0, 500), (GT, 0, 2000), (OR, None, None), (EQ, 1, 3), (EQ, 1, 4), (OR, None, None),
                            (AND, None, None), (GE, 2, 10), (LE, 2, 40), (AND, None, None), (OR, None, None),
                            (LT, 1, 9), (AND, None, None)]),
        "five_on_one": ([cols[0]], [(GE, 0, 10), (LE, 0, 2500), (AND, None, None), (GT, 0, 11), (AND, None, None),
                                     (LT, 0, 2499), (AND, None, None), (GE, 0, 12), (AND, None, None)]),
    }
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        for name, (pc, prog) in progs.items():
            bm, cnt = ctx.eval(pc, prog, n)
            want, wcnt = oracle.eval_tree(pc, prog, n)
            assert (bm == want).all() and cnt == wcnt, name
    finally:
        for k in opts:
            ctx.set_option(k, {"eval_variant": 1, "eval_jit": 1}.get(k, 0))


def test_edge_sizes_and_errors(ctx, oracle):
    import fle_b200
    from fle_b200 import capi
    rng = np.random.default_rng(9)
    for n in (0, 1, 63, 64, 65, 127, 128, 129, 4095, 4096, 4097):
        vals = rng.integers(0, 1 << 9, size=n).astype(np.uint32)
        words = oracle.pack(vals, 9)
        assert (ctx.pack(vals, 9) == words).all()
        if n:
            assert (ctx.unpack(words, 9, 0, n) == vals).all()
        bm, cnt = ctx.eval([(words, 9)], [(LT, 0, 200)], n)
        want, wcnt = oracle.pred(LT, words, 9, 0, n, [200])
        assert (bm == want).all() and cnt == wcnt
    w = oracle.pack(np.arange(640, dtype=np.uint32) % 512, 9)
    for bad_bw in (0, 33, -1):
        with pytest.raises(fle_b200.FleError) as ei:
            ctx.eval([(w, bad_bw)], [(LT, 0, 1)], 640)
        assert ei.value.code == capi.E_INVALID
    with pytest.raises(fle_b200.FleError) as ei:  # rows beyond the buffer == Get() false at the guard
        ctx.unpack(w, 9, 0, 641)
    assert ei.value.code == capi.E_BUFFER
    with pytest.raises(fle_b200.FleError) as ei:
        ctx.eval([(w, 9)], [(LT, 0, 1)], 100, first_row=7)
    assert ei.value.code == capi.E_ALIGN
    with pytest.raises(fle_b200.FleError) as ei:
        ctx.eval([(w, 9)], [(LT, 0, 1), (AND, None, None)], 100)
    assert ei.value.code == capi.E_INVALID
    with pytest.raises(fle_b200.FleError) as ei:
        ctx.eval([(w, 9)], [(LT, 1, 1)], 100)
    assert ei.value.code == capi.E_INVALID
    # literal wider than the column: folded to what the comparison means for 9-bit codes ...
    bm, cnt = ctx.eval([(w, 9)], [(EQ, 0, 512 + 5)], 640)
    assert cnt == 0 and not bm.any()
    bm, cnt = ctx.eval([(w, 9)], [(LT, 0, 512)], 640)
    assert cnt == 640
    bm, cnt = ctx.eval([(w, 9)], [(GE, 0, 1 << 40)], 640)
    assert cnt == 0
    bm, cnt = ctx.eval([(w, 9)], [(IN, 0, [512 + 5, 7, 1 << 33])], 640)
    assert (bm == oracle.pred(EQ, w, 9, 0, 640, [7])[0]).all()
    # ... unless the reference's plane algebra is asked for: only the low bw bits take part (fle-encoding.h:7965-7971)
    ctx.set_option("reference_compat", 1)
    try:
        bm, _ = ctx.eval([(w, 9)], [(EQ, 0, 512 + 5)], 640)
        assert (bm == oracle.pred(EQ, w, 9, 0, 640, [5])[0]).all()
        bm, cnt = ctx.eval([(w, 9)], [(LT, 0, 512)], 640)
        assert cnt == 0
    finally:
        ctx.set_option("reference_compat", 0)


def test_intersect_bitset_vs_oracle(ctx, oracle):
    rng = np.random.default_rng(31)
    for n, p_root in ((1, 1.0), (64, 0.5), (1000, 0.7), (64 * 5000 + 3, 0.97), (64 * 5000 + 3, 0.03), (70000, 1.0)):
        root_bits = rng.random(n) < p_root
        m = int(root_bits.sum())
        sub_bits = rng.random(m) < 0.4
        root = np.packbits(root_bits, bitorder="little")
        root = np.pad(root, (0, (-len(root)) % 8)).view(np.uint64)
        sub = np.packbits(sub_bits, bitorder="little")
        sub = np.pad(sub, (0, (-len(sub)) % 8)).view(np.uint64)
        want, _ = oracle.intersect_bitset(root, n, sub)
        got = ctx.intersect_bitset(root, n, sub, m)
        assert (got == want).all(), (n, p_root)


@pytest.mark.parametrize("null_frac", [0.0, 0.02, 0.6])
def test_nullable_columns_vs_oracle(ctx, oracle, null_frac):
    """ColumnReader<T>::Eq..In for max_def_level > 0 (hdfs-parquet-scanner.cc:333-445)."""
    rng = np.random.default_rng(55)
    n = 64 * 3000 + 21
    notnull = [rng.random(n) >= null_frac, rng.random(n) >= null_frac / 2]
    bws, doms = (12, 5), (2526, 25)
    cols, ocols = [], []
    for nn, bw, d in zip(notnull, bws, doms):
        data = rng.integers(0, d, size=int(nn.sum())).astype(np.uint32)
        cols.append((oracle.pack(data, bw), bw, oracle.pack(nn.astype(np.uint32), 1)))
        ocols.append((data, nn))
    req = rng.integers(0, 50, size=n).astype(np.uint32)
    cols.append((oracle.pack(req, 6), 6))
    prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (IN, 1, [3, 4, 20]), (OR, None, None), (LT, 2, 23),
            (AND, None, None)]

    def leaf(op, ci, lits):
        if ci == 2:
            return oracle.pred(op, cols[2][0], 6, 0, n, lits)[0]
        data, nn = ocols[ci]
        root = oracle.pred(EQ, cols[ci][2], 1, 0, n, [1])[0]           # fle_def_levels_->Eq(n, bits, max_def)
        sub = oracle.pred(op, cols[ci][0], bws[ci], 0, len(data), lits)[0]  # dict_decoder_->Op(count, ...)
        return oracle.intersect_bitset(root, n, sub)[0]                # IntersectBitset

    # a second tree with larger single-column subtrees (fused per column in the data-ordinal domain),
    # a two-leaf subtree on the required column and a whole-program-on-one-column case
    prog2 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (EQ, 0, 5), (OR, None, None),
             (LT, 2, 10), (GT, 2, 40), (OR, None, None), (IN, 1, [3, 4, 20]), (AND, None, None), (AND, None, None)]
    prog3 = [(GE, 1, 3), (LE, 1, 9), (AND, None, None), (EQ, 1, 20), (OR, None, None)]
    # chains the one-kernel NULL path takes (null_fused = 2): two nullable BETWEENs AND a compare on the required
    # column (the Q6 shape), an OR chain, a single leaf, the same nullable column twice, NOT BETWEEN, <= 4 compares
    prog4 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 3), (LE, 1, 9), (AND, None, None), (AND, None, None),
             (LT, 2, 23), (AND, None, None)]
    prog5 = [(EQ, 0, 5), (LT, 1, 7), (OR, None, None), (GT, 2, 40), (OR, None, None)]
    prog6 = [(LT, 0, 100)]
    prog7 = [(LT, 0, 2000), (GE, 1, 3), (AND, None, None), (GT, 0, 10), (AND, None, None)]
    prog8 = [(LT, 0, 730), (GT, 0, 1094), (OR, None, None), (LE, 1, 20), (AND, None, None)]
    prog9 = [(GT, 1, 2), (LT, 1, 20), (AND, None, None), (GE, 1, 4), (AND, None, None), (LE, 1, 18), (AND, None, None),
             (GE, 2, 1), (AND, None, None)]
    chains = (prog4, prog5, prog6, prog7, prog8, prog9)
    for pr in (prog, prog2, prog3) + chains:
        st = []
        for op, ci, lits in pr:
            if op <= IN:
                st.append(leaf(op, ci, lits))
            else:
                b = st.pop()
                st[-1] = st[-1] & b if op == AND else st[-1] | b
        # a chain of single-pass subtrees in one kernel (default; the grid pinned to 1 / 3 CTAs walks the tile pipeline) /
        # single-pass subtrees evaluate + deposit in one kernel per column / data-domain bitmap, then deposit;
        # twice each: the second call knows that the file holds NULLs and does not launch the optimistic scan
        for fused, ctas in ((2, 0), (2, 1), (2, 3), (1, 0), (0, 0)):
            ctx.set_option("null_fused", fused)
            ctx.set_option("null_chain_ctas", ctas)
            try:
                for rep in range(2):
                    l0 = ctx.launch_count
                    bm, cnt = ctx.eval(cols, pr, n)
                    assert (bm == st[0]).all() and cnt == popcount(st[0]), (pr, fused, ctas)
                    if fused == 2 and rep == 1 and null_frac > 0 and pr in chains:
                        assert ctx.launch_count - l0 == 3, (pr, ctx.launch_count - l0)  # tile counts, prefixes, the chain
            finally:
                ctx.set_option("null_fused", 2)
                ctx.set_option("null_chain_ctas", 0)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 16383, 16384, 16385, 3 * 16384, 5 * 16384 + 1, 64 * 777 + 13])
@pytest.mark.parametrize("null_frac", [0.0, 0.3, 1.0])
def test_nullable_chain_sizes(ctx, oracle, n, null_frac):
    """The NULL chain at tile edges (16384 rows per tile), tiny inputs, an all-NULL column and a column without NULLs
    beside one that has them; grid pinned to one CTA as well, so that one CTA walks every tile."""
    rng = np.random.default_rng(n % 997 + int(null_frac * 10))
    nn = [rng.random(n) >= null_frac if null_frac < 1.0 else np.zeros(n, dtype=bool), rng.random(n) >= 0.1]
    bws, doms = (11, 6), (2000, 50)
    cols, want = [], None
    for k in range(2):
        data = rng.integers(0, doms[k], size=int(nn[k].sum())).astype(np.uint32)
        dw = oracle.pack(data, bws[k]) if len(data) else np.zeros(bws[k], dtype=np.uint64)
        defw = oracle.pack(nn[k].astype(np.uint32), 1)
        cols.append((dw, bws[k], defw))
        root = oracle.pred(EQ, defw, 1, 0, n, [1])[0]
        if len(data):
            sub = oracle.pred(LT, dw, bws[k], 0, len(data), [doms[k] // 2])[0]
            leaf = oracle.intersect_bitset(root, n, sub)[0]
        else:
            leaf = np.zeros_like(root)
        want = leaf if want is None else (want & leaf)
    prog = [(LT, 0, doms[0] // 2), (LT, 1, doms[1] // 2), (AND, None, None)]
    for ctas in (0, 1):
        ctx.set_option("null_chain_ctas", ctas)
        try:
            for _ in range(2):
                bm, cnt = ctx.eval(cols, prog, n)
                assert (bm == want).all() and cnt == popcount(want), (n, null_frac, ctas)
        finally:
            ctx.set_option("null_chain_ctas", 0)


@pytest.mark.parametrize("bws", [(20, 20), (32, 32), (32,), (17, 3, 25, 9)])
def test_nullable_wide_columns_chain_or_fallback(ctx, oracle, bws):
    """The one-kernel NULL chain stages two tiles of every column in shared memory: wide columns that do not fit fall
    back to the per-column kernels -- same answer either way (ColumnReader<T>::Eq..In + IntersectBitset)."""
    rng = np.random.default_rng(sum(bws))
    n = 64 * 1500 + 7
    cols, want = [], None
    prog = []
    for c, bw in enumerate(bws):
        nn = rng.random(n) >= 0.05
        dom = min(1 << bw, 1 << 31)
        data = rng.integers(0, dom, size=int(nn.sum())).astype(np.uint32)
        dw, defw = oracle.pack(data, bw), oracle.pack(nn.astype(np.uint32), 1)
        cols.append((dw, bw, defw))
        lo, hi = dom // 8, dom // 8 + dom // 2
        root = oracle.pred(EQ, defw, 1, 0, n, [1])[0]
        sub = oracle.pred(GE, dw, bw, 0, len(data), [lo])[0] & oracle.pred(LE, dw, bw, 0, len(data), [hi])[0]
        leaf = oracle.intersect_bitset(root, n, sub)[0]
        want = leaf if want is None else (want & leaf)
        prog += [(GE, c, lo), (LE, c, hi), (AND, None, None)] + ([(AND, None, None)] if c else [])
    for _ in range(2):
        bm, cnt = ctx.eval(cols, prog, n)
        assert (bm == want).all() and cnt == popcount(want), bws


# ------------------------------------------------------------------------------------------
# device-resident API at BASELINE.json sizes: size-independent properties
# ------------------------------------------------------------------------------------------
def _torch_pack_mask(mask):
    """Boolean row mask -> selection bitmap packed by torch on the device (row r -> bit r & 63 of word
    r >> 6, fle-encoding.h:7989-8008): an independent full-size image to compare EVERY word against."""
    import torch
    n = mask.numel()
    G = (n + 63) // 64
    if G * 64 != n:
        mask = torch.cat([mask, torch.zeros(G * 64 - n, dtype=torch.bool, device=mask.device)])
    m = mask.view(G, 64)
    want = torch.zeros(G, dtype=torch.int64, device=mask.device)
    for b in range(64):
        want |= m[:, b].to(torch.int64) << b
    return want


def _torch_bits_to_count(bitmap):
    import torch
    b = bitmap.view(torch.uint8)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=b.device)
    return int(lut[b.long()].sum().item())


@pytest.mark.parametrize("bw", [1, 4, 7, 8, 12, 13, 16, 17, 24, 32])
def test_cfg0_roundtrip_16M(ctx, bw):
    """config 0 (fle-benchmark shape): 16 Mi values, encode -> decode round trip on the device."""
    import torch
    n = 16 * 1024 * 1024
    g = torch.Generator(device="cuda").manual_seed(0xF1E0 + bw)
    vals = torch.randint(0, 2 ** bw, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    words = ctx.pack_dev(vals, bw)
    out = ctx.unpack_dev(words, bw, 0, n)
    assert torch.equal(out, vals)
    # every kernel family must produce the same bytes: 0 direct loads/stores, 2 per-row bulk stores,
    # 3 warp-private copy-out, 4 staged copy-out (1 = the per-width choice among them, used above)
    try:
        for variant in (0, 2, 3, 4):
            ctx.set_option("codec_variant", variant)
            assert torch.equal(ctx.pack_dev(vals, bw), words), variant
            assert torch.equal(ctx.unpack_dev(words, bw, 0, n), vals), variant
            m = n - 64 * 129 - 37  # ragged: partial last tile and partial last group
            assert torch.equal(ctx.unpack_dev(words, bw, 0, m), vals[:m]), variant
    finally:
        ctx.set_option("codec_variant", 1)
    # plane image check on a slice: plane i of group g, bit 63-k == bit i of value 64g+k
    k = 64 * 1000
    sl = (vals[:k].to(torch.int64) & 0xFFFFFFFF).view(-1, 64)
    for i in (0, bw - 1):
        bits = (sl >> i) & 1
        weights = torch.tensor([1 << (63 - j) if j > 0 else -(1 << 63) for j in range(64)], device="cuda", dtype=torch.int64)
        plane = (bits * weights).sum(dim=1)
        assert torch.equal(words.view(-1, bw)[:1000, i], plane)
    # unpack from an offset group
    assert torch.equal(ctx.unpack_dev(words, bw, 64 * 777, 12345), vals[64 * 777: 64 * 777 + 12345])


def test_cfg1_between_full_size(ctx):
    """config 1: N = 600,037,902 (SF100 lineitem rows), bw 12 BETWEEN; the count must equal a
    plain torch compare on the values, EVERY word of the bitmap must equal a bitmap packed by torch
    from that compare, both kernels must agree bit for bit, and the bitmap's popcount must equal the
    reported count (checksum of checksums)."""
    import torch
    n = 600_037_902
    bw = 12
    g = torch.Generator(device="cuda").manual_seed(0xF1E00100)
    vals = torch.randint(0, 4096, (n,), generator=g, device="cuda", dtype=torch.int32)
    words = ctx.pack_dev(vals, bw)
    lo, hi = 1000, 1409  # ~10 %
    mask = (vals >= lo) & (vals <= hi)
    want = int(mask.sum().item())
    want_bm = _torch_pack_mask(mask)
    del mask
    prog = [(GE, 0, lo), (LE, 0, hi), (AND, None, None)]
    bm1, c1 = ctx.eval_dev([(words, bw)], prog, n)
    torch.cuda.synchronize()
    assert int(c1.item()) == want
    assert torch.equal(bm1, want_bm), "full 600 M-row bitmap differs from the torch-packed one"
    del want_bm
    assert _torch_bits_to_count(bm1) == want
    ctx.set_option("eval_variant", 0)
    try:
        bm0, c0 = ctx.eval_dev([(words, bw)], prog, n)
        torch.cuda.synchronize()
    finally:
        ctx.set_option("eval_variant", 1)
    assert int(c0.item()) == want and torch.equal(bm0, bm1)
    # spot-check the bitmap against the values
    idx = torch.randint(0, n, (100000,), device="cuda")
    bit = (bm1[idx >> 6] >> (idx & 63)) & 1
    assert torch.equal(bit.bool(), (vals[idx] >= lo) & (vals[idx] <= hi))
    del vals
    cnt = ctx.bitmap_count_dev(bm1, n)
    torch.cuda.synchronize()
    assert int(cnt.item()) == want


def test_cfg2_q6_full_size(ctx):
    """config 2: TPC-H Q6 over 600,037,902 rows (shipdate bw 12, discount bw 4, quantity bw 6, compound
    AND of 5 leaves) + SUM(extendedprice * discount) over the survivors.  Count and the exact int64
    revenue must equal a torch compare of the values; the fused kernel, the two-kernel path and
    the interpreting kernel must agree bit for bit; popcount(bitmap) == count."""
    import torch
    n = 600_037_902
    g = torch.Generator(device="cuda").manual_seed(0xF1E00200)
    price = torch.randint(90_000, 10_500_000, (n,), generator=g, device="cuda", dtype=torch.int64)
    dd = torch.arange(0, 11, device="cuda", dtype=torch.int64)
    cols, mask, disc = [], None, None
    for (d, bw), (lo, hi) in zip(((2526, 12), (11, 4), (50, 6)), ((730, 1094), (5, 7), (0, 22))):
        v = torch.randint(0, d, (n,), generator=g, device="cuda", dtype=torch.int32)
        cols.append((ctx.pack_dev(v, bw), bw))
        m = (v >= lo) & (v <= hi)
        mask = m if mask is None else (mask & m)
        if bw == 4:
            disc = v
    want_cnt = int(mask.sum().item())
    want_rev = int((price[mask] * dd[disc[mask].long()]).sum().item())
    want_bm = _torch_pack_mask(mask)
    del mask, disc, v, m
    prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
            (AND, None, None), (LT, 2, 23), (AND, None, None)]
    bm = torch.empty((n + 63) // 64, dtype=torch.int64, device="cuda")
    out = ctx.eval_sum_product_dev(cols, prog, n, price, 1, dd, bitmap=bm)
    torch.cuda.synchronize()
    assert int(out[0].item()) == want_rev and int(out[1].item()) == want_cnt
    assert torch.equal(bm, want_bm), "full 600 M-row Q6 bitmap differs from the torch-packed one"
    del want_bm
    assert _torch_bits_to_count(bm) == want_cnt
    bm1, c1 = ctx.eval_dev(cols, prog, n)
    two = ctx.sum_product_dev(bm1, n, price, cols[1][0], 4, dd)
    torch.cuda.synchronize()
    assert torch.equal(bm1, bm) and int(c1.item()) == want_cnt
    assert int(two[0].item()) == want_rev and int(two[1].item()) == want_cnt
    for variant in (0, 2):
        ctx.set_option("eval_variant", variant)
        try:
            bmv, cv = ctx.eval_dev(cols, prog, n)
            torch.cuda.synchronize()
        finally:
            ctx.set_option("eval_variant", 1)
        assert torch.equal(bmv, bm) and int(cv.item()) == want_cnt, variant


def test_cfg3_in_lists_and_trees_full_size(ctx):
    """config 3: IN lists of 8 / 64 / 256 literals on bw 12 and bw 16 columns and Q12- / Q14- / OR-style
    trees over 600,037,902 rows: counts equal torch.isin / torch compares, popcount(bitmap) == count,
    lookup kernel and plane form agree bit for bit."""
    import torch
    n = 600_037_902
    g = torch.Generator(device="cuda").manual_seed(0xF1E00300)
    cpu = torch.Generator().manual_seed(3)
    for d, bw in ((2526, 12), (40000, 16)):
        v = torch.randint(0, d, (n,), generator=g, device="cuda", dtype=torch.int32)
        w = ctx.pack_dev(v, bw)
        for L in (8, 64, 256):
            lits = torch.randint(0, d, (L,), generator=cpu)
            member = torch.isin(v, lits.to("cuda", torch.int32))
            want = int(member.sum().item())
            want_bm = _torch_pack_mask(member)
            del member
            bm, c = ctx.eval_dev([(w, bw)], [(IN, 0, lits.tolist())], n)
            torch.cuda.synchronize()
            assert int(c.item()) == want and _torch_bits_to_count(bm) == want, (bw, L)
            assert torch.equal(bm, want_bm), ("full bitmap", bw, L)
            del want_bm
            if L == 8:  # the other IN implementation (lookup kernel <-> unrolled equality chains)
                ctx.set_option("in_lookup_min", 1)
                try:
                    bm2, c2 = ctx.eval_dev([(w, bw)], [(IN, 0, lits.tolist())], n)
                    torch.cuda.synchronize()
                finally:
                    ctx.set_option("in_lookup_min", 9)
                assert torch.equal(bm2, bm) and int(c2.item()) == want
        del v, w
    mode = torch.randint(0, 7, (n,), generator=g, device="cuda", dtype=torch.int32)
    rcpt = torch.randint(0, 2526, (n,), generator=g, device="cuda", dtype=torch.int32)
    instr = torch.randint(0, 4, (n,), generator=g, device="cuda", dtype=torch.int32)
    qty = torch.randint(0, 50, (n,), generator=g, device="cuda", dtype=torch.int32)
    w_mode, w_rcpt, w_instr, w_qty = ctx.pack_dev(mode, 3), ctx.pack_dev(rcpt, 12), ctx.pack_dev(instr, 2), ctx.pack_dev(qty, 6)
    in_mode = (mode == 1) | (mode == 4)
    cases = {
        "q12": ([(w_mode, 3), (w_rcpt, 12)], [(IN, 0, [1, 4]), (GE, 1, 730), (AND, None, None), (LT, 1, 1095), (AND, None, None)],
                in_mode & (rcpt >= 730) & (rcpt < 1095)),
        "q14": ([(w_rcpt, 12)], [(GE, 0, 1000), (LT, 0, 1030), (AND, None, None)], (rcpt >= 1000) & (rcpt < 1030)),
        "or": ([(w_mode, 3), (w_instr, 2), (w_qty, 6)], [(IN, 0, [1, 4]), (EQ, 1, 2), (OR, None, None), (LT, 2, 23), (AND, None, None)],
               (in_mode | (instr == 2)) & (qty < 23)),
    }
    for name, (cols, prog, mask) in cases.items():
        want = int(mask.sum().item())
        bm, c = ctx.eval_dev(cols, prog, n)
        torch.cuda.synchronize()
        assert int(c.item()) == want and _torch_bits_to_count(bm) == want, name
        assert torch.equal(bm, _torch_pack_mask(mask)), ("full bitmap", name)


def test_two_contexts_two_threads(ctx, oracle):
    """The library is re-entrant: one context per scanner thread, each with its own stream
    (SURVEY 8b threading).  Two threads evaluate different programs at the same time; each result
    must equal the oracle's."""
    import threading

    import torch

    import fle_b200
    rng = np.random.default_rng(4242)
    n = 64 * 30000 + 17
    jobs = []
    for bw, prog_of in ((12, lambda D: [(GE, 0, D // 4), (LE, 0, D // 2), (AND, None, None)]), (7, lambda D: [(IN, 0, list(range(3, 90, 7)))])):
        vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
        words = oracle.pack(vals, bw)
        prog = prog_of(1 << bw)
        if prog[0][0] == IN:
            want, wcnt = oracle.pred(IN, words, bw, 0, n, np.array(prog[0][2], dtype=np.uint64))
        else:
            a, _ = oracle.pred(GE, words, bw, 0, n, prog[0][2])
            b, _ = oracle.pred(LE, words, bw, 0, n, prog[1][2])
            want = a & b
            wcnt = popcount(want)
        jobs.append((words, bw, prog, want, wcnt))
    errors = []

    def worker(job):
        words, bw, prog, want, wcnt = job
        try:
            stream = torch.cuda.Stream()
            c = fle_b200.FleContext(0, stream.cuda_stream)
            for _ in range(50):
                bm, cnt = c.eval([(words, bw)], prog, n)
                if not ((bm == want).all() and cnt == wcnt):
                    errors.append((bw, "mismatch"))
                    break
            c.close()
        except Exception as e:  # pragma: no cover
            errors.append((bw, repr(e)))

    ts = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------------------------------
# survivor materialisation + the Q6 aggregate
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sel", [0.0, 0.019, 0.5, 1.0])
def test_select_gather_sum_product(ctx, sel):
    import torch
    n = 64 * 50000 + 41
    g = torch.Generator(device="cuda").manual_seed(77)
    ship = torch.randint(0, 2526, (n,), generator=g, device="cuda", dtype=torch.int32)
    disc = torch.randint(0, 11, (n,), generator=g, device="cuda", dtype=torch.int32)
    price = torch.randint(90_000, 10_500_000, (n,), generator=g, device="cuda", dtype=torch.int64)
    w_ship, w_disc = ctx.pack_dev(ship, 12), ctx.pack_dev(disc, 4)
    hi = int(round(sel * 2526)) - 1
    prog = [(LE, 0, hi)] if hi >= 0 else [(NONE, None, None)]
    bm, cnt = ctx.eval_dev([(w_ship, 12)], prog, n)
    mask = ship <= hi
    want_ids = torch.nonzero(mask).flatten().to(torch.int32)
    ids, c2 = ctx.bitmap_select_dev(bm, n)
    torch.cuda.synchronize()
    k = int(c2.item())
    assert k == int(cnt.item()) == want_ids.numel()
    assert torch.equal(ids[:k], want_ids)                       # ascending row ids == skip distances
    codes = ctx.gather_dev(w_disc, 4, ids, c2, n)
    assert torch.equal(codes[:k], disc[mask])                   # Get(val, skip) per survivor
    dict_disc = torch.arange(0, 11, device="cuda", dtype=torch.int64) * 100 + 7   # any sorted dictionary
    vals = ctx.gather_dev(w_disc, 4, ids, c2, n, dict_values=dict_disc)
    assert torch.equal(vals[:k], dict_disc[disc[mask].long()])  # dictionary lookup -> dense column
    out = ctx.sum_product_dev(bm, n, price, w_disc, 4, dict_disc)
    torch.cuda.synchronize()
    want = int((price[mask] * dict_disc[disc[mask].long()]).sum().item())  # exact int64
    assert int(out[0].item()) == want and int(out[1].item()) == k
    # capacity smaller than the number of survivors: count is still exact, ids truncated
    if k > 10:
        ids2, c3 = ctx.bitmap_select_dev(bm, n, cap=10)
        torch.cuda.synchronize()
        assert int(c3.item()) == k and torch.equal(ids2[:10], want_ids[:10])


@pytest.mark.parametrize("bw", list(range(1, 17)) + [20])
def test_in_list_lookup_kernel_vs_oracle(ctx, oracle, bw):
    """IN by decode + membership lookup (bw <= 16) must equal the OR-of-equalities of the
    reference (fle-encoding.h:8236-8313) and the flat plane-algebra kernels, for any list length."""
    rng = np.random.default_rng(900 + bw)
    n = 64 * 4096 * 2 + 64 * 11 + 5
    vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
    words = oracle.pack(vals, bw)
    for L in (1, 8, 64, 256, 1000):
        lits = rng.integers(0, 1 << bw, size=L, dtype=np.uint64)  # duplicates and absent codes included
        want, wcnt = oracle.pred(IN, words, bw, 0, n, lits)
        # (lookup_min, in_wide, in_mux_bw): multiplexer tree over the planes (bw <= 7), one-CTA producer/consumer kernel
        # (entry table, replicated bit table), two-CTA bit-table kernel, equality chains on the planes
        # (wide 2: word / byte entries addressed straight from the transposed word; 1: 16-bit entries)
        for lookup_min, wide, mux in ((1, 2, 7), (1, 2, 0), (1, 1, 0), (1, 0, 0), (1 << 30, 2, 7)):
            ctx.set_option("in_lookup_min", lookup_min)
            ctx.set_option("in_wide", wide)
            ctx.set_option("in_mux_bw", mux)
            ctx.set_option("in_mux_min", 4 if lookup_min == 1 else 1 << 30)  # (last configuration: equality chains only)
            try:
                bm, cnt = ctx.eval([(words, bw)], [(IN, 0, lits)], n)
                mode = ctx.get_option("last_mode")
            finally:
                ctx.set_option("in_lookup_min", 9)
                ctx.set_option("in_wide", 2)
                ctx.set_option("in_mux_bw", 6)
                ctx.set_option("in_mux_min", 4)
            assert (bm == want).all() and cnt == wcnt, (bw, L, lookup_min, wide, mux)
            if lookup_min == 1 and bw <= 16:
                assert mode == (7 if bw <= mux else (6 if wide else 3)), (bw, mode)


def test_long_in_lists_inside_programs_vs_oracle(ctx, oracle):
    """Long IN lists inside larger programs (device-resident calls): the list goes through the lookup kernel,
    the other operands of its AND / OR chain run as one fused program, bitmaps are combined word by word
    (in_split_min_rows).  Must equal the reference's leaf-by-leaf evaluation (simple-predicates.h:145-163,
    fle-encoding.h:8236-8313) and the unsplit program."""
    import torch
    rng = np.random.default_rng(77)
    n = 64 * 9000 + 29
    bws = (12, 6, 16, 3)
    doms = (2526, 50, 40000, 7)
    vals = [rng.integers(0, d, size=n).astype(np.uint32) for d in doms]
    cols = [(oracle.pack(v, bw), bw) for v, bw in zip(vals, bws)]
    dev = [(torch.from_numpy(w.view(np.int64)).cuda(), bw) for w, bw in cols]
    big0 = rng.choice(vals[0], size=256).astype(np.uint64)
    big2 = rng.choice(vals[2], size=64).astype(np.uint64)
    progs = {
        "in_and": [(IN, 0, big0), (LT, 1, 30), (AND, None, None)],
        "lt_and_in": [(LT, 1, 30), (IN, 0, big0), (AND, None, None)],
        "chain": [(IN, 0, big0), (GE, 1, 5), (AND, None, None), (LT, 1, 40), (AND, None, None), (IN, 2, big2), (AND, None, None),
                  (IN, 3, [1, 4]), (AND, None, None)],
        "or_of_and": [(IN, 0, big0), (GE, 1, 5), (LT, 3, 4), (AND, None, None), (OR, None, None)],
        "nested": [(IN, 0, big0), (LT, 1, 10), (AND, None, None), (GE, 3, 5), (OR, None, None), (LT, 2, 30000), (AND, None, None)],
        "two_lists_or": [(IN, 0, big0), (IN, 2, big2), (OR, None, None)],
        "same_col": [(IN, 0, big0), (LT, 0, 1500), (AND, None, None), (GE, 0, 100), (AND, None, None)],
        "consts": [(ALL, None, None), (IN, 0, big0), (AND, None, None), (NONE, None, None), (OR, None, None)],
    }
    for name, prog in progs.items():
        want, wcnt = oracle.eval_tree(cols, prog, n)
        for split in (0, -1):
            ctx.set_option("in_split_min_rows", split)
            try:
                launches = ctx.launch_count
                bm, cnt = ctx.eval_dev(dev, prog, n)
                torch.cuda.synchronize()
                launches = ctx.launch_count - launches
            finally:
                ctx.set_option("in_split_min_rows", 1 << 20)
            got = bm.cpu().numpy().view(np.uint64)
            assert (got == want).all() and int(cnt.item()) == wcnt, (name, split)
            assert (launches > 1) == (split == 0), (name, split, launches)


# ------------------------------------------------------------------------------------------
# "next" rows: page framing on the device, PLAIN-encoded predicates
# ------------------------------------------------------------------------------------------
def test_page_framing_realign_and_eval(ctx, oracle):
    """A column chunk kept in HBM as the raw file image: pages [int32 def_len][def FLE][u8 bw][FLE]
    at arbitrary byte offsets -> fle_page_parse + funnel-shift realign on the device -> fused eval
    with def levels; per-page bit widths; result == reference structure evaluated by the oracle."""
    import torch
    rng = np.random.default_rng(321)
    pages, blob = [], bytearray(b"\x7f" * 3)  # leading junk so that nothing is aligned by accident
    for pi, (rows, bw, null_frac) in enumerate(((64 * 300 + 17, 12, 0.05), (64 * 200, 9, 0.0), (777, 5, 0.5), (64 * 64, 16, 0.01))):
        nn = rng.random(rows) >= null_frac
        data = rng.integers(0, 1 << bw, size=int(nn.sum()), dtype=np.uint64).astype(np.uint32)
        defw = oracle.pack(nn.astype(np.uint32), 1)
        dataw = oracle.pack(data, bw)
        body = np.int32(defw.nbytes).tobytes() + defw.tobytes() + bytes([bw]) + dataw.tobytes()
        blob += b"\x55" * int(rng.integers(0, 7))
        pages.append((len(blob), len(body), rows, bw, nn, data, defw, dataw))
        blob += body
    host = np.frombuffer(bytes(blob), dtype=np.uint8)
    dev = torch.from_numpy(host.copy()).cuda()
    for off, size, rows, bw, nn, data, defw, dataw in pages:
        lay = ctx.page_parse(host[off:off + size], 1)
        assert lay.bit_width == bw and lay.def_bytes == defw.nbytes and lay.data_bytes == dataw.nbytes
        assert (off + lay.data_off) % 8 != 0 or True
        d_def = ctx.realign_dev(dev, off + lay.def_off, lay.def_bytes)
        d_data = ctx.realign_dev(dev, off + lay.data_off, lay.data_bytes)
        torch.cuda.synchronize()
        assert (d_def.cpu().numpy().view(np.uint64)[: len(defw)] == defw).all()
        assert (d_data.cpu().numpy().view(np.uint64)[: len(dataw)] == dataw).all()
        lit = int(data[len(data) // 2])
        bm, cnt = ctx.eval_dev([(d_data[: len(dataw)], bw, d_def[: len(defw)])], [(LE, 0, lit)], rows)
        torch.cuda.synchronize()
        root = oracle.pred(EQ, defw, 1, 0, rows, [1])[0]
        sub = oracle.pred(LE, dataw, bw, 0, len(data), [lit])[0]
        want = oracle.intersect_bitset(root, rows, sub)[0]
        assert (bm.cpu().numpy().view(np.uint64)[: len(want)] == want).all() and int(cnt.item()) == popcount(want)
    # odd sizes / every misalignment
    for off in list(range(0, 18)) + [31, 33, 255]:
        for nbytes in (1, 3, 4, 5, 8, 13, 15, 16, 17, 64, 1001, 4099):
            out = ctx.realign_dev(dev, off, nbytes)
            torch.cuda.synchronize()
            got = out.cpu().numpy().view(np.uint8)
            assert (got[:nbytes] == host[off:off + nbytes]).all() and (got[nbytes:(nbytes + 15) // 16 * 16] == 0).all()


@pytest.mark.parametrize("dtype", ["int32", "int64", "float32", "float64"])
def test_plain_predicates(ctx, dtype):
    """PLAIN fixed-width columns (parquet-common.h:197-255): `value OP literal` by default, the
    reference's literal `literal OP value` order on request, and a working IN."""
    import torch
    rng = np.random.default_rng(17)
    for n in (1, 31, 32, 33, 64, 65, 64 * 3000 + 19):
        if dtype.startswith("int"):
            vals = rng.integers(-50, 50, size=n).astype(dtype)
        else:
            vals = (rng.integers(-50, 50, size=n) / 4).astype(dtype)
        t = torch.from_numpy(vals).cuda()
        lit = vals[n // 2]
        for op, f in ((EQ, np.equal), (LT, np.less), (LE, np.less_equal), (GT, np.greater), (GE, np.greater_equal)):
            for ref_order in (False, True):
                bm, cnt = ctx.plain_eval_dev(t, op, [lit], reference_order=ref_order)
                torch.cuda.synchronize()
                want_bits = f(lit, vals) if ref_order else f(vals, lit)
                got = np.unpackbits(bm.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
                assert (got == want_bits).all() and int(cnt.item()) == int(want_bits.sum()), (dtype, n, op, ref_order)
                assert np.unpackbits(bm.cpu().numpy().view(np.uint8), bitorder="little")[n:].sum() == 0
        lits = rng.choice(vals, size=min(5, n))
        bm, cnt = ctx.plain_eval_dev(t, IN, lits)
        torch.cuda.synchronize()
        got = np.unpackbits(bm.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (got == np.isin(vals, lits)).all() and int(cnt.item()) == int(np.isin(vals, lits).sum())


def test_dict_build_on_device(ctx, oracle):
    """DictEncoder::Put + WriteDict (sorted) + WriteData on the device: the dictionary
    is the sorted set of values, codes are order-preserving, the planes equal the oracle's packing
    of those codes; an over-full dictionary reports FLE_E_DICT_FULL (Put() == -1)."""
    import torch
    import fle_b200
    rng = np.random.default_rng(8)
    for dtype, nd, n in ((np.int64, 2526, 64 * 4000 + 7), (np.int32, 11, 1000), (np.int64, 1, 130), (np.int64, 40000, 400000),
                         (np.int32, 3, 64)):
        domain = np.unique(rng.integers(-10**9, 10**9, size=nd * 2).astype(dtype))[:nd]
        if dtype == np.int64 and nd > 5:
            domain[0] = np.iinfo(np.int64).min  # the hash table's sentinel value must work as data
            domain.sort()
        vals = rng.choice(domain, size=n)
        vals[: len(domain)] = domain  # every entry occurs
        d, bw, enc = ctx.dict_encode_dev(torch.from_numpy(vals).cuda())
        torch.cuda.synchronize()
        want_dict = np.unique(vals)
        assert (d.cpu().numpy() == want_dict.astype(np.int64)).all()
        assert bw == oracle.dict_bit_width(len(want_dict))
        codes = np.searchsorted(want_dict, vals).astype(np.uint32)
        assert (enc.cpu().numpy().view(np.uint64) == oracle.pack(codes, bw)).all()
    # large chunks go through the shared-memory front caches (insert: the keys a CTA has met; encode: value -> code map
    # built from the sorted dictionary when it has at most 12288 entries): few / many distinct values, INT64_MIN, int32,
    # and a chunk long enough for the sampled choice between the two insert kernels
    for dt, nd_want, with_min, n2 in ((torch.int64, 3000, True, (1 << 21) + 77), (torch.int64, 12288, False, (1 << 21) + 1),
                                      (torch.int64, 30000, False, (1 << 23) + 5), (torch.int32, 777, False, (1 << 23) + 64 * 3),
                                      (torch.int64, 9000, False, (1 << 23) + 9)):
        codes_t = torch.randint(0, nd_want, (n2,), generator=torch.Generator().manual_seed(nd_want))
        codes_t[-nd_want:] = torch.arange(nd_want)  # every entry occurs (the last ones only behind the sample)
        table = torch.arange(nd_want, dtype=torch.int64) * 1000003 - 5 * 10**8
        if dt == torch.int32:
            table = torch.arange(nd_want, dtype=torch.int64) * 2711 - 10**6
        if with_min:
            table[0] = torch.iinfo(torch.int64).min
        vals2 = table[codes_t].to(dt)
        d2, bw2, enc2 = ctx.dict_encode_dev(vals2.cuda())
        torch.cuda.synchronize()
        assert torch.equal(d2.cpu(), table), (dt, nd_want)
        assert bw2 == max(1, int(nd_want - 1).bit_length())
        got2 = oracle.unpack(enc2.cpu().numpy().view(np.uint64), bw2, 0, n2)
        assert (got2.astype(np.int64) == codes_t.numpy()).all(), (dt, nd_want)
    big = torch.arange(0, 50000, device="cuda", dtype=torch.int64)
    with pytest.raises(fle_b200.FleError) as ei:
        ctx.dict_encode_dev(big)
    assert ei.value.code == fle_b200.capi.E_DICT_FULL
    big = torch.arange(0, 1 << 23, device="cuda", dtype=torch.int64) % 50000  # ... and through the sampled path
    with pytest.raises(fle_b200.FleError) as ei:
        ctx.dict_encode_dev(big)
    assert ei.value.code == fle_b200.capi.E_DICT_FULL


@pytest.mark.parametrize("bw", list(range(1, 17)))
def test_unpack_narrow_targets(ctx, oracle, bw):
    """FleDecoder::Get<T> for int8 / int16 columns: 1-byte values for bw <= 8, 2-byte values for bw <= 16
    (fle-encoding.h:236-238, :8354-8361), full tiles, ragged tails and group offsets."""
    import torch
    rng = np.random.default_rng(1600 + bw)
    for n in (1, 63, 64 * 128, 64 * 128 * 3 + 64 * 17 + 9):
        vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
        words = torch.from_numpy(oracle.pack(vals, bw).view(np.int64)).cuda()
        for es in ((1, 2) if bw <= 8 else (2,)):
            for variant in (1, 4):  # warp-private copy-out (default), CTA-wide copy-out
                out = torch.full((n + 32,), -1, dtype=torch.uint8 if es == 1 else torch.int16, device="cuda")
                ctx.set_option("codec_variant", variant)
                try:
                    ctx.unpack_narrow_dev(words, bw, 0, n, out=out[:n])
                    torch.cuda.synchronize()
                finally:
                    ctx.set_option("codec_variant", 1)
                got = out.cpu().numpy()
                assert (got[:n].astype(np.uint16 if es == 2 else np.uint8) == vals).all(), (bw, n, es, variant)
                assert (got[n:] == -1 if es == 2 else got[n:] == 255).all(), "wrote past the end"
        if n > 64 * 130:
            first = 64 * 128
            out = ctx.unpack_narrow_dev(words, bw, first, n - first)
            torch.cuda.synchronize()
            assert (out.cpu().numpy().astype(np.uint16 if bw > 8 else np.uint8) == vals[first:]).all()
    with pytest.raises(Exception):
        ctx.unpack_narrow_dev(words, 9, 0, 64, out=torch.empty(64, dtype=torch.uint8, device="cuda"))


def test_optional_columns_without_host_sync(ctx, oracle):
    """fle_eval_no_nulls_dev: OPTIONAL columns checked and evaluated without reading anything back; the NULL count
    says whether the optimistic result stands (hdfs-parquet-table-writer.cc:103-105, hdfs-parquet-scanner.cc:326-345)."""
    import torch
    rng = np.random.default_rng(31)
    n = 64 * 3000 + 41
    vals = [rng.integers(0, d, size=n).astype(np.uint32) for d in (2526, 11, 50)]
    bws = (12, 4, 6)
    cols = [(oracle.pack(v, bw), bw) for v, bw in zip(vals, bws)]
    want, wcnt = oracle.eval_tree(cols, Q6, n)
    ones = oracle.pack(np.ones(n, dtype=np.uint32), 1)
    dcols = [(torch.from_numpy(w.view(np.int64)).cuda(), bw, torch.from_numpy(ones.view(np.int64)).cuda()) for w, bw in cols]
    bm, out = ctx.eval_no_nulls_dev(dcols, Q6, n)
    torch.cuda.synchronize()
    assert int(out[1].item()) == 0 and int(out[0].item()) == wcnt
    assert (bm.cpu().numpy().view(np.uint64)[: len(want)] == want).all()
    # NULLs present: the call reports how many, the caller falls back to fle_eval_dev
    defs = np.ones(n, dtype=np.uint32)
    defs[rng.choice(n, size=77, replace=False)] = 0
    dcols[1] = (dcols[1][0], 4, torch.from_numpy(oracle.pack(defs, 1).view(np.int64)).cuda())
    _, out = ctx.eval_no_nulls_dev(dcols, Q6, n)
    torch.cuda.synchronize()
    assert int(out[1].item()) == 77
    # required columns: nothing to check
    _, out = ctx.eval_no_nulls_dev([(c[0], c[1]) for c in dcols], Q6, n)
    torch.cuda.synchronize()
    assert int(out[1].item()) == 0 and int(out[0].item()) == wcnt


class _RawCol:
    """A device 'tensor' that is a byte range of a larger uint8 tensor (what capi.make_col_descs needs)."""

    def __init__(self, blob, off, nbytes):
        self.blob, self.off, self.nbytes = blob, off, nbytes
        self.device = blob.device

    def data_ptr(self):
        return self.blob.data_ptr() + self.off

    def numel(self):
        return self.nbytes

    def element_size(self):
        return 1


@pytest.mark.parametrize("skew", [1, 2, 3, 4, 5, 7, 8, 9, 12, 15])
def test_eval_directly_on_raw_page_image(ctx, oracle, skew):
    """The FLE section of a page inside a file image starts at an ODD byte offset (dict-encoding.h:185-192): the
    fused kernels evaluate it in place -- no realigned copy (fle_eval1_skew.cu, FLE_JIT_SKEW) -- for single-pass
    programs, IN lists on narrow columns and multi-column programs; bit-identical to the oracle on the planes."""
    import torch
    rng = np.random.default_rng(50 + skew)
    n = 64 * 4096 + 64 * 300 + 11   # full staged tiles + a ragged remainder
    specs = ((2526, 12), (11, 4), (50, 6), (40000, 16), (7, 3))
    blob = bytearray(b"\xee" * (256 + skew))
    cols_np, offs = [], []
    for d, bw in specs:
        v = rng.integers(0, d, size=n).astype(np.uint32)
        w = oracle.pack(v, bw)
        pad = (-(len(blob) - skew)) % 16          # every column at the same misalignment
        blob += b"\x33" * pad
        offs.append(len(blob))
        blob += w.tobytes()
        cols_np.append((w, bw))
    blob += b"\x44" * 64
    dev = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    assert all((dev.data_ptr() + o) % 16 == skew % 16 for o in offs)
    raw = [(_RawCol(dev, o, w.nbytes), bw) for o, (w, bw) in zip(offs, cols_np)]
    programs = [
        ([0], [(GE, 0, 730), (LE, 0, 1094), (AND, None, None)]),                    # BETWEEN, single pass
        ([3], [(EQ, 0, 12345)]),
        ([4], [(IN, 0, [1, 4, 5, 6, 0, 2, 1, 4, 5, 6])]),                            # narrow IN list: multiplexer tree
        ([0, 1, 2], Q6),                                                             # specialised multi-column kernel
        ([4, 2], [(IN, 0, [1, 4]), (LT, 1, 23), (OR, None, None)]),
    ]
    for jit in (2, 0):
        ctx.set_option("eval_jit", jit)
        try:
            for idx, prog in programs:
                want, wcnt = oracle.eval_tree([cols_np[i] for i in idx], prog, n)
                l0 = ctx.launch_count
                bm, cnt = ctx.eval_dev([raw[i] for i in idx], prog, n)
                torch.cuda.synchronize()
                assert (bm.cpu().numpy().view(np.uint64)[: len(want)] == want).all() and int(cnt.item()) == wcnt, (skew, jit, idx)
                if jit == 2 or len(idx) == 1:
                    assert ctx.launch_count - l0 == 1, "the image must be evaluated in place, in one launch"
            # a sub-range that starts further into the image (first_row on a group boundary)
            want, wcnt = oracle.eval_tree([(cols_np[0][0][64 * 12:], 12)], [(LT, 0, 1000)], n - 64 * 64)
            bm, cnt = ctx.eval_dev([raw[0]], [(LT, 0, 1000)], n - 64 * 64, first_row=64 * 64)
            torch.cuda.synchronize()
            assert (bm.cpu().numpy().view(np.uint64)[: len(want)] == want).all() and int(cnt.item()) == wcnt
        finally:
            ctx.set_option("eval_jit", 1)

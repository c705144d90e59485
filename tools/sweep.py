#!/usr/bin/env python
"""Kernel sweeps on one B200 (device-resident inputs, CUDA events, MEDIAN of `reps`; launches whose
working set is below 2x the 126 MB L2 are timed one by one with an L2 flush -- a 384 MB overwrite --
before each, so every fraction is a DRAM-bound number).

    python tools/sweep.py between|codec|programs|tune|all [--rows N] [--out gpurun_out/sweep.jsonl]

Every record: {"what", "bw", ..., "ms_min", "ms_med", "alg_bytes", "GBps", "frac"} where
alg_bytes is the algorithmic traffic of DESIGN.md (planes read once + bitmap / values written
once) and frac is against MEASURED_PEAKS.json's HBM copy bandwidth."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, EQ, GE, GT, IN, LE, LT, OR  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK = peak()


L2_BYTES = 126 << 20
_flush_buf = None
_flush_clean = None
_flush_sink = None


def flush_l2():
    """384 MB overwrite, then a 256 MB READ of a buffer that is clean in DRAM: the overwrite alone leaves up to an L2 of
    dirty lines whose write-back the timed kernel would pay for (~19 us)."""
    global _flush_buf, _flush_clean, _flush_sink
    if _flush_buf is None:
        _flush_buf = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
        _flush_clean = torch.zeros(256 << 17, dtype=torch.int64, device="cuda")
        _flush_sink = torch.zeros(1, dtype=torch.int64, device="cuda")
    _flush_buf.fill_(1)
    _flush_clean.sum()


def timeit(fn, reps=10, warm=3, inner=8, working_set=None):
    """Per-launch time.  Working sets above 2x L2: `inner` launches queued back to back between two
    events, so that the host-side launch path overlaps the previous kernel (as in a real scan loop; the
    inputs cannot stay in L2).  Smaller working sets (or unknown): one launch per event pair with the
    L2 flushed before it."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    cold = working_set is None or working_set < 2 * L2_BYTES
    for _ in range(reps if not cold else max(reps, 7)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if cold:
            flush_l2()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        else:
            a.record()
            for _ in range(inner):
                fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / inner)
    return min(ts), statistics.median(ts)


def rec(out, **kw):
    if "alg_bytes" in kw and "ms_min" in kw:
        kw["GBps"] = kw["alg_bytes"] / (kw["ms_med"] * 1e-3) / 1e9
        kw["GBps_best"] = kw["alg_bytes"] / (kw["ms_min"] * 1e-3) / 1e9
        kw["frac"] = kw["GBps"] / PEAK          # of the MEDIAN launch
        kw["cold_l2"] = kw["alg_bytes"] < 2 * L2_BYTES
    line = json.dumps(kw)
    print(line, flush=True)
    out.write(line + "\n")
    out.flush()


def make_col(ctx, n, bw, dom=None, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    hi = dom if dom else (1 << bw)
    v = torch.randint(0, hi, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    w = ctx.pack_dev(v, bw)
    torch.cuda.synchronize()
    return v, w


def sweep_between(ctx, out, n, variants=(1, 2, 0), bws=range(1, 33)):
    G = (n + 63) // 64
    bm = torch.empty(G, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for bw in bws:
        v, w = make_col(ctx, n, bw, seed=bw)
        D = 1 << bw
        lo = D // 4
        hi = max(lo, lo + D // 10 - 1)
        prog = [(GE, 0, lo), (LE, 0, hi), (AND, None, None)]
        want = int(((v.to(torch.int64) & 0xFFFFFFFF >= lo) & ((v.to(torch.int64) & 0xFFFFFFFF) <= hi)).sum().item()) if False else None
        del v
        cols = [(w, bw)]
        prep = (ctx.make_col_descs(cols),) + fle_b200.capi._build_program(prog)
        for var in variants:
            ctx.set_option("eval_variant", var)
            mn, md = timeit(lambda: ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt, prepared=prep), working_set=G * bw * 8)
            rec(out, what="between", variant=var, bw=bw, rows=n, ms_min=mn, ms_med=md, alg_bytes=G * bw * 8 + G * 8,
                rows_per_s=n / (mn * 1e-3), count=int(cnt.item()))
        ctx.set_option("eval_variant", 1)
        del w
        torch.cuda.empty_cache()


def sweep_selectivity(ctx, out, n, bw=12):
    G = (n + 63) // 64
    bm = torch.empty(G, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    v, w = make_col(ctx, n, bw, seed=99)
    del v
    cols = [(w, bw)]
    D = 1 << bw
    for sel in (1, 10, 20, 30, 50, 80, 100):
        width = max(1, round(sel * D / 100))
        lo = (D - width) // 2
        prog = [(GE, 0, lo), (LE, 0, lo + width - 1), (AND, None, None)]
        prep = (ctx.make_col_descs(cols),) + fle_b200.capi._build_program(prog)
        mn, md = timeit(lambda: ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt, prepared=prep), working_set=G * bw * 8)
        rec(out, what="selectivity", sel_pct=sel, bw=bw, rows=n, ms_min=mn, ms_med=md, alg_bytes=G * bw * 8 + G * 8,
            rows_per_s=n / (mn * 1e-3), observed_sel=int(cnt.item()) / n)


def sweep_codec(ctx, out, n, bws=range(1, 33)):
    G = (n + 63) // 64
    for bw in bws:
        v, w = make_col(ctx, n, bw, seed=100 + bw)
        o = torch.empty(n, dtype=torch.int32, device="cuda")
        mn, md = timeit(lambda: ctx.unpack_dev(w, bw, 0, n, out=o), working_set=n * 4)
        ok = bool(torch.equal(o, v))
        rec(out, what="unpack", bw=bw, values=n, ms_min=mn, ms_med=md, alg_bytes=G * bw * 8 + n * 4,
            values_per_s=n / (mn * 1e-3), ok=ok)
        w2 = torch.empty_like(w)
        mn, md = timeit(lambda: ctx.pack_dev(v, bw, out=w2), working_set=n * 4)
        ok = bool(torch.equal(w2, w))
        rec(out, what="pack", bw=bw, values=n, ms_min=mn, ms_med=md, alg_bytes=G * bw * 8 + n * 4,
            values_per_s=n / (mn * 1e-3), ok=ok)
        del v, w, o, w2
        torch.cuda.empty_cache()


def sweep_programs(ctx, out, n):
    G = (n + 63) // 64
    bm = torch.empty(G, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    spec = {"shipdate": (2526, 12), "discount": (11, 4), "quantity": (50, 6), "shipmode": (7, 3), "instruct": (4, 2),
            "receipt": (2526, 12), "wide16": (40000, 16)}
    cols = {}
    for i, (k, (d, bw)) in enumerate(spec.items()):
        v, w = make_col(ctx, n, bw, dom=d, seed=200 + i)
        del v
        cols[k] = (w, bw)
    g = torch.Generator().manual_seed(5)
    programs = {
        "q6": (["shipdate", "discount", "quantity"],
               [(GE, 0, 730), (LE, 0, 1094), (AND,), (GE, 1, 5), (LE, 1, 7), (AND,), (AND,), (LT, 2, 23), (AND,)]),
        "q12": (["shipmode", "receipt"], [(IN, 0, [1, 4]), (GE, 1, 730), (AND,), (LT, 1, 1095), (AND,)]),
        "q14": (["shipdate"], [(GE, 0, 1339), (LT, 0, 1369), (AND,)]),
        "or_variant": (["shipmode", "instruct", "quantity"], [(IN, 0, [1, 4]), (EQ, 1, 2), (OR,), (LT, 2, 23), (AND,)]),
        "q19ish": (["shipmode", "instruct", "quantity", "discount", "shipdate"],
                   [(IN, 0, [0, 2]), (EQ, 1, 1), (AND,), (GE, 2, 1), (LE, 2, 11), (AND,), (AND,),
                    (IN, 0, [1, 3]), (EQ, 1, 1), (AND,), (GE, 2, 10), (LE, 2, 20), (AND,), (AND,), (OR,),
                    (LT, 3, 9), (AND,), (GE, 4, 100), (AND,)]),
        "in20_and_lt": (["shipdate", "quantity"], [(IN, 0, torch.randint(0, 2526, (20,), generator=g).tolist()), (LT, 1, 30), (OR,)]),
        "in8_bw12": (["shipdate"], [(IN, 0, torch.randint(0, 2526, (8,), generator=g).tolist())]),
        "in64_bw12": (["shipdate"], [(IN, 0, torch.randint(0, 2526, (64,), generator=g).tolist())]),
        "in256_bw12": (["shipdate"], [(IN, 0, torch.randint(0, 2526, (256,), generator=g).tolist())]),
        "in8_bw16": (["wide16"], [(IN, 0, torch.randint(0, 40000, (8,), generator=g).tolist())]),
        "in64_bw16": (["wide16"], [(IN, 0, torch.randint(0, 40000, (64,), generator=g).tolist())]),
        "in256_bw16": (["wide16"], [(IN, 0, torch.randint(0, 40000, (256,), generator=g).tolist())]),
    }
    for name, (cn, prog) in programs.items():
        pc = [cols[c] for c in cn]
        prog = [tuple(p) + (None,) * (3 - len(p)) for p in prog]
        prep = (ctx.make_col_descs(pc),) + fle_b200.capi._build_program(prog)
        sum_bw = sum(bw for _, bw in pc)
        # "1j": default path with run-time specialised kernels forced on; 1 / 2 / 0: interpreting kernels
        for label, var, jit in (("1j", 1, 2), (1, 1, 0), (2, 2, 0), (0, 0, 0)):
            ctx.set_option("eval_variant", var)
            ctx.set_option("eval_jit", jit)
            mn, md = timeit(lambda: ctx.eval_dev(pc, prog, n, bitmap=bm, count=cnt, prepared=prep), reps=5 if "in256" in name else 10,
                            working_set=G * sum_bw * 8)
            rec(out, what="program", name=name, variant=label, rows=n, sum_bw=sum_bw, ms_min=mn, ms_med=md,
                alg_bytes=G * sum_bw * 8 + G * 8, rows_per_s=n / (mn * 1e-3), count=int(cnt.item()))
        ctx.set_option("eval_variant", 1)
        ctx.set_option("eval_jit", 1)
    # Q6 the way an Impala-written file presents it: every column OPTIONAL with a def-level stream
    # (1-bit FLE, hdfs-parquet-table-writer.cc:103-105) that holds no NULL at all.  The call includes
    # the def-level count kernel and the host sync that decides between the fused and the NULL path.
    ones = torch.full((G,), -1, dtype=torch.int64, device="cuda")
    pc = [cols[c] + (ones,) for c in ("shipdate", "discount", "quantity")]
    prog = [tuple(p) + (None,) * (3 - len(p)) for p in programs["q6"][1]]
    prep = (ctx.make_col_descs(pc),) + fle_b200.capi._build_program(prog)
    mn, md = timeit(lambda: ctx.eval_dev(pc, prog, n, bitmap=bm, count=cnt, prepared=prep), working_set=G * 25 * 8)
    rec(out, what="program", name="q6_optional_cols_no_nulls", variant="1j", rows=n, sum_bw=25, ms_min=mn, ms_med=md,
        alg_bytes=G * 25 * 8 + G * 8, rows_per_s=n / (mn * 1e-3), count=int(cnt.item()))
    # ... and through fle_eval_no_nulls_dev: the same two kernels, the NULL count stays on the device (no host sync)
    o2 = torch.zeros(2, dtype=torch.int64, device="cuda")
    mn, md = timeit(lambda: ctx.eval_no_nulls_dev(pc, prog, n, bitmap=bm, out=o2, prepared=prep), working_set=G * 25 * 8)
    rec(out, what="program", name="q6_optional_cols_no_nulls_async", variant="1j", rows=n, sum_bw=25, ms_min=mn, ms_med=md,
        alg_bytes=G * 25 * 8 + G * 8, rows_per_s=n / (mn * 1e-3), count=int(o2[0].item()), null_rows=int(o2[1].item()))


def sweep_nulls(ctx, out, n):
    """Q6 over three OPTIONAL columns that DO hold NULLs (2 %): per column one fused evaluation in the data-ordinal domain
    and one deposit through the def-level planes, folded into the output bitmap (hdfs-parquet-scanner.cc:326-345);
    the call includes the def-level count kernel, the optimistic launch and the host sync that picks the path."""
    G = (n + 63) // 64
    g = torch.Generator(device="cuda").manual_seed(77)
    bm = torch.empty(G, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    pc, want, by = [], None, G * 8
    for (d, bw), (lo, hi) in zip(((2526, 12), (11, 4), (50, 6)), ((730, 1094), (5, 7), (0, 22))):
        nn = torch.rand(n, generator=g, device="cuda") >= 0.02
        m = int(nn.sum().item())
        data = torch.randint(0, d, (m,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
        leaf = torch.zeros(n, dtype=torch.bool, device="cuda")
        leaf[nn] = (data >= lo) & (data <= hi)
        want = leaf if want is None else (want & leaf)
        pc.append((ctx.pack_dev(data, bw), bw, ctx.pack_dev(nn.to(torch.int32), 1)))
        by += ((m + 63) // 64) * bw * 8 + G * 8
        del nn, data, leaf
    prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
            (AND, None, None), (LT, 2, 23), (AND, None, None)]
    prep = (ctx.make_col_descs(pc),) + fle_b200.capi._build_program(prog)
    wantbm = None
    for mode, label in ((2, "q6_optional_cols_2pct_nulls"), (1, "q6_optional_cols_2pct_nulls_per_column_kernels")):
        ctx.set_option("null_fused", mode)
        ctx.eval_dev(pc, prog, n, bitmap=bm, count=cnt, prepared=prep)  # (the first call learns that the file holds NULLs)
        l0 = ctx.launch_count
        ctx.eval_dev(pc, prog, n, bitmap=bm, count=cnt, prepared=prep)
        torch.cuda.synchronize()
        per_call = ctx.launch_count - l0
        assert int(cnt.item()) == int(want.sum().item()), "Q6 with NULLs: count differs from the torch compare"
        if wantbm is None:
            wantbm = bm.clone()
        assert bool((bm == wantbm).all()), "Q6 with NULLs: the two NULL paths disagree"
        mn, md = timeit(lambda: ctx.eval_dev(pc, prog, n, bitmap=bm, count=cnt, prepared=prep), working_set=by)
        rec(out, what="program", name=label, variant="1j", rows=n, sum_bw=25, ms_min=mn, ms_med=md,
            alg_bytes=by, rows_per_s=n / (mn * 1e-3), count=int(cnt.item()), launches_per_call=per_call)
    ctx.set_option("null_fused", 2)


def sweep_tune(ctx, out, n):
    G = (n + 63) // 64
    bm = torch.empty(G, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    cases = {}
    v, w12 = make_col(ctx, n, 12, seed=1)
    del v
    cases["between12"] = ([(w12, 12)], [(GE, 0, 1000), (LE, 0, 1409), (AND, None, None)], 12)
    q6c = []
    for i, (d, bw) in enumerate(((2526, 12), (11, 4), (50, 6))):
        v, w = make_col(ctx, n, bw, dom=d, seed=300 + i)
        del v
        q6c.append((w, bw))
    cases["q6"] = (q6c, [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
                         (AND, None, None), (LT, 2, 23), (AND, None, None)], 22)
    for bwx in (1, 4, 32):
        v, w = make_col(ctx, n, bwx, seed=400 + bwx)
        del v
        cases[f"lt{bwx}"] = ([(w, bwx)], [(LT, 0, (1 << bwx) // 2)], bwx)
    v, w3 = make_col(ctx, n, 3, dom=7, seed=500)
    del v
    cases["q12"] = ([(w3, 3), q6c[0]], [(IN, 0, [1, 4]), (GE, 1, 730), (AND, None, None), (LT, 1, 1095), (AND, None, None)], 15)
    v, w2 = make_col(ctx, n, 2, dom=4, seed=501)
    del v
    cases["or_variant"] = ([(w3, 3), (w2, 2), q6c[2]], [(IN, 0, [1, 4]), (EQ, 1, 2), (OR, None, None), (LT, 2, 23), (AND, None, None)], 11)
    for cname, (pc, prog, sum_bw) in cases.items():
        prep = (ctx.make_col_descs(pc),) + fle_b200.capi._build_program(prog)
        for variant in (2, "jit"):
            if variant == "jit" and len(pc) == 1:
                continue
            for stages in (0, 2, 3, 4):
                for gpt in (0, 1, 2, 4, 8):
                    for cps in (0, 1, 2, 3):
                        ctx.set_option("eval_variant", 1 if variant == "jit" else variant)
                        ctx.set_option("eval_jit", 2 if variant == "jit" else 0)
                        ctx.set_option("eval_stages", stages)
                        ctx.set_option("eval_gpt", gpt)
                        ctx.set_option("eval_ctas_per_sm", cps)
                        try:
                            mn, md = timeit(lambda: ctx.eval_dev(pc, prog, n, bitmap=bm, count=cnt, prepared=prep), reps=4, warm=2, working_set=G * sum_bw * 8)
                        except fle_b200.FleError as e:
                            continue
                        rec(out, what="tune", case=cname, block=variant, stages=stages, gpt=gpt, ctas=cps, ms_min=mn,
                            ms_med=md, alg_bytes=G * sum_bw * 8 + G * 8)
    ctx.set_option("eval_variant", 1)
    ctx.set_option("eval_jit", 1)
    for k, d in (("eval_block", 256), ("eval_stages", 0), ("eval_gpt", 0), ("eval_ctas_per_sm", 0)):
        ctx.set_option(k, d)


def sweep_aux(ctx, out, n):
    """The kernels either side of the fused eval: bitmap count / select / gather / intersect,
    PLAIN predicates, realign, dictionary build."""
    from fle_b200 import GE
    G = (n + 63) // 64
    g = torch.Generator(device="cuda").manual_seed(9)
    v, w = make_col(ctx, n, 12, dom=2526, seed=900)
    bm, cnt = ctx.eval_dev([(w, 12)], [(GE, 0, 2480)], n)  # ~1.8 % survivors
    torch.cuda.synchronize()
    k = int(cnt.item())
    mn, md = timeit(lambda: ctx.bitmap_count_dev(bm, n, count=cnt))
    rec(out, what="aux", name="bitmap_count", rows=n, ms_min=mn, ms_med=md, alg_bytes=G * 8)
    ids = torch.empty(k + 64, dtype=torch.int32, device="cuda")
    mn, md = timeit(lambda: ctx.bitmap_select_dev(bm, n, cap=k + 64, ids=ids, count=cnt))
    rec(out, what="aux", name=f"bitmap_select ({k} survivors, single pass with look-back)", rows=n, ms_min=mn, ms_med=md,
        alg_bytes=G * 8 + k * 4)
    o = torch.empty(k + 64, dtype=torch.int32, device="cuda")
    mn, md = timeit(lambda: ctx.gather_dev(w, 12, ids, cnt, k + 64, out=o))
    rec(out, what="aux", name="gather codes of survivors (64 B fetched per survivor group)", rows=n, ms_min=mn, ms_med=md,
        alg_bytes=k * (64 + 4 + 4))
    full = torch.full((G,), -1, dtype=torch.int64, device="cuda")
    sub = bm.clone()  # n_sub = popcount(root) = n bits
    root = full.clone()
    mn, md = timeit(lambda: ctx.intersect_bitset_dev(root, n, sub, n))
    rec(out, what="aux", name="intersect_bitset (all-ones root)", rows=n, ms_min=mn, ms_med=md, alg_bytes=3 * G * 8)
    del root, full, sub, ids, o, v, w
    for dt, nb in ((torch.int32, 4), (torch.int64, 8), (torch.float64, 8)):
        m = min(n, 1 << 28)
        x = torch.randint(0, 1 << 20, (m,), generator=g, device="cuda", dtype=torch.int64).to(dt)
        pb = torch.empty((m + 63) // 64, dtype=torch.int64, device="cuda")
        mn, md = timeit(lambda: ctx.plain_eval_dev(x, GE, [1 << 19], bitmap=pb, count=cnt), working_set=m * nb)
        rec(out, what="aux", name=f"plain_eval {str(dt).split('.')[-1]}", rows=m, ms_min=mn, ms_med=md,
            alg_bytes=m * nb + (m + 63) // 64 * 8)
        del x, pb
    blob = torch.randint(0, 255, ((1 << 28) + 64,), generator=g, device="cuda", dtype=torch.int64).to(torch.uint8)
    dst = torch.empty((1 << 28) // 8 + 2, dtype=torch.int64, device="cuda")
    mn, md = timeit(lambda: ctx.realign_dev(blob, 5, 1 << 28, out=dst), working_set=2 << 28)
    rec(out, what="aux", name="realign (funnel-shift copy, 256 MiB)", rows=1 << 28, ms_min=mn, ms_med=md, alg_bytes=2 << 28)
    del blob, dst
    m = 1 << 24
    x = torch.randint(0, 2526, (m,), generator=g, device="cuda", dtype=torch.int64)
    mn, md = timeit(lambda: ctx.dict_encode_dev(x), reps=3, warm=1)
    rec(out, what="aux", name="dict_encode int64 (2526 distinct, 16 Mi values; includes its host sync)", rows=m, ms_min=mn,
        ms_med=md, alg_bytes=m * 8 + m * 12 // 8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="+")
    ap.add_argument("--rows", type=int, default=1 << 28)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    torch.cuda.set_device(0)
    ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
    with open(args.out, "a") as out:
        what = set(args.what)
        if what & {"between", "all"}:
            sweep_between(ctx, out, args.rows)
        if what & {"selectivity", "all"}:
            sweep_selectivity(ctx, out, args.rows)
        if what & {"codec", "all"}:
            sweep_codec(ctx, out, min(args.rows, 1 << 26))
        if what & {"programs", "all"}:
            sweep_programs(ctx, out, args.rows)
        if what & {"nulls", "all"}:
            sweep_nulls(ctx, out, args.rows)
        if what & {"tune"}:
            sweep_tune(ctx, out, args.rows)
        if what & {"aux", "all"}:
            sweep_aux(ctx, out, args.rows)
    ctx.close()


if __name__ == "__main__":
    main()

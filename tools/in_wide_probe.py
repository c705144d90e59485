#!/usr/bin/env python
"""IN lists on bw <= 12 columns: entry-table kernel (in_wide = 1) against
the bit-table kernel (in_wide = 0), 600 M rows, bitmaps compared bit for bit.
python tools/in_wide_probe.py [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import IN  # noqa: E402
from tools.sweep import PEAK, make_col, timeit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_037_902
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
G = (n + 63) // 64
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
gen = torch.Generator().manual_seed(7)
ctx.set_option("in_lookup_min", 1)
for bw, dom in ((16, 40000), (14, 16384), (13, 8192), (12, 2526), (11, 2048), (10, 1000), (8, 256), (6, 50), (3, 7))[:int(os.environ.get("NBW", "9"))]:
    v, w = make_col(ctx, n, bw, dom=dom, seed=bw)
    del v
    by = G * bw * 8 + G * 8
    for L in (9, 256):
        lits = torch.randint(0, dom, (L,), generator=gen).tolist()
        prog = [(IN, 0, lits)]
        res = {}
        for name, wide, nt, stages in (("bit", 0, 0, 0), ("wide", 1, 0, 0)):
            ctx.set_option("in_wide", wide)
            ctx.set_option("eval_stages", stages)
            bm = torch.empty(G, dtype=torch.int64, device="cuda")
            try:
                mn, md = timeit(lambda: ctx.eval_dev([(w, bw)], prog, n, bitmap=bm, count=cnt))
            except Exception as e:  # stage count that does not fit
                print(f"bw{bw} L{L} {name}: {e}", flush=True)
                continue
            res[name] = (mn, bm, ctx.get_option("last_mode"))
        ctx.set_option("eval_stages", 0)
        line = f"bw{bw:2d} dom{dom:5d} L{L:3d}:"
        for name, (mn, bm, mode) in res.items():
            same = bool(torch.equal(bm, res["bit"][1]))
            line += f" {name} {mn*1e3:6.1f} us frac {by/mn/1e6/PEAK:.3f} mode {mode} same {same} |"
            assert same, name
        print(line, flush=True)
    del w
    torch.cuda.empty_cache()

# long IN lists inside larger programs: regrouped (lookup kernel + fused rest + combine) against the whole program on the
# specialised kernel (in_split_min_rows = -1)
from fle_b200 import AND, LT  # noqa: E402

if os.environ.get("SPLIT", "1") == "1":
    ctx.set_option("in_lookup_min", 9)
    v, w12 = make_col(ctx, n, 12, dom=2526, seed=12)
    del v
    v, w6 = make_col(ctx, n, 6, dom=50, seed=6)
    del v
    by = G * 12 * 8 + G * 6 * 8 + G * 8
    for L in (20, 64, 256):
        prog = [(IN, 0, torch.randint(0, 2526, (L,), generator=gen).tolist()), (LT, 1, 23), (AND, None, None)]
        res = {}
        for name, split in (("whole", -1), ("regrouped", 0)):
            ctx.set_option("in_split_min_rows", split)
            bm = torch.empty(G, dtype=torch.int64, device="cuda")
            mn, md = timeit(lambda: ctx.eval_dev([(w12, 12), (w6, 6)], prog, n, bitmap=bm, count=cnt))
            res[name] = (mn, bm)
        ctx.set_option("in_split_min_rows", 1 << 20)
        same = bool(torch.equal(res["whole"][1], res["regrouped"][1]))
        print(f"IN({L}) on bw12 AND bw6 < 23: whole {res['whole'][0]*1e3:7.1f} us frac {by/res['whole'][0]/1e6/PEAK:.3f} | "
              f"regrouped {res['regrouped'][0]*1e3:7.1f} us frac {by/res['regrouped'][0]/1e6/PEAK:.3f} identical {same}", flush=True)
        assert same

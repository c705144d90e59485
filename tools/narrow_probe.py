#!/usr/bin/env python
"""Single-pass predicates on narrow columns: unstaged coalesced kernel (eval_narrow_bw) against the
TMA-staged one, same inputs, bitmaps compared bit for bit.  python tools/narrow_probe.py [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, EQ, GE, LE, LT  # noqa: E402
from tools.sweep import PEAK, make_col, timeit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_037_902
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
G = (n + 63) // 64
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for bw in (1, 2, 4):
    v, w = make_col(ctx, n, bw, seed=bw)
    del v
    D = 1 << bw
    progs = {"between": [(GE, 0, D // 4), (LE, 0, max(D // 4, D // 2 - 1)), (AND, None, None)],
             "lt": [(LT, 0, D - 1)], "eq": [(EQ, 0, D - 1)]}
    for name, prog in progs.items():
        res = {}
        for nb in (0, 4):
            ctx.set_option("eval_narrow_bw", nb)
            bm = torch.empty(G, dtype=torch.int64, device="cuda")
            mn, md = timeit(lambda: ctx.eval_dev([(w, bw)], prog, n, bitmap=bm, count=cnt))
            res[nb] = (mn, md, bm, int(cnt.item()))
        same = bool(torch.equal(res[0][2], res[4][2])) and res[0][3] == res[4][3]
        by = G * bw * 8 + G * 8
        print(f"bw{bw} {name:8s} staged {res[0][0]*1e3:7.1f} us frac {by/res[0][0]/1e6/PEAK:.3f} | "
              f"narrow {res[4][0]*1e3:7.1f} us frac {by/res[4][0]/1e6/PEAK:.3f} | identical {same}", flush=True)
        assert same
    del w
    torch.cuda.empty_cache()

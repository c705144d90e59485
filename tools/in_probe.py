#!/usr/bin/env python
"""IN lists: the decode + lookup kernel against the plane form (in_lookup_min) per list length, to
place the crossover.  Both bitmaps are compared bit for bit.  python tools/in_probe.py [rows]"""
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
for bw, dom in ((12, 2526), (16, 40000)):
    v, w = make_col(ctx, n, bw, dom=dom, seed=bw)
    del v
    by = G * bw * 8 + G * 8
    for L in (2, 4, 8, 9, 12, 16, 64, 256):
        lits = torch.randint(0, dom, (L,), generator=gen).tolist()
        prog = [(IN, 0, lits)]
        out = {}
        for lm in (1, 1 << 30):
            ctx.set_option("in_lookup_min", lm)
            bm = torch.empty(G, dtype=torch.int64, device="cuda")
            mn, md = timeit(lambda: ctx.eval_dev([(w, bw)], prog, n, bitmap=bm, count=cnt))
            out[lm] = (mn, bm)
        same = bool(torch.equal(out[1][1], out[1 << 30][1]))
        print(f"bw{bw} L{L:3d}: lookup {out[1][0]*1e3:7.1f} us frac {by/out[1][0]/1e6/PEAK:.3f} | planes {out[1 << 30][0]*1e3:7.1f} us "
              f"frac {by/out[1 << 30][0]/1e6/PEAK:.3f} identical {same}", flush=True)
        assert same
    ctx.set_option("in_lookup_min", 9)
    del w
    torch.cuda.empty_cache()

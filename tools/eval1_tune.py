#!/usr/bin/env python
"""Single-pass predicate kernels (eval1_kernel): stages x groups-per-thread x CTAs-per-SM per bit width,
against the launcher's automatic choice.  python tools/eval1_tune.py [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, GE, LE  # noqa: E402
from tools.sweep import PEAK, make_col, timeit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_037_902
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
ctx.set_option("eval_narrow_bw", 0)
G = (n + 63) // 64
bm = torch.empty(G, dtype=torch.int64, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for bw in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32):
    v, w = make_col(ctx, n, bw, seed=bw)
    del v
    D = 1 << bw
    prog = [(GE, 0, D // 4), (LE, 0, max(D // 4, D // 2 - 1)), (AND, None, None)]
    prep = (ctx.make_col_descs([(w, bw)]),) + fle_b200.capi._build_program(prog)
    by = G * bw * 8 + G * 8
    res = []
    for st in (0, 2, 3, 4, 6):
        for gpt in (0, 1, 2, 4, 8):
            for ctas in (0, 1, 2, 3, 4, 6):
                if (st == 0) != (gpt == 0) or (st == 0) != (ctas == 0):
                    continue
                ctx.set_option("eval_stages", st); ctx.set_option("eval_gpt", gpt); ctx.set_option("eval_ctas_per_sm", ctas)
                try:
                    mn, md = timeit(lambda: ctx.eval_dev([(w, bw)], prog, n, bitmap=bm, count=cnt, prepared=prep), reps=4, warm=2)
                except fle_b200.FleError:
                    continue
                if ctx.get_option("last_mode") != 2:  # fell through to another kernel (stages do not fit)
                    continue
                res.append((by / mn / 1e6 / PEAK, st, gpt, ctas))
    auto = [r for r in res if r[1] == 0][0]
    res.sort(reverse=True)
    print(f"bw{bw:2d} auto {auto[0]:.3f} | best " + " ".join(f"{f:.3f}(s{st} g{g} c{c})" for f, st, g, c in res[:5]), flush=True)
    del w
    torch.cuda.empty_cache()

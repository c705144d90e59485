#!/usr/bin/env python
"""Runs ONE program a few times so that ncu can capture it:  python tools/profile_one.py q6|between12|unpack12|pack12 [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, EQ, GE, IN, LE, LT, OR  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "between12"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 600_037_902
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
for o in sys.argv[4:]:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
g = torch.Generator(device="cuda").manual_seed(1)


def col(d, bw):
    v = torch.randint(0, d, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    w = ctx.pack_dev(v, bw)
    return v, w


G = (n + 63) // 64
bm = torch.empty(G, dtype=torch.int64, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
if case == "q6":
    cols = [(col(2526, 12)[1], 12), (col(11, 4)[1], 4), (col(50, 6)[1], 6)]
    prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
            (AND, None, None), (LT, 2, 23), (AND, None, None)]
    for _ in range(reps):
        ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt)
elif case == "q6nulls":
    # Q6 over three OPTIONAL columns with 2 % NULLs each (the NULL path: tile counts, prefixes, chain kernel)
    cols = []
    for d, bw in ((2526, 12), (11, 4), (50, 6)):
        nn = torch.rand(n, generator=g, device="cuda") >= 0.02
        data = torch.randint(0, d, (int(nn.sum().item()),), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
        cols.append((ctx.pack_dev(data, bw), bw, ctx.pack_dev(nn.to(torch.int32), 1)))
        del nn, data
    prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
            (AND, None, None), (LT, 2, 23), (AND, None, None)]
    for _ in range(reps):
        ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt)
elif case == "q6agg":
    cols = [(col(2526, 12)[1], 12), (col(11, 4)[1], 4), (col(50, 6)[1], 6)]
    prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
            (AND, None, None), (LT, 2, 23), (AND, None, None)]
    price = torch.randint(90_000, 10_500_000, (n,), generator=g, device="cuda", dtype=torch.int64)
    dd = torch.arange(0, 11, device="cuda", dtype=torch.int64)
    for _ in range(reps):
        out = ctx.eval_sum_product_dev(cols, prog, n, price, 1, dd, bitmap=bm)
    cnt = out[1:2]
elif case == "or_variant":
    cols = [(col(7, 3)[1], 3), (col(4, 2)[1], 2), (col(50, 6)[1], 6)]
    prog = [(IN, 0, [1, 4]), (EQ, 1, 2), (OR, None, None), (LT, 2, 23), (AND, None, None)]
    for _ in range(reps):
        ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt)
elif case == "q12":
    cols = [(col(7, 3)[1], 3), (col(2526, 12)[1], 12)]
    prog = [(IN, 0, [1, 4]), (GE, 1, 730), (AND, None, None), (LT, 1, 1095), (AND, None, None)]
    for _ in range(reps):
        ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt)
elif case.startswith("between"):
    bw = int(case[len("between"):])
    cols = [(col(1 << bw, bw)[1], bw)]
    D = 1 << bw
    prog = [(GE, 0, D // 4), (LE, 0, D // 4 + max(1, D // 10) - 1), (AND, None, None)]
    for _ in range(reps):
        ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt)
elif case.startswith("in"):
    # in<L> (bw 12, 2526 codes) or in<L>x<bw>
    L, _, bwx = case[2:].partition("x")
    L, bwx = int(L), int(bwx or 12)
    dom = 2526 if bwx == 12 else min(1 << bwx, 40000)
    cols = [(col(dom, bwx)[1], bwx)]
    prog = [(IN, 0, torch.randint(0, dom, (L,)).tolist())]
    for _ in range(reps):
        ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt)
elif case.startswith("unpack") or case.startswith("pack"):
    bw = int(case.lstrip("unpack"))
    v, w = col(1 << bw, bw)
    o = torch.empty(n, dtype=torch.int32, device="cuda")
    w2 = torch.empty_like(w)
    for _ in range(reps):
        if case.startswith("unpack"):
            ctx.unpack_dev(w, bw, 0, n, out=o)
        else:
            ctx.pack_dev(v, bw, out=w2)
torch.cuda.synchronize()
print(case, n, "count", int(cnt.item()), "launches", ctx.launch_count)
ctx.close()

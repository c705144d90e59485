#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Sizes are chosen to hit full TMA tiles, ragged tiles and ragged rows while staying fast under
the sanitizer.  Results are checked against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, EQ, GE, IN, LE, LT, OR  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
orc = Oracle()
rng = np.random.default_rng(5)
n = 64 * 2048 * 2 + 64 * 3 + 7
for bw in (1, 5, 12, 16, 17, 32):
    vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
    w = orc.pack(vals, bw)
    for cv in (1, 0, 2, 3, 4):
        ctx.set_option("codec_variant", cv)
        assert (ctx.pack(vals, bw) == w).all()
        assert (ctx.unpack(w, bw, 0, n) == vals).all()
    ctx.set_option("codec_variant", 1)
    lo, hi = (1 << bw) // 4, (3 << bw) // 4
    for ev in (1, 2, 0):
        ctx.set_option("eval_variant", ev)
        for prog in ([(GE, 0, lo), (LE, 0, hi), (AND, None, None)], [(EQ, 0, int(vals[3]))],
                     [(IN, 0, rng.choice(vals, size=40).astype(np.uint64))]):
            bm, c = ctx.eval([(w, bw)], prog, n)
            want, wc = orc.eval_tree([(w, bw)], prog, n)
            assert (bm == want).all() and c == wc
    ctx.set_option("eval_variant", 1)
cols = [(orc.pack(rng.integers(0, d, size=n).astype(np.uint32), bw), bw) for d, bw in ((2526, 12), (11, 4), (50, 6))]
q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None), (AND, None, None),
      (LT, 2, 23), (AND, None, None)]
orv = [(IN, 1, [2, 9]), (EQ, 2, 17), (OR, None, None), (LT, 0, 1200), (AND, None, None)]
for ev in (1, 0):
    ctx.set_option("eval_variant", ev)
    for prog in (q6, orv):
        bm, c = ctx.eval(cols, prog, n)
        want, wc = orc.eval_tree(cols, prog, n)
        assert (bm == want).all() and c == wc
ctx.set_option("eval_variant", 1)
# nullable, materialisation, aggregate, PLAIN, realign, dictionary build
nn = rng.random(n) >= 0.1
data = rng.integers(0, 2526, size=int(nn.sum())).astype(np.uint32)
bm, c = ctx.eval([(orc.pack(data, 12), 12, orc.pack(nn.astype(np.uint32), 1))], [(LT, 0, 1000)], n)
dw = torch.from_numpy(cols[1][0].view(np.int64)).cuda()
sw = torch.from_numpy(cols[0][0].view(np.int64)).cuda()
bmd, cd = ctx.eval_dev([(sw, 12)], [(LT, 0, 300)], n)
ids, c2 = ctx.bitmap_select_dev(bmd, n)
codes = ctx.gather_dev(dw, 4, ids, c2, n)
price = torch.randint(1, 1000, (n,), device="cuda", dtype=torch.int64)
out = ctx.sum_product_dev(bmd, n, price, dw, 4, torch.arange(0, 11, device="cuda", dtype=torch.int64))
pv = torch.randint(-50, 50, (n,), device="cuda", dtype=torch.int32)
ctx.plain_eval_dev(pv, LE, [3])
ctx.plain_eval_dev(pv.to(torch.float64), IN, [1.0, 2.0])
blob = torch.randint(0, 255, (5000,), device="cuda", dtype=torch.uint8)
ctx.realign_dev(blob, 3, 4001)
d, bw, enc = ctx.dict_encode_dev(torch.randint(0, 3000, (n,), device="cuda", dtype=torch.int64))
torch.cuda.synchronize()
print("sanitize_small ok", int(cd.item()), int(c2.item()), int(out[0].item()), bw, d.numel())
ctx.close()

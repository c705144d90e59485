#!/usr/bin/env python
"""Unpack kernels side by side (codec_variant 1 = per-width choice, 2 = per-row bulk stores, 3 = warp-private
copy-out, 4 = staged copy-out), outputs compared bit for bit.  python tools/codec_probe.py [values]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from tools.sweep import PEAK, make_col, timeit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
for bw in (1, 2, 4, 6, 8, 10, 12, 14, 16, 17, 20, 24, 28, 32):
    v, w = make_col(ctx, n, bw, seed=bw)
    by = (n + 63) // 64 * bw * 8 + n * 4
    res = {}
    for var in (1, 2, 3, 4):
        ctx.set_option("codec_variant", var)
        o = torch.empty(n, dtype=torch.int32, device="cuda")
        mn, md = timeit(lambda: ctx.unpack_dev(w, bw, 0, n, out=o))
        res[var] = (mn, o)
    ok = all(bool(torch.equal(res[k][1], v)) for k in res)
    print(f"bw{bw:2d} unpack " + " | ".join(f"v{k} {res[k][0]*1e3:6.1f} us frac {by/res[k][0]/1e6/PEAK:.3f}" for k in res) + f" | ok {ok}", flush=True)
    assert ok
    del v, w
# ragged sizes through the bulk-store and warp-private kernels
for var in (2, 3):
    ctx.set_option("codec_variant", var)
    for m in (1, 63, 64, 65, 8191, 8192, 8193, 128 * 64 * 3 + 77, 128 * 64 * 1200 + 5):
        for bw in (3, 12, 16, 25, 32):
            v, w = make_col(ctx, m, bw, seed=m + bw)
            o = ctx.unpack_dev(w, bw, 0, m)
            assert torch.equal(o, v), (var, m, bw)
print("ragged ok")

#!/usr/bin/env python
"""cfg0 kernels under the bench's conditions (L2 flushed before every timed launch, median): unpack variants
(2 = per-row bulk stores, 3 = warp-private copy-out, 4 = staged copy-out) x CTAs per SM, pack x CTAs per SM.
    python tools/codec_cold_probe.py [values]"""
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
for bw in (1, 2, 4, 8, 12, 16, 24, 32):
    v, w = make_col(ctx, n, bw, seed=bw)
    by = (n + 63) // 64 * bw * 8 + n * 4
    o = torch.empty(n, dtype=torch.int32, device="cuda")
    cells = []
    ctx.set_option("codec_variant", 1)
    ctx.set_option("codec_ctas_per_sm", 0)
    _, md = timeit(lambda: ctx.unpack_dev(w, bw, 0, n, out=o), working_set=0)
    cells.append(f"auto {by / md / 1e6 / PEAK:.3f}")
    for var in (2, 3, 4):
        for ctas in (2, 3, 4):
            ctx.set_option("codec_variant", var)
            ctx.set_option("codec_ctas_per_sm", ctas)
            _, md = timeit(lambda: ctx.unpack_dev(w, bw, 0, n, out=o), working_set=0)
            cells.append(f"v{var}c{ctas} {by / md / 1e6 / PEAK:.3f}")
    print(f"bw{bw:2d} unpack " + " ".join(cells), flush=True)
    assert torch.equal(o, v)
    w2 = torch.empty_like(w)
    cells = []
    ctx.set_option("codec_variant", 1)
    for ctas in (0, 1, 2, 3):
        ctx.set_option("codec_ctas_per_sm", ctas)
        _, md = timeit(lambda: ctx.pack_dev(v, bw, out=w2), working_set=0)
        cells.append(f"c{ctas} {by / md / 1e6 / PEAK:.3f} ({md * 1e3:.1f} us)")
    ctx.set_option("codec_ctas_per_sm", 0)
    print(f"bw{bw:2d} pack   " + " ".join(cells), flush=True)
    assert torch.equal(w2, w)
    del v, w, o, w2

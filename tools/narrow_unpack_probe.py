#!/usr/bin/env python
"""Narrow decode targets (fle_unpack_u8/u16_dev): warp-private copy-out (codec_variant 1) against the first form
(codec_variant 4), cold L2, median of 9; frac = (bw * 8 + 64 * sizeof(out)) bytes per group / time / measured HBM peak.
    python tools/narrow_unpack_probe.py [ctas_per_sm ...]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
clean = torch.zeros(256 << 17, dtype=torch.int64, device="cuda")
sink = torch.zeros(1, dtype=torch.int64, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(7)


def cold(fn, reps=9):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        clean.sum()  # leave clean lines in the L2, not the overwrite's dirty ones
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


ctas_list = [int(x) for x in sys.argv[1:]] or [0]
print(f"{'values':>9} {'bw':>3} {'out':>4} | " + " | ".join(f"v{v} ctas{c}: ms  frac" for v in (1, 4) for c in ctas_list))
for n in (1 << 24, 1 << 26):
    G = n // 64
    for bw in (1, 4, 8, 12, 16):
        v = torch.randint(0, 1 << bw, (n,), generator=gen, device="cuda", dtype=torch.int64).to(torch.int32)
        w = ctx.pack_dev(v, bw)
        for es in ((1, 2) if bw <= 8 else (2,)):
            o = torch.empty(n, dtype=torch.uint8 if es == 1 else torch.int16, device="cuda")
            by = G * bw * 8 + n * es
            cells = []
            for variant in (1, 4):
                for c in ctas_list:
                    ctx.set_option("codec_variant", variant)
                    ctx.set_option("codec_ctas_per_sm", c)
                    ms = cold(lambda: ctx.unpack_narrow_dev(w, bw, 0, n, out=o))
                    assert torch.equal(o.to(torch.int32) & ((1 << (8 * es)) - 1), v)
                    cells.append(f"{ms:8.4f} {by / ms / 1e6 / PEAK:5.3f}")
            ctx.set_option("codec_variant", 1)
            ctx.set_option("codec_ctas_per_sm", 0)
            print(f"{n:>9} {bw:>3} {'u' + str(8 * es):>4} | " + " | ".join(cells), flush=True)
        del v, w
ctx.close()

#!/usr/bin/env python
"""Dictionary build + encode (fle_dict_encode_dev) at growing sizes: python tools/dict_probe.py
frac = (values read once + packed planes written once) / time / measured HBM peak; the call includes its host sync."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402

peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(3)
for dt, es in ((torch.int64, 8), (torch.int32, 4)):
    for logn in (24, 26, 28):
        for nd in (2526, 30000):
            n = 1 << logn
            x = torch.randint(0, nd, (n,), generator=g, device="cuda", dtype=torch.int64).to(dt) * 7 - 1000
            for _ in range(2):
                d, bw, enc = ctx.dict_encode_dev(x)
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                d, bw, enc = ctx.dict_encode_dev(x)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            by = n * es + enc.numel() * 8
            print(f"{str(dt).split('.')[-1]:6s} n=2^{logn} distinct={d.numel():6d} bw={bw:2d}  {ms:8.3f} ms  {by / ms / 1e6:8.1f} GB/s  frac {by / ms / 1e6 / peak:.3f}", flush=True)
            del x, enc

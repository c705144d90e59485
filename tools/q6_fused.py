#!/usr/bin/env python
"""Times TPC-H Q6 on device-resident columns: predicate + aggregate in one specialised kernel
(fle_eval_sum_product_dev) against predicate kernel + aggregate kernel, over tile configurations.
  python tools/q6_fused.py [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, GE, LE, LT  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_037_902
torch.cuda.set_device(0)
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1)
cols = []
for d, bw in ((2526, 12), (11, 4), (50, 6)):
    v = torch.randint(0, d, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    cols.append((ctx.pack_dev(v, bw), bw))
    del v
price = torch.randint(90_000, 10_500_000, (n,), generator=g, device="cuda", dtype=torch.int64)
dd = torch.arange(0, 11, device="cuda", dtype=torch.int64)
prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None), (AND, None, None),
        (LT, 2, 23), (AND, None, None)]
G = (n + 63) // 64
bm = torch.empty(G, dtype=torch.int64, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
out = torch.zeros(2, dtype=torch.int64, device="cuda")
prep = (ctx.make_col_descs(cols),) + fle_b200.capi._build_program(prog)


def timeit(fn, reps=6, inner=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / inner)
    return best


def two():
    ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt, prepared=prep)
    ctx.sum_product_dev(bm, n, price, cols[1][0], 4, dd, out=out)


t2 = timeit(two)
ref = out.clone()
print(f"two kernels           {t2:.4f} ms   scan alone {timeit(lambda: ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt, prepared=prep)):.4f} ms")
for gpt in (0, 1):
    for stages in (0, 3):
        for ctas in (0,):
            ctx.set_option("eval_gpt", gpt)
            ctx.set_option("eval_stages", stages)
            ctx.set_option("eval_ctas_per_sm", ctas)
            for with_bm in (True, False):
                try:
                    t = timeit(lambda: ctx.eval_sum_product_dev(cols, prog, n, price, 1, dd, bitmap=bm if with_bm else None,
                                                                out=out, prepared=prep))
                except fle_b200.FleError as e:
                    print("gpt", gpt, "stages", stages, "ctas", ctas, "failed", e)
                    continue
                ok = bool(torch.equal(out, ref))
                print(f"fused gpt {gpt} stages {stages} ctas {ctas} bitmap {int(with_bm)}: {t:.4f} ms  ok={ok}")
ctx.close()

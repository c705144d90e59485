#!/usr/bin/env python
"""IN-list sweep on one B200: bw 3..16 x L in {8, 64, 256} at 600,037,902 rows (inputs > L2), median of 7
launches, fraction of the measured HBM peak; counts checked against torch.isin.
    python tools/in_sweep.py [--rows N] [--out gpurun_out/in_sweep.txt] [--opt name=value ...]"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, IN, LT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=600_037_902)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "in_sweep.txt"))
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--bws", default="3,4,5,6,7,8,9,10,11,12,13,14,15,16")
    args = ap.parse_args()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    torch.cuda.set_device(0)
    ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    n = args.rows
    G = (n + 63) // 64
    bm = torch.empty(G, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0xF1E00300)
    cg = torch.Generator().manual_seed(5)
    qty = torch.randint(0, 50, (n,), generator=gen, device="cuda", dtype=torch.int32)
    qcol = (ctx.pack_dev(qty, 6), 6)
    lines = [f"IN-list sweep, {n} rows, median of 7 launches (CUDA events, a 384 MB overwrite queued before each), frac = (sum bw + 1) * 8 B per 64 rows / time / {peak} GB/s; "
             f"options {args.opt}", f"{'bw':>3} {'L':>4} {'ms':>8} {'GB/s':>8} {'frac':>6} {'mode':>5} {'launches':>8} | IN(L) AND qty<23: {'ms':>8} {'frac':>6} {'launches':>8}"]

    flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
    clean = torch.zeros(256 << 17, dtype=torch.int64, device="cuda")
    sink = torch.zeros(1, dtype=torch.int64, device="cuda")

    def timed(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1)  # cold L2, and the GPU is busy while the host enqueues the timed launch
            clean.sum()  # ... and no dirty lines left for the timed kernel to write back
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    for bw in [int(x) for x in args.bws.split(",")]:
        d = min(1 << bw, 40000)
        v = torch.randint(0, d, (n,), generator=gen, device="cuda", dtype=torch.int32)
        col = (ctx.pack_dev(v, bw), bw)
        for L in (8, 64, 256):
            L_eff = min(L, max(1, d // 2))
            lits = torch.randperm(d, generator=cg)[:L_eff].tolist()
            prog = [(IN, 0, lits)]
            prep = (ctx.make_col_descs([col]),) + fle_b200.capi._build_program(prog)
            fn = lambda: ctx.eval_dev([col], prog, n, bitmap=bm, count=cnt, prepared=prep)  # noqa: E731
            l0 = ctx.launch_count
            fn()
            torch.cuda.synchronize()
            per = ctx.launch_count - l0
            mode = ctx.get_option("last_mode")
            member = torch.isin(v, torch.tensor(lits, device="cuda", dtype=torch.int32))
            assert int(cnt.item()) == int(member.sum().item()), (bw, L)
            ms = timed(fn)
            by = G * (bw + 1) * 8
            prog2 = [(IN, 0, lits), (LT, 1, 23), (AND, None, None)]
            prep2 = (ctx.make_col_descs([col, qcol]),) + fle_b200.capi._build_program(prog2)
            fn2 = lambda: ctx.eval_dev([col, qcol], prog2, n, bitmap=bm, count=cnt, prepared=prep2)  # noqa: E731
            l0 = ctx.launch_count
            fn2()
            torch.cuda.synchronize()
            per2 = ctx.launch_count - l0
            assert int(cnt.item()) == int((member & (qty < 23)).sum().item()), (bw, L, "and")
            ms2 = timed(fn2)
            by2 = G * (bw + 6 + 1) * 8
            lines.append(f"{bw:>3} {L_eff:>4} {ms:8.4f} {by / ms / 1e6:8.0f} {by / ms / 1e6 / peak:6.3f} {mode:>5} {per:>8} | "
                         f"{'':>17} {ms2:8.4f} {by2 / ms2 / 1e6 / peak:6.3f} {per2:>8}")
            print(lines[-1], flush=True)
            del member
        del v, col
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()

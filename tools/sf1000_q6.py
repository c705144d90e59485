#!/usr/bin/env python
"""BASELINE.json configs[4]: TPC-H Q6 over SF1000-shaped lineitem (5,999,989,709 rows in row groups of
4,194,304 rows), the row groups dealt out contiguously to the GPUs of one box (strong scaling):

    python tools/sf1000_q6.py                                   # 1 GPU (65 GB of columns fit one B200)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/sf1000_q6.py

Every rank builds its shard on the device (synthetic, SURVEY 8d cfg2 domains), runs the fused
predicate + SUM(extendedprice * discount) kernel over it and the job ends with ONE all-reduce of
{revenue, count} (NCCL).  Times are CUDA events on the launching stream, max over ranks; the
reduced count and the exact int64 revenue are checked against a torch compare of the same data.
Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fle_b200  # noqa: E402
from fle_b200 import AND, GE, LE, LT  # noqa: E402

from fle_b200 import partition  # noqa: E402

SF1000_ROWS = 5_999_989_709
COLS = ((2526, 12, 730, 1094), (11, 4, 5, 7), (50, 6, 0, 22))  # (domain, bit width, lo, hi): shipdate, discount, quantity
PROG = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
        (AND, None, None), (LT, 2, 23), (AND, None, None)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-rows", type=int, default=SF1000_ROWS)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunk-row-groups", type=int, default=1,
                    help="row groups generated per seed; 1 => the data (and so count and revenue) is the same for any GPU count")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    ctx = fle_b200.FleContext(local_rank, stream.cuda_stream)

    first, end = partition.rank_rows(args.total_rows, world, rank)
    n = end - first
    G = (n + 63) // 64
    RG = partition.ROW_GROUP_ROWS
    planes = [torch.empty(G * bw, dtype=torch.int64, device=dev) for _, bw, _, _ in COLS]
    price = torch.empty(n, dtype=torch.int64, device=dev)
    disc_dict = torch.arange(0, 11, device=dev, dtype=torch.int64)
    want = torch.zeros(2, dtype=torch.int64, device=dev)  # revenue, count by torch compare
    # shard built row group by row group (seeded per row group => the data does not depend on the GPU count)
    chunk = args.chunk_row_groups * RG
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        gen = torch.Generator(device=dev).manual_seed(0xF1E00400 + (first + r0) // RG)
        p = torch.randint(90_000, 10_500_000, (m,), generator=gen, device=dev, dtype=torch.int64)
        price[r0:r0 + m] = p
        mask = None
        dv = None
        for c, (d, bw, lo, hi) in enumerate(COLS):
            v = torch.randint(0, d, (m,), generator=gen, device=dev, dtype=torch.int32)
            g0 = r0 // 64
            ctx.pack_dev(v, bw, out=planes[c][g0 * bw:(g0 + (m + 63) // 64) * bw])
            mk = (v >= lo) & (v <= hi)
            mask = mk if mask is None else (mask & mk)
            if c == 1:
                dv = v
        want[0] += (p[mask] * disc_dict[dv[mask].long()]).sum()
        want[1] += mask.sum()
        del p, v, mk, mask, dv
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    cols = [(planes[c], COLS[c][1]) for c in range(3)]
    prep = (ctx.make_col_descs(cols),) + fle_b200.capi._build_program(PROG)
    bitmap = torch.empty(G, dtype=torch.int64, device=dev)
    agg = torch.zeros(2, dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)

    def q6():
        ctx.eval_sum_product_dev(cols, PROG, n, price, 1, disc_dict, bitmap=bitmap, out=agg, prepared=prep)
        if world > 1:
            dist.all_reduce(agg)

    def scan():
        ctx.eval_dev(cols, PROG, n, bitmap=bitmap, count=count, prepared=prep)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            fn()
        b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    q6_ms = timed(q6)
    if world > 1:
        dist.all_reduce(want)
    ok = bool(torch.equal(agg, want))
    scan_ms = timed(scan)
    cnt_total = count.clone()
    if world > 1:
        dist.all_reduce(cnt_total)
    ok = ok and int(cnt_total.item()) == int(want[1].item())

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    Gt = (args.total_rows + 63) // 64
    scan_bytes = Gt * (12 + 4 + 6) * 8 + Gt * 8
    if rank == 0:
        print(json.dumps({
            "workload": "cfg4 TPC-H Q6 over SF1000-shaped lineitem, row groups of 4194304 rows partitioned contiguously "
                        "across the GPUs (strong scaling), fused predicate + exact int64 SUM(price*discount) + one NCCL all-reduce",
            "total_rows": args.total_rows, "n_gpus": world, "rows_rank0": n, "steps": args.steps,
            "q6_ms": q6_ms, "q6_rows_per_s": args.total_rows / (q6_ms * 1e-3),
            "scan_ms": scan_ms, "scan_rows_per_s": args.total_rows / (scan_ms * 1e-3),
            "scan_GBps": scan_bytes / (scan_ms * 1e-3) / 1e9, "scan_frac_of_n_x_peak": scan_bytes / (scan_ms * 1e-3) / 1e9 / (peak * world),
            "peak_GBps_per_gpu": peak, "survivors": int(want[1].item()), "revenue_exact_int64_1e-4_dollars": int(want[0].item()),
            "matches_torch_compare": ok}))
    assert ok, "Q6 count / revenue differ from the torch compare"
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (see DESIGN.md "Measurement").

A *step* is one fused predicate pass over one rank's synthetic lineitem shard:
  workload cfg1 (default, BASELINE.json configs[1]): single-column BETWEEN on an FLE-encoded
  int32 column, bw 12, N = 600,037,902 rows per GPU (SF100 lineitem), selectivity cycling
  1/10/20/30/50/80/100 % over the steps; for N > 1 every rank scans its own shard (row groups
  partitioned, weak scaling) and the per-rank match counts are summed with one NCCL all-reduce.
  The same run also times the TPC-H Q6 compound scan (3 columns, 5 leaves) and reports it
  under "q6".

value      rows/s, whole job, inputs resident in HBM, CUDA events, max over ranks
e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of the encoded column
           and D2H of bitmap+count inside the timed region
roofline   dominant kernel (fused eval): algorithmic bytes / mean launch duration vs the
           measured HBM copy peak (MEASURED_PEAKS.json)
cpu_baseline  the reference's own fle-encoding.h (oracle/_ref) on this box's host cores,
           bounded sample of the same column
--impl reference: only that CPU arm, as its own JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; stdout carries exactly one JSON line
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
# ... and whatever NCCL does log goes to stderr, not into the JSON stream
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

N_ROWS = 600_037_902          # SF100 lineitem
BW = 12
SELECTIVITIES = (1, 10, 20, 30, 50, 80, 100)
METRIC = "rows_per_sec_filtered"
UNIT = "rows/s"
CPU_SAMPLE_ROWS = 1 << 27     # 134,217,728 rows of the same column for the CPU arm


def between_bounds(sel_pct: int, domain: int):
    width = max(1, round(sel_pct * domain / 100))
    lo = (domain - width) // 2
    return lo, lo + width - 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.index = index
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, k), int):
                names[getattr(nv, k)] = k[len("nvmlClocksThrottleReason"):]
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (r & bit) and nm not in ("None", "All", "GpuIdle"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------------------------
def cpu_reference_run(words_sample: np.ndarray, n_sample: int, steps: int, warmup: int, budget_s: float = 20.0):
    """Times the reference's FleDecoder::Ge + Le + `&=` (BETWEEN as And(Ge, Le)) on host cores,
    one decoder per thread over disjoint row-group ranges.  Returns dict for the JSON line."""
    from oracle.oracle import GE, LE, Ref, ref_available
    cores = os.cpu_count() or 1
    if ref_available():
        ref = Ref()
        kind = "reference"
    else:
        raise RuntimeError("oracle/_ref/libfle_ref.so missing; build it where /root/reference is mounted")
    padded = np.concatenate([words_sample, np.zeros(16, dtype=np.uint64)])
    times, count = [], 0
    t_start = time.time()
    it = 0
    while it < warmup + steps and (time.time() - t_start < budget_s or it < warmup + 1):
        sel = SELECTIVITIES[it % len(SELECTIVITIES)]
        lo, hi = between_bounds(sel, 1 << BW)
        s, count = ref.time_and_leaves([(padded, BW)], [(GE, 0, lo), (LE, 0, hi)], n_sample, cores, mode=0, reps=1)
        if it >= warmup:
            times.append(s)
        it += 1
    mean_s = sum(times) / len(times)
    # scanner-faithful variant: 1024-row windows + Skip (hdfs-parquet-scanner.cc:1837-1865)
    lo, hi = between_bounds(10, 1 << BW)
    s_win, _ = ref.time_and_leaves([(padded, BW)], [(GE, 0, lo), (LE, 0, hi)], n_sample, cores, mode=1, reps=1)
    s_1t, _ = ref.time_and_leaves([(padded, BW)], [(GE, 0, lo), (LE, 0, hi)], min(n_sample, 1 << 25), 1, mode=0, reps=1)
    return {
        "value": n_sample / mean_s, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"{n_sample} rows of the same bw{BW} column, BETWEEN = Ge+Le+&= whole-range calls, "
                  f"{len(times)} passes, selectivity cycling {'/'.join(map(str, SELECTIVITIES))}%",
        "ms_per_pass": mean_s * 1e3,
        "windowed_1024_rows_per_s": n_sample / s_win,
        "single_thread_rows_per_s": min(n_sample, 1 << 25) / s_1t,
        "steps_timed": len(times),
    }


def make_cpu_sample(seed: int):
    """The first CPU_SAMPLE_ROWS rows of the synthetic column, generated and packed on the host
    (numpy + the C oracle's pack) so that the CPU arm needs no GPU."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(seed)
    vals = rng.integers(0, 1 << BW, size=CPU_SAMPLE_ROWS, dtype=np.uint32)
    return Oracle().pack(vals, BW)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        words = make_cpu_sample(0xF1E00100)
        cb = cpu_reference_run(words, CPU_SAMPLE_ROWS, args.steps, args.warmup, budget_s=120.0)
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:200]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": cb["steps_timed"], "warmup": args.warmup, "ms_per_step": cb["ms_per_pass"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"cfg1 single-column BETWEEN, bw{BW}, uniform codes; each step = one pass over a "
                               f"{CPU_SAMPLE_ROWS}-row sample of the {N_ROWS}-row column on host cores",
                   "rows_per_step": CPU_SAMPLE_ROWS, "selectivities_pct": list(SELECTIVITIES)},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "windowed_1024_rows_per_s": cb["windowed_1024_rows_per_s"],
        "single_thread_rows_per_s": cb["single_thread_rows_per_s"],
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=70)
    ap.add_argument("--warmup", type=int, default=7)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows per GPU (default: SF100 lineitem)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value")
    ap.add_argument("--no-reduce", action="store_true", help="diagnostic: skip the per-step count all-reduce (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import fle_b200
    from fle_b200 import AND, GE, LE, LT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    ctx = fle_b200.FleContext(local_rank, stream.cuda_stream)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))

    n = args.rows
    G = (n + 63) // 64
    seed = 0xF1E00100 + rank
    gen = torch.Generator(device=dev).manual_seed(seed)
    # cfg1 column: uniform codes in [0, 2^12)
    vals = torch.randint(0, 1 << BW, (n,), generator=gen, device=dev, dtype=torch.int32)
    col = ctx.pack_dev(vals, BW)
    # Q6 columns: shipdate D=2526 bw12, discount D=11 bw4, quantity D=50 bw6
    q6_cols = []
    q6_mask = None
    price = torch.randint(90_000, 10_500_000, (n,), generator=gen, device=dev, dtype=torch.int64)  # l_extendedprice, cents (PLAIN)
    disc_dict = torch.arange(0, 11, device=dev, dtype=torch.int64)                                  # l_discount dictionary, cents
    for (d, bw), (lo_, hi_) in zip(((2526, 12), (11, 4), (50, 6)), ((730, 1094), (5, 7), (0, 22))):
        v = torch.randint(0, d, (n,), generator=gen, device=dev, dtype=torch.int32)
        q6_cols.append((ctx.pack_dev(v, bw), bw))
        m_ = (v >= lo_) & (v <= hi_)
        q6_mask = m_ if q6_mask is None else (q6_mask & m_)
        if d == 11:
            disc_vals = v
        else:
            del v
    q6_want_count = int(q6_mask.sum().item())
    q6_want_revenue = int((price[q6_mask] * disc_dict[disc_vals[q6_mask].long()]).sum().item())
    del q6_mask, disc_vals, m_
    q6_prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
               (AND, None, None), (LT, 2, 23), (AND, None, None)]
    torch.cuda.synchronize()
    want_counts = {}
    for sel in SELECTIVITIES:
        lo, hi = between_bounds(sel, 1 << BW)
        want_counts[sel] = int(((vals >= lo) & (vals <= hi)).sum().item())
    del vals
    torch.cuda.empty_cache()

    bitmap = torch.empty(G, dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    progs = {}
    for sel in SELECTIVITIES:
        lo, hi = between_bounds(sel, 1 << BW)
        progs[sel] = [(GE, 0, lo), (LE, 0, hi), (AND, None, None)]
    cols = [(col, BW)]
    prepared = {sel: (ctx.make_col_descs(cols),) + fle_b200.capi._build_program(progs[sel]) for sel in SELECTIVITIES}

    count_ring = torch.zeros(max(args.steps, 16), dtype=torch.int64, device=dev)
    pending = []

    def step(i, timed=False):
        sel = SELECTIVITIES[i % len(SELECTIVITIES)]
        if world > 1 and timed and not args.no_reduce:
            # every step's match count is summed over the ranks (8 bytes over NVLink); the
            # reduce of query i is in flight while the kernel of query i+1 runs
            c = count_ring[i:i + 1]
            ctx.eval_dev(cols, progs[sel], n, bitmap=bitmap, count=c, prepared=prepared[sel])
            pending.append(dist.all_reduce(c, async_op=True))
        else:
            ctx.eval_dev(cols, progs[sel], n, bitmap=bitmap, count=count, prepared=prepared[sel])
            if world > 1 and not (timed and args.no_reduce):
                total.copy_(count)
                dist.all_reduce(total)
        return sel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (+ correctness of every selectivity once)
    for i in range(max(args.warmup, len(SELECTIVITIES))):
        sel = step(i)
        torch.cuda.synchronize()
        assert int(count.item()) == want_counts[sel], (sel, int(count.item()), want_counts[sel])

    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    sampler.start()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        step(i, timed=True)
        ev[i][1].record(stream)
    for w in pending:
        w.wait()  # all reduced counts are complete before the clock stops
    t1.record(stream)
    barrier()
    if world > 1 and not args.no_reduce:
        want_total = torch.tensor([want_counts[SELECTIVITIES[i % len(SELECTIVITIES)]] for i in range(args.steps)],
                                  dtype=torch.int64, device=dev)
        dist.all_reduce(want_total)
        assert torch.equal(count_ring[:args.steps], want_total), "reduced match counts differ from the reference compare"
    sampler.stop()
    launches = ctx.launch_count - launches0
    elapsed_ms = t0.elapsed_time(t1)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # TPC-H Q6 end to end on the device: fused compound scan -> bitmap -> SUM(price * discount)
    # over the survivors (exact int64) -> one all-reduce of {count, revenue}
    q6_prep = (ctx.make_col_descs(q6_cols),) + fle_b200.capi._build_program(q6_prog)
    agg = torch.zeros(2, dtype=torch.int64, device=dev)

    def q6_step():
        # predicate + aggregate in ONE specialised kernel (bitmap written as well); where no such kernel
        # can be compiled the library runs the predicate kernel and the aggregate kernel back to back
        ctx.eval_sum_product_dev(q6_cols, q6_prog, n, price, 1, disc_dict, bitmap=bitmap, out=agg, prepared=q6_prep)
        if world > 1:
            dist.all_reduce(agg)

    def q6_two_kernel_step():
        ctx.eval_dev(q6_cols, q6_prog, n, bitmap=bitmap, count=count, prepared=q6_prep)
        ctx.sum_product_dev(bitmap, n, price, q6_cols[1][0], 4, disc_dict, out=agg)

    for _ in range(3):
        q6_step()
    torch.cuda.synchronize()
    if world == 1:
        assert int(agg[0].item()) == q6_want_revenue and int(agg[1].item()) == q6_want_count
    q6_two_kernel_step()  # (also the warm-up of the two-kernel timing below)
    torch.cuda.synchronize()
    if world == 1:
        assert int(agg[0].item()) == q6_want_revenue and int(agg[1].item()) == q6_want_count == int(count.item())
    barrier()
    q6_steps = max(10, args.steps // 4)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(q6_steps):
        q6_step()
    b.record(stream)
    barrier()
    q6_ms = a.elapsed_time(b) / q6_steps
    a.record(stream)
    for _ in range(q6_steps):
        q6_two_kernel_step()
    b.record(stream)
    barrier()
    q6_two_ms = a.elapsed_time(b) / q6_steps
    # scan kernel alone (no aggregate, no reduce)
    a.record(stream)
    for _ in range(q6_steps):
        ctx.eval_dev(q6_cols, q6_prog, n, bitmap=bitmap, count=count, prepared=q6_prep)
    b.record(stream)
    barrier()
    q6_scan_ms = a.elapsed_time(b) / q6_steps
    if world > 1:
        t = torch.tensor([q6_ms, q6_scan_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q6_ms, q6_scan_ms = float(t[0].item()), float(t[1].item())
    q6_count = int(count.item())
    q6_revenue_cents2 = int(agg[0].item())
    q6_bytes = G * (12 + 4 + 6) * 8 + G * 8

    peak, peak_src = measured_peak()
    alg_bytes = G * BW * 8 + G * 8
    mean_kern_ms = sum(kern_ms) / len(kern_ms)
    achieved = alg_bytes / (mean_kern_ms * 1e-3) / 1e9

    # e2e: host (pinned) buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host_words = torch.empty(col.numel(), dtype=torch.int64, pin_memory=True)
        host_words.copy_(col)
        host_bm = torch.empty(G, dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize()
        hw = host_words.numpy().view(np.uint64)
        hb = host_bm.numpy().view(np.uint64)
        import ctypes as C
        descs = (fle_b200.capi.ColDesc * 1)()
        descs[0].planes = hw.ctypes.data
        descs[0].bytes = hw.nbytes
        descs[0].bit_width = BW
        lib = ctx.lib
        cnt = C.c_uint64(0)

        def e2e_step(i):
            sel = SELECTIVITIES[i % len(SELECTIVITIES)]
            nodes = prepared[sel][1]
            rc = lib.fle_eval(ctx.h, descs, 1, nodes, len(nodes), 0, n, hb.ctypes.data, C.byref(cnt))
            assert rc == 0, lib.fle_last_error(ctx.h)
            assert cnt.value == want_counts[sel]

        e2e_step(1)
        barrier()
        w0 = time.perf_counter()
        for i in range(args.e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - w0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": n * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(hw.nbytes),
               "d2h_bytes_per_step": int(hb.nbytes + 8), "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "timing": "host wall clock around fle_eval() (call returns with the result in host memory)"}
        del host_words, host_bm

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            sample = col[: (CPU_SAMPLE_ROWS // 64) * BW].cpu().numpy().view(np.uint64) if n >= CPU_SAMPLE_ROWS else \
                col.cpu().numpy().view(np.uint64)
            ns = min(n, CPU_SAMPLE_ROWS)
            cb = cpu_reference_run(sample, ns, 14, 1, budget_s=20.0)
            cpu_baseline = cb
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"unavailable": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": f"cfg1 single-column BETWEEN on FLE bw{BW}, {n} rows per GPU (SF100 lineitem), "
                            f"selectivity cycling {'/'.join(map(str, SELECTIVITIES))}%, uniform codes, "
                            "row groups partitioned per rank + NCCL all-reduce of match counts",
                "rows_per_gpu": n, "bit_width": BW, "selectivities_pct": list(SELECTIVITIES),
                "l2_policy": f"input {alg_bytes / 1e6:.0f} MB per step > 126 MB L2, no flush needed",
                "parallelism": f"rowgroup-partition x{world}",
                "eval_variant": ctx.get_option("eval_variant"),
            },
            "encoded_GBps": alg_bytes * world / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": 971.3e6, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                         "profiles/r01_eval1_bw12_ncu_full_summary.txt (900.1 MB + 71.2 MB; algorithmic 975.1 MB)",
                         "peak_source": peak_src, "kernel": "fle::eval1_kernel<IS_LT=true, K=2, BW=12> (TMA-staged single-pass BETWEEN, 1 CTA per SM x 3 stages of 48 KB)",
                         "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": mean_kern_ms,
                         "min_launch_ms": min(kern_ms), "frac_of_nominal_8TBps": achieved / 8000.0},
            "q6": {"ms": q6_ms, "scan_ms": q6_scan_ms, "two_kernel_ms": q6_two_ms, "rows_per_s": n * world / (q6_ms * 1e-3),
                   "jit_launches": ctx.get_option("jit_launches"),
                   "what": "3-column 5-leaf scan + SUM(extendedprice*discount) over survivors (exact int64) in one "
                           "kernel specialised for the program shape at run time (two_kernel_ms: predicate kernel "
                           "then aggregate kernel; scan_ms: predicate kernel alone)"
                           + (" + all-reduce{count,revenue}" if world > 1 else ""),
                   "survivors_rank0": q6_count, "revenue": q6_revenue_cents2 / 1e4,
                   "revenue_exact_int64_1e-4_dollars": q6_revenue_cents2,
                   "scan_achieved_GBps_per_gpu": q6_bytes / (q6_scan_ms * 1e-3) / 1e9,
                   "scan_frac_of_peak": q6_bytes / (q6_scan_ms * 1e-3) / 1e9 / peak, "steps": q6_steps},
            "clocks": sampler.summary(),
            "gpu_launches": launches,
            "e2e": e2e,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()

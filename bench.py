#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (see DESIGN.md "Measurement").

A *step* is one fused predicate pass over one rank's synthetic lineitem shard:
  workload cfg1 (BASELINE.json configs[1]): single-column BETWEEN on an FLE-encoded int32 column,
  bw 12, N = 600,037,902 rows per GPU (SF100 lineitem), selectivity cycling 1/10/20/30/50/80/100 %
  over the steps.  For N > 1 every rank scans its own shard (weak scaling) and every step's match
  count is summed over the GPUs INSIDE the scan kernel (fle_eval_group_dev: the kernel's last CTA
  posts the count into the peers' mailboxes over NVLink and adds up theirs) -- no NCCL call and no
  host work per step beyond the one kernel launch.

The same run also measures (sub-records of the one JSON line):
  q6        TPC-H Q6 compound scan + exact int64 revenue, SF100 per GPU (cfg2)
  cfg4      SF1000-shaped lineitem (5,999,989,709 rows in row groups of 4,194,304 rows) cut into
            contiguous row-group ranges per GPU by partition.rank_rows, Q6 end to end in ONE kernel
            per GPU incl. the {count, revenue} reduction; strong scaling; totals asserted identical
            at every N (and equal to a torch compare of the same data)
  codec     cfg0: FLE pack / unpack at the reference benchmark's size (16 Mi values) and 2^26, cold L2
  in_lists  cfg3: IN lists of 8/64/256 literals on bw 12 / 16 and Q12-/Q14-/OR-style trees
  e2e       the headline metric through the C ABI with HOST (pinned) buffers, copies timed, next to
            a pinned H2D-only copy of the same bytes (the PCIe roofline of that call)
  cpu_baseline  the reference's own fle-encoding.h (oracle/_ref) on this box's host cores, all rows

value      rows/s, whole job, inputs resident in HBM, CUDA events, max over ranks
roofline   dominant kernel (fused eval): algorithmic bytes / mean launch duration vs the measured HBM
           copy peak (MEASURED_PEAKS.json)
--impl reference: only the CPU arm, as its own JSON line (all 600,037,902 rows per pass).
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
SF1000_ROWS = 5_999_989_709   # cfg4
BW = 12
SELECTIVITIES = (1, 10, 20, 30, 50, 80, 100)
METRIC = "rows_per_sec_filtered"
UNIT = "rows/s"
# cfg4 totals (seeded per row group => independent of the GPU count; profiles/r01_sf1000_q6_scaling.txt)
CFG4_EXPECT = {"survivors": 108_760_274, "revenue_1e-4_dollars": 3_453_552_581_274_335}
Q6_COLS = ((2526, 12, 730, 1094), (11, 4, 5, 7), (50, 6, 0, 22))  # (domain, bit width, lo, hi): shipdate, discount, quantity


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
    """Samples SM clock and throttle reasons through NVML while the measured regions run (a thread;
    the NVML calls go through ctypes, which releases the GIL for their duration)."""

    def __init__(self, index: int, period_s: float = 0.005):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.index = index
        self.period_s = period_s
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
            time.sleep(self.period_s)

    def start(self):
        if self.nv and self._t is None:
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
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sampled_over": "warm-up + timed steps of the headline loop and of the Q6 loops (5 ms period)"}


# --------------------------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------------------------
def cpu_reference_run(words: np.ndarray, n_rows: int, steps: int, warmup: int, budget_s: float = 20.0):
    """Times the reference's FleDecoder::Ge + Le + `&=` (BETWEEN as And(Ge, Le)) on host cores,
    one decoder per thread over disjoint row-group ranges.  Returns dict for the JSON line."""
    from oracle.oracle import GE, LE, Ref, ref_available
    cores = os.cpu_count() or 1
    if ref_available():
        ref = Ref()
        kind = "reference"
    else:
        raise RuntimeError("oracle/_ref/libfle_ref.so missing; build it where /root/reference is mounted")
    padded = np.concatenate([words, np.zeros(16, dtype=np.uint64)])
    times, count = [], 0
    t_start = time.time()
    it = 0
    while it < warmup + steps and (time.time() - t_start < budget_s or it < warmup + 1):
        sel = SELECTIVITIES[it % len(SELECTIVITIES)]
        lo, hi = between_bounds(sel, 1 << BW)
        s, count = ref.time_and_leaves([(padded, BW)], [(GE, 0, lo), (LE, 0, hi)], n_rows, cores, mode=0, reps=1)
        if it >= warmup:
            times.append(s)
        it += 1
    mean_s = sum(times) / len(times)
    # scanner-faithful variant: 1024-row windows + Skip (hdfs-parquet-scanner.cc:1837-1865)
    lo, hi = between_bounds(10, 1 << BW)
    n_small = min(n_rows, 1 << 27)
    s_win, _ = ref.time_and_leaves([(padded, BW)], [(GE, 0, lo), (LE, 0, hi)], n_small, cores, mode=1, reps=1)
    s_1t, _ = ref.time_and_leaves([(padded, BW)], [(GE, 0, lo), (LE, 0, hi)], min(n_rows, 1 << 25), 1, mode=0, reps=1)
    out = {
        "value": n_rows / mean_s, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"all {n_rows} rows of the bw{BW} column per pass, BETWEEN = Ge+Le+&= whole-range calls in 2^20-row chunks, "
                  f"one FleDecoder per thread, {len(times)} passes, selectivity cycling {'/'.join(map(str, SELECTIVITIES))}%",
        "ms_per_pass": mean_s * 1e3,
        "windowed_1024_rows_per_s": n_small / s_win,
        "single_thread_rows_per_s": min(n_rows, 1 << 25) / s_1t,
        "steps_timed": len(times),
    }
    # the reference's Put/Flush and Get on all cores, and the scalar restatement
    try:
        out["codec"] = cpu_codec_baselines(ref, cores)
    except Exception as e:  # pragma: no cover
        out["codec"] = {"unavailable": str(e)[:160]}
    return out


def cpu_codec_baselines(ref, cores: int):
    """SURVEY 8d (iii)/(iv), fle-benchmark's own shape (16 Mi values): the reference's FleEncoder::Put x n + Flush
    and FleDecoder::Get x n (AVX2 paths, fle-encoding.h compiled unmodified) on all host cores, one encoder /
    decoder per thread over disjoint ranges; and the scalar formulas of fle-benchmark.cc:395-411 (TestNativeEncode)
    and :458-474 (TestNativeDecode) as restated in oracle/fle_oracle.c, one core, on a 4 Mi-value sample."""
    from oracle.oracle import GE, Oracle
    n, m = 1 << 24, 1 << 22
    rng = np.random.default_rng(0xF1E00000)
    out = {"values": n, "cores": cores, "kind": "reference (AVX2 Put/Flush, Get) + port (scalar formulas, 1 core, 4 Mi values)",
           "unit": "values/s"}
    orc = Oracle()
    for bw in (1, 8, 12, 16, 32):
        vals = rng.integers(0, 1 << bw, size=n, dtype=np.uint64).astype(np.uint32)
        words = orc.pack(vals, bw)
        padded = np.concatenate([words, np.zeros(16, dtype=np.uint64)])
        r = {"avx2_pack_all_cores": n / ref.time_pack(vals, bw, cores, 2),
             "avx2_unpack_all_cores": n / ref.time_unpack(padded, n, bw, cores, 2)[0],
             "avx2_pack_1core": m / ref.time_pack(vals[:m], bw, 1),
             "avx2_unpack_1core": m / ref.time_unpack(padded, m, bw, 1)[0]}
        t = time.perf_counter(); orc.pack(vals[:m], bw); r["scalar_pack_1core"] = m / (time.perf_counter() - t)
        t = time.perf_counter(); orc.unpack(words, bw, 0, m); r["scalar_unpack_1core"] = m / (time.perf_counter() - t)
        t = time.perf_counter(); orc.pred(GE, words, bw, 0, m, [(1 << bw) // 2]); r["scalar_compare_rows_per_s_1core"] = m / (time.perf_counter() - t)
        out[f"bw{bw}"] = r
    return out


def make_cpu_column(seed: int, n: int):
    """The synthetic bw-12 column generated and packed on the host (numpy + the C oracle's pack) so
    that the CPU arm needs no GPU.  Chunked to bound the temporary."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(seed)
    orc = Oracle()
    parts = []
    chunk = 1 << 27
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        vals = rng.integers(0, 1 << BW, size=m, dtype=np.uint32)
        parts.append(orc.pack(vals, BW))
    return np.concatenate(parts)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.rows
    try:
        words = make_cpu_column(0xF1E00100, n)
        cb = cpu_reference_run(words, n, args.steps, args.warmup, budget_s=150.0)
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:200]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": cb["steps_timed"], "warmup": args.warmup, "ms_per_step": cb["ms_per_pass"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"cfg1 single-column BETWEEN on FLE bw{BW}, {n} rows (SF100 lineitem), uniform codes; each step = "
                               f"one pass over all {n} rows on the host cores (reference fle-encoding.h compiled unmodified)",
                   "rows_per_step": n, "rows_per_gpu": n, "bit_width": BW, "selectivities_pct": list(SELECTIVITIES)},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "windowed_1024_rows_per_s": cb["windowed_1024_rows_per_s"],
        "single_thread_rows_per_s": cb["single_thread_rows_per_s"],
        "codec": cb.get("codec"),
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import fle_b200
        self.torch, self.dist, self.fle = torch, dist, fle_b200
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        self.ctx = fle_b200.FleContext(self.local_rank, self.stream.cuda_stream)
        for o in args.opt:
            k, v = o.split("=")
            self.ctx.set_option(k, int(v))
        self.peak, self.peak_src = measured_peak()
        self.reduce_impl = "none (one GPU)"
        self.group = False
        if self.world > 1:
            self.open_group()
        self.flush_buf = None

    # -- plumbing -------------------------------------------------------------------------------
    def open_group(self):
        """One process per GPU: exchange the CUDA-IPC handles of the mailboxes over torch.distributed
        and open the group; the per-step reductions then need no collective library at all."""
        torch, dist = self.torch, self.dist
        if self.args.reduce == "nccl":
            self.reduce_impl = "ncclAllReduce per step, async (--reduce nccl)"
            return
        try:
            h = self.ctx.group_export()
            mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(self.dev)
            alls = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(alls, mine)
            self.ctx.group_open(self.rank, self.world, [bytes(t.cpu().numpy().tobytes()) for t in alls])
            ok = torch.ones(1, device=self.dev)
        except Exception as e:  # pragma: no cover
            print(f"[bench rank {self.rank}] fle_group_open failed: {e}", file=sys.stderr)
            ok = torch.zeros(1, device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() > 0:
            self.group = True
            self.ctx.set_option("group_deferred", 0 if self.args.reduce == "group-immediate" else 1)
            self.reduce_impl = ("in-kernel: last CTA of the scan posts {count, sum} to every peer's mailbox with peer "
                                "stores over NVLink and sums the peers' entries (fle_eval_group_dev, CUDA-IPC mailboxes; "
                                + ("collected inside the same kernel)" if self.args.reduce == "group-immediate" else
                                   "collected by the next query's kernel, the last one by fle_group_flush inside the timed region)"))
        else:  # pragma: no cover
            try:
                self.ctx.group_close()
            except Exception:
                pass
            self.reduce_impl = "ncclAllReduce per step, async (fle_group_open unavailable on this box)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def gather_ranks(self, val: float):
        if self.world == 1:
            return [val]
        t = self.torch.tensor([val], device=self.dev, dtype=self.torch.float64)
        alls = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(alls, t)
        return [float(x.item()) for x in alls]

    def flush_l2(self):
        """Between cold-timed launches: overwrite a buffer larger than twice the 126 MB L2, then READ a second buffer of
        twice the L2 that is clean in DRAM.  The overwrite alone leaves up to 126 MB of DIRTY lines behind, whose
        write-back the next kernel pays for (~19 us at the HBM rate -- more than the whole roofline time of a 16 Mi-value
        codec kernel); after the read pass the L2 holds clean lines of a buffer nobody uses."""
        torch = self.torch
        if self.flush_buf is None:
            self.flush_buf = torch.empty(384 << 20, dtype=torch.uint8, device=self.dev)
            self.flush_clean = torch.zeros(256 << 17, dtype=torch.int64, device=self.dev)  # 256 MB
            self.flush_sink = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.flush_buf.zero_()
        self.flush_clean.sum()

    def cold_ms(self, fn, reps=7, warm=2):
        """Median launch time with the L2 flushed before every timed launch (CUDA events on the stream)."""
        torch = self.torch
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            self.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            fn()
            b.record(self.stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts), min(ts)

    def loop_ms(self, fn, steps, warm=3):
        torch = self.torch
        for _ in range(warm):
            fn()
        if self.group:
            self.ctx.group_flush()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for _ in range(steps):
            fn()
        if self.group:
            self.ctx.group_flush()  # deferred mode: the last query's totals are collected inside the timed region
        b.record(self.stream)
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b) / steps)[0]

    # -- cfg1 headline --------------------------------------------------------------------------
    def run_headline(self):
        torch, dist, fle = self.torch, self.dist, self.fle
        from fle_b200 import AND, GE, LE
        args, ctx, dev, world = self.args, self.ctx, self.dev, self.world
        n = args.rows
        G = (n + 63) // 64
        gen = torch.Generator(device=dev).manual_seed(0xF1E00100 + self.rank)
        vals = torch.randint(0, 1 << BW, (n,), generator=gen, device=dev, dtype=torch.int32)
        col = ctx.pack_dev(vals, BW)
        torch.cuda.synchronize()
        want_counts = {}
        for sel in SELECTIVITIES:
            lo, hi = between_bounds(sel, 1 << BW)
            want_counts[sel] = int(((vals >= lo) & (vals <= hi)).sum().item())
        # full-bitmap parity of the headline kernel against a torch-packed bitmap of the same compare
        lo, hi = between_bounds(30, 1 << BW)
        self.headline_bitmap_check = self.full_bitmap_equal([(col, BW)], [(GE, 0, lo), (LE, 0, hi), (AND, None, None)], n,
                                                            (vals >= lo) & (vals <= hi))
        del vals
        torch.cuda.empty_cache()
        self.col, self.want_counts = col, want_counts

        bitmap = torch.empty(G, dtype=torch.int64, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        progs = {}
        for sel in SELECTIVITIES:
            lo, hi = between_bounds(sel, 1 << BW)
            progs[sel] = [(GE, 0, lo), (LE, 0, hi), (AND, None, None)]
        cols = [(col, BW)]
        prepared = {sel: (ctx.make_col_descs(cols),) + fle.capi._build_program(progs[sel]) for sel in SELECTIVITIES}
        self.prepared, self.bitmap = prepared, bitmap
        steps = args.steps
        count_ring = torch.zeros(max(steps, 16), dtype=torch.int64, device=dev)
        pending = []

        def step(i, mode):
            """mode: 'local' (no reduction), 'group' (in-kernel reduction), 'nccl' (async ncclAllReduce)"""
            sel = SELECTIVITIES[i % len(SELECTIVITIES)]
            c = count_ring[i % count_ring.numel():i % count_ring.numel() + 1]
            if mode == "group":
                ctx.eval_group_dev(cols, progs[sel], n, bitmap=bitmap, count=c, prepared=prepared[sel])
            else:
                ctx.eval_dev(cols, progs[sel], n, bitmap=bitmap, count=c, prepared=prepared[sel])
                if mode == "nccl":
                    pending.append(dist.all_reduce(c, async_op=True))
            return sel

        def want_totals(k):
            t = torch.tensor([want_counts[SELECTIVITIES[i % len(SELECTIVITIES)]] for i in range(k)], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(t)
            return t

        main_mode = "local" if world == 1 else ("group" if self.group else "nccl")
        sampler = self.sampler = ClockSampler(self.local_rank)
        sampler.start()
        # warm-up (+ correctness of every selectivity once, reductions included)
        nwarm = max(args.warmup, len(SELECTIVITIES))
        for i in range(nwarm):
            step(i, main_mode)
        for w in pending:
            w.wait()
        pending.clear()
        if main_mode == "group":
            ctx.group_flush()
        torch.cuda.synchronize()
        assert torch.equal(count_ring[:nwarm], want_totals(nwarm)[:nwarm]), "warm-up counts differ from the torch compare"

        def timed_loop(mode):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            self.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(self.stream)
            for i in range(steps):
                ev[i][0].record(self.stream)
                step(i, mode)
                ev[i][1].record(self.stream)
            for w in pending:
                w.wait()  # all reduced counts are complete before the clock stops
            pending.clear()
            if mode == "group":
                ctx.group_flush()  # ... in deferred mode too: the last query's totals are collected here
            t1.record(self.stream)
            self.barrier()
            return t0.elapsed_time(t1), [a.elapsed_time(b) for a, b in ev]

        launches0 = ctx.launch_count
        elapsed_ms, kern_ms = timed_loop(main_mode)
        launches = ctx.launch_count - launches0
        if world > 1:
            assert torch.equal(count_ring[:steps], want_totals(steps)), "reduced match counts differ from the torch compare"
            if self.group:
                st = ctx.group_status()
                assert st["peers_timed_out"] == 0, st
        else:
            assert torch.equal(count_ring[:steps], want_totals(steps))
        per_rank = self.gather_ranks(elapsed_ms)
        worst_step = self.gather_ranks(max(kern_ms))
        elapsed_max = max(per_rank)
        ms_per_step = elapsed_max / steps
        self.headline = {
            "ms_per_step": ms_per_step, "value": n * world / (ms_per_step * 1e-3), "launches": launches,
            "mean_kern_ms": sum(kern_ms) / len(kern_ms), "min_kern_ms": min(kern_ms),
            "per_rank_elapsed_ms": per_rank, "per_rank_worst_step_ms": worst_step,
        }
        # attribution of the multi-GPU step: the same loop without any reduction, and with the NCCL path
        if world > 1:
            loc_ms, loc_k = timed_loop("local")
            loc = max(self.gather_ranks(loc_ms)) / steps
            self.headline["scan_only_ms_per_step"] = loc
            self.headline["reduce_exposed_ms_per_step"] = ms_per_step - loc
            if main_mode == "group" and not args.no_nccl_compare:
                nc_ms, _ = timed_loop("nccl")
                self.headline["nccl_async_allreduce_ms_per_step"] = max(self.gather_ranks(nc_ms)) / steps

    def full_bitmap_equal(self, cols, prog, n, mask_bool):
        """Every word of the kernel's bitmap against a bitmap packed by torch from the boolean compare."""
        torch = self.torch
        G = (n + 63) // 64
        bm, cnt = self.ctx.eval_dev(cols, prog, n)
        pad = G * 64 - n
        m = mask_bool
        if pad:
            m = torch.cat([m, torch.zeros(pad, dtype=torch.bool, device=m.device)])
        want = torch.zeros(G, dtype=torch.int64, device=m.device)
        m = m.view(G, 64)
        for b in range(64):  # row r -> bit r & 63 of word r >> 6
            want |= m[:, b].to(torch.int64) << b
        torch.cuda.synchronize()
        ok = bool(torch.equal(bm[:G], want)) and int(cnt.item()) == int(mask_bool.sum().item())
        assert ok, "kernel bitmap differs from the torch-packed bitmap"
        return {"rows": n, "words_compared": G, "equal": ok}

    # -- cfg2: Q6 at SF100 per GPU --------------------------------------------------------------------
    def run_q6(self):
        torch, dist, fle = self.torch, self.dist, self.fle
        from fle_b200 import AND, GE, LE, LT
        args, ctx, dev, world = self.args, self.ctx, self.dev, self.world
        n = args.rows
        G = (n + 63) // 64
        gen = torch.Generator(device=dev).manual_seed(0xF1E00200 + self.rank)
        q6_cols, q6_mask = [], None
        price = torch.randint(90_000, 10_500_000, (n,), generator=gen, device=dev, dtype=torch.int64)  # l_extendedprice, cents (PLAIN)
        disc_dict = torch.arange(0, 11, device=dev, dtype=torch.int64)                                  # l_discount dictionary, cents
        for d, bw, lo_, hi_ in Q6_COLS:
            v = torch.randint(0, d, (n,), generator=gen, device=dev, dtype=torch.int32)
            q6_cols.append((ctx.pack_dev(v, bw), bw))
            m_ = (v >= lo_) & (v <= hi_)
            q6_mask = m_ if q6_mask is None else (q6_mask & m_)
            if d == 11:
                disc_vals = v
            else:
                del v
        want_count = int(q6_mask.sum().item())
        want_revenue = int((price[q6_mask] * disc_dict[disc_vals[q6_mask].long()]).sum().item())
        q6_prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
                   (AND, None, None), (LT, 2, 23), (AND, None, None)]
        bm_check = self.full_bitmap_equal(q6_cols, q6_prog, n, q6_mask)
        del q6_mask, disc_vals, m_
        torch.cuda.empty_cache()
        want = torch.tensor([want_revenue, want_count], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(want)
        prep = (ctx.make_col_descs(q6_cols),) + fle.capi._build_program(q6_prog)
        agg = torch.zeros(2, dtype=torch.int64, device=dev)
        bitmap, count = self.bitmap, torch.zeros(1, dtype=torch.int64, device=dev)

        def q6_step():
            # predicate + aggregate (+ the cross-GPU sum) in ONE kernel specialised for the program shape
            if self.group:
                ctx.eval_sum_product_group_dev(q6_cols, q6_prog, n, price, 1, disc_dict, bitmap=bitmap, out=agg, prepared=prep)
            else:
                ctx.eval_sum_product_dev(q6_cols, q6_prog, n, price, 1, disc_dict, bitmap=bitmap, out=agg, prepared=prep)
                if world > 1:
                    dist.all_reduce(agg)

        def q6_two_kernel_step():
            ctx.eval_dev(q6_cols, q6_prog, n, bitmap=bitmap, count=count, prepared=prep)
            ctx.sum_product_dev(bitmap, n, price, q6_cols[1][0], 4, disc_dict, out=agg)

        q6_steps = max(10, args.steps // 2)
        q6_ms = self.loop_ms(q6_step, q6_steps)
        assert torch.equal(agg, want), "Q6 {revenue, count} differ from the torch compare"
        two_ms = self.loop_ms(q6_two_kernel_step, q6_steps)
        if world == 1:
            assert torch.equal(agg, want) and int(count.item()) == want_count
        scan_ms = self.loop_ms(lambda: ctx.eval_dev(q6_cols, q6_prog, n, bitmap=bitmap, count=count, prepared=prep), q6_steps)
        q6_bytes = G * (12 + 4 + 6) * 8 + G * 8
        # the same scan as an Impala-written file presents it: every column OPTIONAL with a def-level stream (1-bit FLE,
        # hdfs-parquet-table-writer.cc:103-105) that holds no NULL; checked and evaluated without a host sync
        optional = None
        if world == 1 and hasattr(ctx, "eval_no_nulls_dev"):
            ones = torch.full((G,), -1, dtype=torch.int64, device=dev)
            ocols = [c + (ones,) for c in q6_cols]
            oprep = (ctx.make_col_descs(ocols),) + fle.capi._build_program(q6_prog)
            o2 = torch.zeros(2, dtype=torch.int64, device=dev)
            opt_ms = self.loop_ms(lambda: ctx.eval_no_nulls_dev(ocols, q6_prog, n, bitmap=bitmap, out=o2, prepared=oprep), q6_steps)
            assert int(o2[0].item()) == want_count and int(o2[1].item()) == 0
            obytes = q6_bytes + 3 * G * 8
            optional = {"ms": opt_ms, "GBps": obytes / (opt_ms * 1e-3) / 1e9, "frac_of_peak": obytes / (opt_ms * 1e-3) / 1e9 / self.peak,
                        "what": "scan_ms with three OPTIONAL columns through fle_eval_no_nulls_dev: def-level count kernel + fused scan, "
                                "NULL count left on the device (no host sync); bytes = the scan's + 3 bits per row of def levels"}
            del ones, ocols
        # ... and with NULLs actually present (2 % per column): ColumnReader<T>::Eq..In + IntersectBitset
        # (hdfs-parquet-scanner.cc:326-345) = tile counts, prefixes and ONE chain kernel that evaluates the data streams,
        # deposits through the def levels and folds the conjuncts; every word checked against a torch-packed bitmap
        with_nulls = None
        if world == 1:
            ncols, nmask, nbytes = [], None, G * 8
            for d, bw, lo_, hi_ in Q6_COLS:
                nn = torch.rand(n, generator=gen, device=dev) >= 0.02
                data = torch.randint(0, d, (int(nn.sum().item()),), generator=gen, device=dev, dtype=torch.int32)
                leaf = torch.zeros(n, dtype=torch.bool, device=dev)
                leaf[nn] = (data >= lo_) & (data <= hi_)
                nmask = leaf if nmask is None else (nmask & leaf)
                ncols.append((ctx.pack_dev(data, bw), bw, ctx.pack_dev(nn.to(torch.int32), 1)))
                nbytes += ((data.numel() + 63) // 64) * bw * 8 + G * 8
                del nn, data, leaf
            ctx.eval_dev(ncols, q6_prog, n, bitmap=bitmap, count=count)  # (the first call learns that the file holds NULLs)
            null_check = self.full_bitmap_equal(ncols, q6_prog, n, nmask)
            del nmask
            nprep = (ctx.make_col_descs(ncols),) + fle.capi._build_program(q6_prog)
            l0 = ctx.launch_count
            ctx.eval_dev(ncols, q6_prog, n, bitmap=bitmap, count=count, prepared=nprep)
            per_call = ctx.launch_count - l0
            null_ms = self.loop_ms(lambda: ctx.eval_dev(ncols, q6_prog, n, bitmap=bitmap, count=count, prepared=nprep), q6_steps)
            with_nulls = {"ms": null_ms, "null_fraction": 0.02, "launches_per_call": per_call, "GBps": nbytes / (null_ms * 1e-3) / 1e9,
                          "frac_of_peak": nbytes / (null_ms * 1e-3) / 1e9 / self.peak, "full_bitmap_check": null_check,
                          "what": "scan_ms with three OPTIONAL columns that hold 2 % NULLs each, through fle_eval_dev (its host sync "
                                  "included); bytes = data streams + def levels + bitmap, each once"}
            del ncols
            torch.cuda.empty_cache()
        self.q6 = {
            "ms": q6_ms, "scan_ms": scan_ms, "two_kernel_ms": two_ms, "rows_per_s": n * world / (q6_ms * 1e-3),
            "jit_launches": ctx.get_option("jit_launches"),
            "what": "cfg2, SF100 per GPU: 3-column 5-leaf scan + SUM(extendedprice*discount) over survivors (exact int64) in one "
                    "kernel specialised for the program shape at run time (two_kernel_ms: predicate kernel then aggregate "
                    "kernel; scan_ms: predicate kernel alone)" + (f"; totals over {world} GPUs: {self.reduce_impl}" if world > 1 else ""),
            "survivors_total": int(want[1].item()), "revenue": int(want[0].item()) / 1e4,
            "revenue_exact_int64_1e-4_dollars": int(want[0].item()),
            "scan_achieved_GBps_per_gpu": q6_bytes / (scan_ms * 1e-3) / 1e9,
            "scan_frac_of_peak": q6_bytes / (scan_ms * 1e-3) / 1e9 / self.peak, "steps": q6_steps,
            "full_bitmap_check": bm_check, "optional_columns_no_host_sync": optional, "optional_columns_2pct_nulls": with_nulls,
        }
        del price, q6_cols, agg
        torch.cuda.empty_cache()

    # -- e2e: host buffers through the C ABI -----------------------------------------------------------
    def run_e2e(self):
        import ctypes as C
        torch, fle = self.torch, self.fle
        args, ctx, world = self.args, self.ctx, self.world
        n = args.rows
        G = (n + 63) // 64
        col = self.col
        host_words = torch.empty(col.numel(), dtype=torch.int64, pin_memory=True)
        host_words.copy_(col)
        host_bm = torch.empty(G, dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize()
        hw = host_words.numpy().view(np.uint64)
        hb = host_bm.numpy().view(np.uint64)
        descs = (fle.capi.ColDesc * 1)()
        descs[0].planes = hw.ctypes.data
        descs[0].bytes = hw.nbytes
        descs[0].bit_width = BW
        lib = ctx.lib
        cnt = C.c_uint64(0)

        def e2e_step(i):
            sel = SELECTIVITIES[i % len(SELECTIVITIES)]
            nodes = self.prepared[sel][1]
            rc = lib.fle_eval(ctx.h, descs, 1, nodes, len(nodes), 0, n, hb.ctypes.data, C.byref(cnt))
            assert rc == 0, lib.fle_last_error(ctx.h)
            assert cnt.value == self.want_counts[sel]

        e2e_step(1)
        self.barrier()
        w0 = time.perf_counter()
        for i in range(args.e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = self.max_over_ranks((time.perf_counter() - w0) / args.e2e_steps)[0]
        # the PCIe roofline of that call on this box: the same bytes, pinned -> device, nothing else
        dst = torch.empty_like(col)
        for _ in range(2):
            dst.copy_(host_words, non_blocking=True)
        self.barrier()
        w0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(host_words, non_blocking=True)
        torch.cuda.synchronize()
        h2d_s = self.max_over_ranks((time.perf_counter() - w0) / 3)[0]
        self.e2e = {"value": n * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(hw.nbytes),
                    "d2h_bytes_per_step": int(hb.nbytes + 8), "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
                    "timing": "host wall clock around fle_eval() (call returns with the result in host memory)",
                    "achieved_h2d_GBps_per_gpu": hw.nbytes / e2e_s / 1e9,
                    "h2d_peak_GBps": hw.nbytes / h2d_s / 1e9,
                    "h2d_peak_what": f"pinned->device cudaMemcpyAsync of the same {hw.nbytes} bytes alone, all {world} GPUs at once, per GPU",
                    "frac_of_h2d_peak": h2d_s / e2e_s}
        del host_words, host_bm, dst

    # -- the column as it sits inside a file image: FLE section at an odd byte offset, evaluated in place -------------
    def run_page_image(self):
        torch, ctx, dev = self.torch, self.ctx, self.dev
        from fle_b200 import AND, GE, LE
        n = self.args.rows
        G = (n + 63) // 64
        nbytes = self.col.numel() * 8
        image = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
        off = 17  # one byte after a 16-byte boundary: the page's width byte precedes the groups (dict-encoding.h:185-192)
        image[off:off + nbytes].copy_(self.col.view(torch.uint8))

        class Raw:
            device = dev
            def data_ptr(self_): return image.data_ptr() + off
            def numel(self_): return nbytes
            def element_size(self_): return 1
        lo, hi = between_bounds(30, 1 << BW)
        prog = [(GE, 0, lo), (LE, 0, hi), (AND, None, None)]
        cols = [(Raw(), BW)]
        prep = (ctx.make_col_descs(cols),) + self.fle.capi._build_program(prog)
        bm = torch.empty(G, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        fn = lambda: ctx.eval_dev(cols, prog, n, bitmap=bm, count=cnt, prepared=prep)  # noqa: E731
        l0 = ctx.launch_count
        fn()
        torch.cuda.synchronize()
        per_call = ctx.launch_count - l0
        ref_bm, ref_cnt = ctx.eval_dev([(self.col, BW)], prog, n)
        torch.cuda.synchronize()
        assert torch.equal(bm, ref_bm) and int(cnt.item()) == int(ref_cnt.item()) == self.want_counts[30]
        med, mn = self.cold_ms(fn)
        by = G * BW * 8 + G * 8
        self.page_image = {"what": f"cfg1 BETWEEN on the same bw{BW} column placed at byte offset {off} of a device buffer (as the FLE section "
                                   "of a page inside a file image), evaluated in place by the skew-capable kernel -- no realigned copy; "
                                   "bitmap identical to the aligned column's", "ms": med, "ms_min": mn,
                           "GBps": by / (med * 1e-3) / 1e9, "frac": by / (med * 1e-3) / 1e9 / self.peak, "launches_per_call": per_call}
        del image, bm

    # -- cfg0 / cfg3 sub-records (one GPU) ----------------------------------------------------------------
    def run_codec(self):
        torch, ctx, dev = self.torch, self.ctx, self.dev
        out = {"what": "cfg0: FLE pack / unpack, uniform values, L2 flushed (384 MB overwrite, then a 256 MB read of a clean buffer so that no dirty lines are left) before every timed launch, median of 7; "
                       "frac = (bw*8 + 64*sizeof(out)) bytes per group / time / measured HBM peak", "records": []}
        gen = torch.Generator(device=dev).manual_seed(0xF1E00000)
        for n in (1 << 24, 1 << 26):
            G = n // 64
            for bw in (1, 4, 8, 12, 16, 24, 32):
                v = torch.randint(0, 1 << bw, (n,), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
                w = ctx.pack_dev(v, bw)
                o = torch.empty(n, dtype=torch.int32, device=dev)
                w2 = torch.empty_like(w)
                rec = {"values": n, "bw": bw}
                for name, fn, by in (("unpack", lambda: ctx.unpack_dev(w, bw, 0, n, out=o), G * bw * 8 + n * 4),
                                     ("pack", lambda: ctx.pack_dev(v, bw, out=w2), G * bw * 8 + n * 4)):
                    med, mn = self.cold_ms(fn)
                    rec[name] = {"ms": med, "ms_min": mn, "GBps": by / (med * 1e-3) / 1e9, "frac": by / (med * 1e-3) / 1e9 / self.peak,
                                 "values_per_s": n / (med * 1e-3)}
                assert torch.equal(o, v) and torch.equal(w2, w)
                if hasattr(ctx, "unpack_narrow_dev") and bw <= 16:
                    es = 1 if bw <= 8 else 2
                    on = torch.empty(n, dtype=torch.uint8 if es == 1 else torch.int16, device=dev)
                    med, mn = self.cold_ms(lambda: ctx.unpack_narrow_dev(w, bw, 0, n, out=on))
                    by = G * bw * 8 + n * es
                    rec[f"unpack_u{8 * es}"] = {"ms": med, "ms_min": mn, "GBps": by / (med * 1e-3) / 1e9,
                                                "frac": by / (med * 1e-3) / 1e9 / self.peak, "values_per_s": n / (med * 1e-3)}
                    assert torch.equal(on.to(torch.int32) & ((1 << (8 * es)) - 1), v)
                out["records"].append(rec)
                del v, w, o, w2
        # the writer side of the dictionary: DictEncoder<T>::Put x n + WriteDict + WriteData as one device call
        # (dict-encoding.h:234-274, :393-447) -- values in, sorted dictionary + FLE planes out; includes its host sync
        try:
            n = 1 << 26
            x = torch.randint(0, 2526, (n,), generator=gen, device=dev, dtype=torch.int64) * 7 - 1000
            d, bw, enc = ctx.dict_encode_dev(x)
            assert d.numel() == 2526 and bw == 12
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(self.stream)
                ctx.dict_encode_dev(x)
                b.record(self.stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            by = n * 8 + enc.numel() * 8
            out["dict_build"] = {"values": n, "distinct": 2526, "elem": "int64", "ms": ms, "values_per_s": n / (ms * 1e-3), "GBps": by / (ms * 1e-3) / 1e9,
                                 "frac": by / (ms * 1e-3) / 1e9 / self.peak,
                                 "what": "fle_dict_encode_dev: hash dedupe behind per-CTA shared-memory sets, rank, value -> code through a "
                                         "shared-memory map, pack; frac = (values read once + planes written once) / time / measured HBM peak"}
            del x, d, enc
        except Exception as e:  # pragma: no cover
            out["dict_build"] = {"unavailable": str(e)[:200]}
        torch.cuda.empty_cache()
        self.codec = out

    def run_in_lists(self):
        torch, ctx, dev, fle = self.torch, self.ctx, self.dev, self.fle
        from fle_b200 import AND, EQ, GE, IN, LT, OR
        n = self.args.rows
        G = (n + 63) // 64
        gen = torch.Generator(device=dev).manual_seed(0xF1E00300)
        spec = {"shipdate": (2526, 12), "wide16": (40000, 16), "shipmode": (7, 3), "instruct": (4, 2), "quantity": (50, 6),
                "receipt": (2526, 12)}
        cols, vals = {}, {}
        for k, (d, bw) in spec.items():
            v = torch.randint(0, d, (n,), generator=gen, device=dev, dtype=torch.int32)
            cols[k] = (ctx.pack_dev(v, bw), bw)
            vals[k] = v
        cg = torch.Generator().manual_seed(5)
        lit = lambda d, L: torch.randperm(d, generator=cg)[:L].tolist()  # noqa: E731  L distinct present codes
        programs = {}
        for L in (8, 64, 256):
            programs[f"in{L}_bw12"] = (["shipdate"], [(IN, 0, lit(2526, L))])
            programs[f"in{L}_bw16"] = (["wide16"], [(IN, 0, lit(40000, L))])
        programs["q12_style"] = (["shipmode", "receipt"], [(IN, 0, [1, 4]), (GE, 1, 730), (AND,), (LT, 1, 1095), (AND,)])
        programs["q14_style"] = (["shipdate"], [(GE, 0, 1339), (LT, 0, 1369), (AND,)])
        programs["or_variant"] = (["shipmode", "instruct", "quantity"], [(IN, 0, [1, 4]), (EQ, 1, 2), (OR,), (LT, 2, 23), (AND,)])
        programs["in256_and_lt"] = (["shipdate", "quantity"], [(IN, 0, lit(2526, 256)), (LT, 1, 23), (AND,)])
        bm = torch.empty(G, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        out = {"what": f"cfg3 at {n} rows (inputs > L2, no flush needed), median of 7 launches; frac = (sum bw + 1) * 8 bytes per "
                       "64 rows / time / measured HBM peak; counts checked against torch.isin / compares", "records": []}

        def torch_eval(cn, prog):
            st = []
            for p in prog:
                if p[0] == AND:
                    b = st.pop(); st[-1] = st[-1] & b
                elif p[0] == OR:
                    b = st.pop(); st[-1] = st[-1] | b
                else:
                    v = vals[cn[p[1]]]
                    if p[0] == IN:
                        st.append(torch.isin(v, torch.tensor(p[2], device=dev, dtype=torch.int32)))
                    elif p[0] == EQ:
                        st.append(v == p[2])
                    elif p[0] == GE:
                        st.append(v >= p[2])
                    elif p[0] == LT:
                        st.append(v < p[2])
            return int(st[0].sum().item())

        for name, (cn, prog) in programs.items():
            pc = [cols[c] for c in cn]
            progf = [tuple(p) + (None,) * (3 - len(p)) for p in prog]
            prep = (ctx.make_col_descs(pc),) + fle.capi._build_program(progf)
            sum_bw = sum(bw for _, bw in pc)
            fn = lambda: ctx.eval_dev(pc, progf, n, bitmap=bm, count=cnt, prepared=prep)  # noqa: E731
            l0 = ctx.launch_count
            fn()
            torch.cuda.synchronize()
            per_call = ctx.launch_count - l0
            assert int(cnt.item()) == torch_eval(cn, prog), name
            med, mn = self.cold_ms(fn)
            by = G * sum_bw * 8 + G * 8
            out["records"].append({"name": name, "sum_bw": sum_bw, "ms": med, "ms_min": mn, "rows_per_s": n / (med * 1e-3),
                                   "GBps": by / (med * 1e-3) / 1e9, "frac": by / (med * 1e-3) / 1e9 / self.peak,
                                   "launches_per_call": per_call, "count": int(cnt.item())})
        del cols, vals, bm
        torch.cuda.empty_cache()
        self.in_lists = out

    # -- cfg4: SF1000-shaped lineitem partitioned by row groups (strong scaling) ----------------------------
    def run_cfg4(self):
        torch, dist, fle = self.torch, self.dist, self.fle
        from fle_b200 import AND, GE, LE, LT, partition
        args, ctx, dev, world, rank = self.args, self.ctx, self.dev, self.world, self.rank
        total = args.cfg4_rows
        first, end = partition.rank_rows(total, world, rank)
        n = end - first
        G = (n + 63) // 64
        RG = partition.ROW_GROUP_ROWS
        planes = [torch.empty(G * bw, dtype=torch.int64, device=dev) for _, bw, _, _ in Q6_COLS]
        price = torch.empty(n, dtype=torch.int64, device=dev)
        disc_dict = torch.arange(0, 11, device=dev, dtype=torch.int64)
        want = torch.zeros(2, dtype=torch.int64, device=dev)  # revenue, count by torch compare
        t_gen = time.time()
        # shard built row group by row group, seeded per row group => the table (and so count and revenue)
        # does not depend on how many GPUs it is dealt out to
        for r0 in range(0, n, RG):
            m = min(RG, n - r0)
            gen = torch.Generator(device=dev).manual_seed(0xF1E00400 + (first + r0) // RG)
            p = torch.randint(90_000, 10_500_000, (m,), generator=gen, device=dev, dtype=torch.int64)
            price[r0:r0 + m] = p
            mask, dv = None, None
            for c, (d, bw, lo, hi) in enumerate(Q6_COLS):
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
        t_gen = time.time() - t_gen
        torch.cuda.empty_cache()
        if world > 1:
            dist.all_reduce(want)
        prog = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
                (AND, None, None), (LT, 2, 23), (AND, None, None)]
        cols = [(planes[c], Q6_COLS[c][1]) for c in range(3)]
        prep = (ctx.make_col_descs(cols),) + fle.capi._build_program(prog)
        bitmap = torch.empty(max(G, 1), dtype=torch.int64, device=dev)
        agg = torch.zeros(2, dtype=torch.int64, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)

        def q6():
            if self.group:
                ctx.eval_sum_product_group_dev(cols, prog, n, price, 1, disc_dict, bitmap=bitmap, out=agg, prepared=prep)
            else:
                ctx.eval_sum_product_dev(cols, prog, n, price, 1, disc_dict, bitmap=bitmap, out=agg, prepared=prep)
                if world > 1:
                    dist.all_reduce(agg)

        def scan():
            if self.group:
                ctx.eval_group_dev(cols, prog, n, bitmap=bitmap, count=count, prepared=prep)
            else:
                ctx.eval_dev(cols, prog, n, bitmap=bitmap, count=count, prepared=prep)
                if world > 1:
                    dist.all_reduce(count)

        steps = max(10, args.steps)
        q6_ms = self.loop_ms(q6, steps)
        ok = bool(torch.equal(agg, want))
        scan_ms = self.loop_ms(scan, steps)
        ok = ok and int(count.item()) == int(want[1].item())
        per_rank_rows = self.gather_ranks(float(n))
        expect_ok = None
        if total == SF1000_ROWS:
            expect_ok = (int(want[1].item()) == CFG4_EXPECT["survivors"] and int(want[0].item()) == CFG4_EXPECT["revenue_1e-4_dollars"])
        Gt = (total + 63) // 64
        scan_bytes = Gt * (12 + 4 + 6) * 8 + Gt * 8
        self.cfg4 = {
            "workload": f"cfg4 TPC-H Q6 over SF1000-shaped lineitem: {total} rows in row groups of {RG} rows dealt out as contiguous "
                        f"row-group ranges (partition.rank_rows) to {world} GPU(s), strong scaling; fused predicate + exact int64 "
                        "SUM(extendedprice*discount) + {revenue, count} summed over the GPUs: " + self.reduce_impl,
            "scaling": "strong", "total_rows": total, "n_gpus": world, "rows_per_rank": [int(x) for x in per_rank_rows],
            "steps": steps, "q6_ms": q6_ms, "q6_rows_per_s": total / (q6_ms * 1e-3),
            "scan_ms": scan_ms, "scan_rows_per_s": total / (scan_ms * 1e-3),
            "scan_GBps": scan_bytes / (scan_ms * 1e-3) / 1e9,
            "scan_frac_of_n_x_peak": scan_bytes / (scan_ms * 1e-3) / 1e9 / (self.peak * world),
            "survivors": int(want[1].item()), "revenue_exact_int64_1e-4_dollars": int(want[0].item()),
            "matches_torch_compare": ok, "matches_totals_of_every_other_gpu_count": expect_ok,
            "build_shard_s": t_gen,
        }
        assert ok, "cfg4: Q6 count / revenue differ from the torch compare"
        assert expect_ok in (None, True), "cfg4: totals differ from the ones every other GPU count produces"
        del planes, price, bitmap
        torch.cuda.empty_cache()

    # -- drop-in scanner path (C++ host program bound to the C ABI) ------------------------------------------
    def run_scanner(self):
        import subprocess
        exe = os.path.join(ROOT, "impala-avx2-parquet-scanner_b200", "host", "tests", "scanner-bench")
        if not os.path.exists(exe):
            self.scanner = {"unavailable": "host/tests/scanner-bench not built"}
            return
        try:
            r = subprocess.run([exe, "--json", "--threads", "1,2,3,4"], capture_output=True, text=True, timeout=300)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            self.scanner = json.loads(line[-1]) if line else {"unavailable": (r.stdout + r.stderr)[-300:]}
            if line:
                self.scanner["what"] = ("the reference-facing drop-in, host/exec/hdfs-parquet-scanner.h: HdfsParquetScanner::AssembleRows on "
                                        "Impala-shaped row groups (3 OPTIONAL dictionary columns, Q6 conjuncts, survivors materialised as "
                                        "tuples), pages in pinned host memory, H2D + D2H inside the timed region, T scanner threads; "
                                        "frac_of_h2d = against the same bytes copied pinned->device and nothing else")
        except Exception as e:  # pragma: no cover
            self.scanner = {"unavailable": str(e)[:200]}

    def cpu_baseline(self):
        try:
            words = self.col.cpu().numpy().view(np.uint64)
            return cpu_reference_run(words, self.args.rows, 14, 1, budget_s=25.0)
        except Exception as e:  # pragma: no cover
            return {"unavailable": str(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=70)
    ap.add_argument("--warmup", type=int, default=7)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows per GPU (default: SF100 lineitem)")
    ap.add_argument("--cfg4-rows", type=int, default=SF1000_ROWS, help="total rows of the cfg4 table (default: SF1000 lineitem)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the codec / in_lists / scanner sub-records")
    ap.add_argument("--no-nccl-compare", action="store_true")
    ap.add_argument("--reduce", default="group", choices=["group", "group-immediate", "nccl"], help="N > 1: how per-step totals are summed")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    b = Bench(args)
    torch = b.torch
    n, world, rank, ctx = args.rows, b.world, b.rank, b.ctx
    G = (n + 63) // 64
    b.run_headline()
    b.run_q6()
    b.sampler.stop()
    b.clocks = b.sampler.summary()
    e2e = None
    if not args.no_e2e:
        b.run_e2e()
        e2e = b.e2e
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = b.cpu_baseline()
    page_image = None
    if world == 1 and not args.no_sub:
        b.run_page_image()
        page_image = b.page_image
    del b.col
    torch.cuda.empty_cache()
    codec = in_lists = scanner = None
    if world == 1 and not args.no_sub:
        b.run_codec()
        b.run_in_lists()
        b.run_scanner()
        codec, in_lists, scanner = b.codec, b.in_lists, b.scanner
    cfg4 = None
    if not args.no_cfg4:
        b.bitmap = None
        torch.cuda.empty_cache()
        b.run_cfg4()
        cfg4 = b.cfg4

    if rank == 0:
        h = b.headline
        alg_bytes = G * BW * 8 + G * 8
        achieved = alg_bytes / (h["mean_kern_ms"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": h["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": f"cfg1 single-column BETWEEN on FLE bw{BW}, {n} rows per GPU (SF100 lineitem), "
                            f"selectivity cycling {'/'.join(map(str, SELECTIVITIES))}%, uniform codes"
                            + (", every GPU scans its own shard and every step's match count is summed over the GPUs: " + b.reduce_impl
                               if world > 1 else ""),
                "rows_per_gpu": n, "bit_width": BW, "selectivities_pct": list(SELECTIVITIES),
                "l2_policy": f"input {alg_bytes / 1e6:.0f} MB per step > 126 MB L2, no flush needed",
                "parallelism": f"rowgroup-partition x{world}",
                "eval_variant": ctx.get_option("eval_variant"),
            },
            "encoded_GBps": alg_bytes * world / (h["ms_per_step"] * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": b.peak, "unit": "GB/s", "frac": achieved / b.peak,
                         "traffic": 972.2e6,
                         "traffic_source": "CITED, not measured in this run: ncu --set full dram__bytes_read.sum + dram__bytes_write.sum "
                                           "per launch, profiles/r02_between12_ncu_full_summary.txt (900.1 MB + 72.0 MB; algorithmic 975.1 MB)",
                         "peak_source": b.peak_src,
                         "kernel": "fle::eval1_kernel<IS_LT=true, K=2, BW=12> (TMA-staged single-pass BETWEEN, 1 CTA per SM x 3 stages of 48 KB)",
                         "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": h["mean_kern_ms"],
                         "min_launch_ms": h["min_kern_ms"], "frac_of_nominal_8TBps": achieved / 8000.0},
            "parity_in_run": {"headline_full_bitmap": b.headline_bitmap_check},
            "multi_gpu": {k: h[k] for k in ("per_rank_elapsed_ms", "per_rank_worst_step_ms", "scan_only_ms_per_step",
                                            "reduce_exposed_ms_per_step", "nccl_async_allreduce_ms_per_step") if k in h}
            | {"reduce": b.reduce_impl},
            "q6": b.q6,
            "cfg4": cfg4,
            "codec": codec,
            "in_lists": in_lists,
            "page_image": page_image,
            "e2e_scanner": scanner,
            "clocks": b.clocks,
            "gpu_launches": h["launches"],
            "e2e": e2e,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        b.dist.barrier()
        if b.group:
            ctx.group_close()
        b.dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()

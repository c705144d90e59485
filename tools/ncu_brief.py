#!/usr/bin/env python
"""Prints the handful of ncu metrics the kernel notes in DESIGN.md quote:  python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_elapsed.avg.per_second"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:90])
    for i, n in enumerate(h):
        if n in WANT or ("issue_stalled" in n and n.endswith("_per_warp_active.pct") and float(r[i] or 0) > 2.0):
            print(f"  {n:95s} {r[i]:>16s} {u[i]}")

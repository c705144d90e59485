#!/usr/bin/env python
"""Prints gpurun_out/sweep.jsonl (tools/sweep.py) as tables."""
import json
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.jsonl"
rows = [json.loads(l) for l in open(path)]
b = {}
for r in rows:
    if r["what"] == "between":
        b.setdefault(r["bw"], {})[r["variant"]] = r
if b:
    print("BETWEEN  bw : staged ms GB/s frac | direct GB/s frac | rows/s")
    for bw in sorted(b):
        t, d = b[bw].get(1), b[bw].get(0)
        line = f"bw{bw:2d}"
        if t:
            line += f" tma {t['ms_min']:.4f}ms {t['GBps']:6.0f} {t['frac']:.3f}"
        if b[bw].get(2):
            line += f" | generic-staged {b[bw][2]['GBps']:6.0f} {b[bw][2]['frac']:.3f}"
        if d:
            line += f" | direct {d['GBps']:6.0f} {d['frac']:.3f}"
        if t:
            line += f" | {t['rows_per_s']:.3e}"
        print(line)
for r in rows:
    if r["what"] == "selectivity":
        print(f"sel {r['sel_pct']:3d}% {r['ms_min']:.4f}ms {r['GBps']:.0f}GB/s frac {r['frac']:.3f} obs {r['observed_sel']:.4f}")
c = {}
for r in rows:
    if r["what"] in ("pack", "unpack"):
        c.setdefault(r["bw"], {})[r["what"]] = r
for bw in sorted(c):
    u, p = c[bw].get("unpack"), c[bw].get("pack")
    print(f"bw{bw:2d} unpack {u['ms_min']:.3f}ms {u['GBps']:6.0f}GB/s {u['frac']:.3f} ok={u['ok']} | pack {p['ms_min']:.3f}ms {p['GBps']:6.0f} {p['frac']:.3f} ok={p['ok']}")
for r in rows:
    if r["what"] == "program":
        print(f"{r['name']:12s} v{r['variant']} {r['ms_min']:.4f}ms {r['GBps']:7.0f}GB/s frac {r['frac']:.3f} rows/s {r['rows_per_s']:.3e} cnt {r['count']}")
t = [r for r in rows if r["what"] == "tune" and "ms_min" in r]
for case in sorted(set(r["case"] for r in t)):
    best = sorted([r for r in t if r["case"] == case], key=lambda r: r["ms_min"])[:6]
    for r in best:
        print(f"tune {case:10s} block {r['block']} stages {r['stages']} gpt {r['gpt']} ctas {r['ctas']}: {r['ms_min']:.4f}ms {r['GBps']:.0f}GB/s {r['frac']:.3f}")
for r in rows:
    if r["what"] == "aux":
        print(f"aux {r['name']:70s} {r['ms_min']:.4f}ms {r['GBps']:7.0f}GB/s frac {r['frac']:.3f}")

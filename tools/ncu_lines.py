#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
stall samples, executed warp instructions, top stall reason.  usage: ncu_lines.py dump.csv [topN]"""
import csv
import sys
from collections import defaultdict

rows = csv.reader(open(sys.argv[1], newline=""))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, line, text, hdr = "", "", "", None
agg = defaultdict(lambda: defaultdict(float))
src = {}
kernel = 0
seen_files = set()
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        if fname in seen_files:  # the dump repeats every file once per profiled launch: keep the first
            kernel = 2
        elif kernel < 2:
            seen_files.add(fname)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if kernel > 1 or hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0]:
        line, text = r[0], r[1]
        src[(fname, line)] = text
    if len(r) > 8 and r[2].startswith("0x"):
        k = (fname, line)
        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        agg[k]["samples"] += num(r[hdr.index("Warp Stall Sampling (All Samples)")])
        agg[k]["inst"] += num(r[hdr.index("Instructions Executed")])
        for name in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_lg", "stall_barrier", "stall_mio",
                     "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_dispatch"):
            if name in hdr:
                agg[k][name] += num(r[hdr.index(name)])
import os
import re
by_fn = "--functions" in sys.argv
if by_fn:  # roll lines up to the enclosing __device__ function of the repo's sources
    regions = {}
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in ("mcts_kernels.cu", "bitboard.cuh", "philox.cuh", "env_kernels.cu"):
        path = os.path.join(here, "alphazero_othello_b200", "csrc", f)
        regions[f] = [(i, m.group(1)) for i, l in enumerate(open(path), 1)
                      for m in [re.match(r"\s*(?:template.*>\s*)?(?:static\s+)?__(?:device|global)__ .*?(\w+)\(", l)] if m]
    def region(k):
        f, ln = k
        name = f
        for start, n in regions.get(f, []):
            if ln.isdigit() and int(ln) >= start:
                name = f + ":" + n
        return name
    rolled = defaultdict(lambda: defaultdict(float))
    for k, v in agg.items():
        for n, x in v.items():
            rolled[(region(k), "")][n] += x
    agg = rolled
    src = {}
tot = sum(v["samples"] for v in agg.values())
toti = sum(v["inst"] for v in agg.values())
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    reasons = sorted(((n, x) for n, x in v.items() if n.startswith("stall_")), key=lambda t: -t[1])[:2]
    rs = " ".join(f"{n[6:]}={x:.0f}" for n, x in reasons)
    print(f"{k[0]}:{k[1]:>4} smp {100 * v['samples'] / tot:5.1f}% inst {100 * v['inst'] / toti:5.1f}%  {rs:32s} | {src.get(k, '').strip()[:90]}")

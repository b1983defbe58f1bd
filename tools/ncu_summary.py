#!/usr/bin/env python
"""Key metrics of every captured launch in an .ncu-rep -> profiles/<tag>.json (tracked).
usage: ncu_summary.py <tag> <report.ncu-rep> [kernel regex] [note]"""
import csv
import json
import re
import subprocess
import sys

tag, rep = sys.argv[1:3]
pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else ".")
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__cycles_elapsed.max"]
caps = []


def to_bytes(s):
    v, u = s.split()[:2]
    return float(v.replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "Tbyte": 1e12}[u]


def to_us(s):
    v, u = s.split()[:2]
    return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)


for r in rr[2:]:
    name = r[h.index("Kernel Name")]
    if not pat.search(name):
        continue
    d = {"kernel": name[:110]}
    for k in keys:
        if k in h:
            d[k] = f"{r[h.index(k)]} {units[h.index(k)]}".strip()
    if "dram__bytes_read.sum" in d:
        d["dram_bytes"] = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
        d["dram_gbs"] = d["dram_bytes"] / (to_us(d["gpu__time_duration.sum"]) * 1e-6) / 1e9
    caps.append(d)
out = {"tag": tag, "note": note, "command": "ncu --set full --clock-control none --import-source on (tools/r2_ncu.sh); per-launch figures are "
                                             "cold-cache and serialised", "launches": caps}
if caps and "dram_bytes" in caps[0]:
    out["dram_bytes_per_launch"] = sum(c["dram_bytes"] for c in caps) / len(caps)
json.dump(out, open(f"profiles/{tag}.json", "w"), indent=1)
for c in caps:
    print(c["kernel"][:60], c.get("gpu__time_duration.sum"), "dram MB %.2f" % (c.get("dram_bytes", 0) / 1e6), "GB/s %.0f" % c.get("dram_gbs", 0),
          "regs", c.get("launch__registers_per_thread"), "alu%", c.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
          "issue%", c.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps%", c.get("sm__warps_active.avg.pct_of_peak_sustained_active"))

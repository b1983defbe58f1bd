#!/usr/bin/env python
"""Turn gpurun_out/{launches_*.csv, prof_*.ncu-rep} into the tracked summaries under profiles/.
usage: summarize_profile.py <tag> <launches.csv> <report.ncu-rep> [kernel regex]"""
import collections
import csv
import json
import re
import subprocess
import sys

tag, launches, rep = sys.argv[1:4]
pat = re.compile(sys.argv[4] if len(sys.argv) > 4 else "k_mcts_step")
out = {"tag": tag}

rows = list(csv.reader(open(launches)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[r[ki]][0] += 1
    agg[r[ki]][1] += v
tot = sum(v[1] for v in agg.values())
out["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none (tools/gpu_profile.sh)",
                      "launches": sum(v[0] for v in agg.values()), "total_us": tot,
                      "kernels": [{"kernel": k[:100], "launches": v[0], "avg_us": v[1] / v[0], "share": v[1] / tot}
                                  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]]}
mine = [(k, v) for k, v in agg.items() if pat.search(k)]
out["kernel_share_of_step"] = sum(v[1] for _, v in mine) / tot

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_alu.sum", "sm__inst_executed.sum"]
caps = []
for r in rr[2:]:
    if not pat.search(r[h.index("Kernel Name")]):
        continue
    d = {"kernel": r[h.index("Kernel Name")][:100]}
    for k in keys:
        if k in h:
            d[k] = f"{r[h.index(k)]} {units[h.index(k)]}".strip()
    caps.append(d)
out["full_capture"] = {"command": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2", "launches": caps}
if caps:
    def mb(s):
        v, u = s.split()[:2]
        return float(v) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[u]
    out["dram_bytes_per_launch"] = sum(mb(c["dram__bytes_read.sum"]) + mb(c["dram__bytes_write.sum"]) for c in caps) / len(caps)
json.dump(out, open(f"profiles/{tag}.json", "w"), indent=1)
print(json.dumps({k: out[k] for k in ("tag", "kernel_share_of_step", "dram_bytes_per_launch") if k in out}))

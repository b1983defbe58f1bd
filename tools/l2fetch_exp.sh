#!/bin/bash
# Experiment: C4 pipeline under different cudaLimitMaxL2FetchGranularity settings.
for g in 32 64 128; do
  OTH_L2_FETCH=$g python bench.py --steps 2 --no-cpu-baseline --no-aux 2>/dev/null | tail -1 > gpurun_out/l2_$g.json
  python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
d = json.load(open(f"gpurun_out/l2_{g}.json")); r = d["roofline"]
print("L2 fetch", g, "value", round(d["value"]), "kernel avg/median us", round(r["launch_ms_avg"] * 1000), round(r["launch_ms_median"] * 1000), "frac", round(r["frac"], 3))
PY
done

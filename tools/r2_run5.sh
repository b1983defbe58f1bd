#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_production_path_gpu.py -q -x --timeout 600 -k "dedup" 2>&1 | tail -4
timeout 600 python bench.py --workload c4 --steps 8 --warmup 3 --dedup 1 --no-aux --no-cpu-baseline > gpurun_out/dd_c4_1.json 2> gpurun_out/dd_c4_1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/dd_c4_1.json").read().strip().splitlines()[-1])
    print("c4 dedup 1: value %.4g e2e %.4g per-step ms %s" % (d["value"], d["e2e"]["value"], [round(x) for x in d["per_step_ms"]]), d["engine"]["evaluation_dedup"], d["network_roofline"]["frac"])
except Exception as e:
    print("FAILED", e); print("\n".join(l for l in open("gpurun_out/dd_c4_1.err").read().splitlines() if not l.startswith("frame"))[-1500:])
PY

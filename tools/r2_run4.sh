#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_production_path_gpu.py -q -x --timeout 600 -k "dedup" 2>&1 | tail -15
for dd in 1 0; do
  timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --dedup $dd --no-aux --no-cpu-baseline > gpurun_out/dd_c4_$dd.json 2> gpurun_out/dd_c4_$dd.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dd_c4_$dd.json").read().strip().splitlines()[-1])
    print("c4 dedup $dd: value %.4g e2e %.4g per-step ms %s" % (d["value"], d["e2e"]["value"], [round(x) for x in d["per_step_ms"]]), d["engine"]["evaluation_dedup"]["iterations_by_bucket_since_start"], d["network_roofline"]["frac"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/dd_c4_$dd.err").read()[-1500:])
PY
done
timeout 600 python bench.py --workload c3 --steps 8 --warmup 3 --dedup 1 --no-aux --no-cpu-baseline > gpurun_out/dd_c3_1.json 2> gpurun_out/dd_c3_1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/dd_c3_1.json").read().strip().splitlines()[-1])
    print("c3 dedup 1: value %.4g e2e %.4g per-step ms %s" % (d["value"], d["e2e"]["value"], [round(x, 1) for x in d["per_step_ms"]]), d["engine"]["evaluation_dedup"]["iterations_by_bucket_since_start"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/dd_c3_1.err").read()[-1500:])
PY

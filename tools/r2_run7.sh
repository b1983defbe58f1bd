#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/r2_run6.sh
echo "== full gpu suite"
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 --durations=8 > gpurun_out/r02_gpu_suite.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/r02_gpu_suite.log
echo "== default bench"
timeout 1500 python bench.py > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/r02_bench_c4_1gpu.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c4_1gpu.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g per-step %s" % (d["value"], d["e2e"]["value"], [round(x) for x in d["per_step_ms"]]))
print("without dedup", d["aux"]["without_dedup"]["sims_per_s"], "c3", d["aux"]["other_architecture"]["sims_per_s"], d["aux"]["other_architecture_without_dedup"]["sims_per_s"])
r = d["roofline"]; print("share", r["kernel_share_of_iteration"], r["kernel_share_of_iteration_net_of_event_overhead"], "ovh", r["event_pair_overhead_ms"])
r = d["aux"]["other_architecture"]["roofline"]; print("c3 share", r["kernel_share_of_iteration"], r["kernel_share_of_iteration_net_of_event_overhead"])
PY

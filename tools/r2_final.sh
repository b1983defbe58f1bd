#!/bin/bash
# End-of-round pass: smoke, the whole -m gpu suite, the default bench and the reference arm with the driver's flags.
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 --durations=5 > gpurun_out/r02_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -9 gpurun_out/r02_gpu_suite.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/r02_bench_c4_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c4_1gpu.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g per-step %s" % (d["value"], d["e2e"]["value"], [round(x) for x in d["per_step_ms"]]))
a = d["aux"]
print("without dedup", a["without_dedup"]["sims_per_s"], "| c3", a["other_architecture"]["sims_per_s"], "| c2", a["c2_one_game"]["sims_per_s"], a["c2_one_self_play_dropin"]["sims_per_s"], "| search-only", a["search_only"]["sims_per_s"], "| tf32", a["network_precision"]["tf32"]["sims_per_s"])
r = d["roofline"]; print("share", r["kernel_share_of_iteration"], r["kernel_share_of_iteration_net_of_event_overhead"], "frac", r["frac"], "rnd", r["random_access"]["frac"], "step us", 1e3 * r["launch_ms_avg"], "move us", 1e3 * r["move_kernel"]["launch_ms_avg"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["port"]["value"], "clocks", d["clocks"])
print(d["engine"]["evaluation_dedup"]["iterations_by_bucket_since_start"], d["network_roofline"]["frac"])
PY
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref_arm.err; echo "ref rc=$?"; cut -c1-330 gpurun_out/r02_bench_reference_arm.json

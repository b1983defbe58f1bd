#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== default bench (c4, aux, cpu baseline)"
timeout 1500 python bench.py > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/r02_bench_c4_1gpu.err; echo "rc=$?"
tail -c 600 gpurun_out/r02_bench_c4_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c4_1gpu.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g share %.3f step %.1f us move %.1f us" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_share_of_iteration"], 1e3*d["roofline"]["launch_ms_avg"], 1e3*d["roofline"]["move_kernel"]["launch_ms_avg"]))
print("cpu_baseline", {k: (v if k != "sample" else v[:60]) for k, v in d["cpu_baseline"].items() if k != "port"}, "port", d["cpu_baseline"].get("port", {}).get("value"))
a = d["aux"]
print("search_only", {k: v for k, v in a["search_only"].items() if k != "workload"})
o = a["other_architecture"]; print("c3", o["sims_per_s"], "share", o["roofline"]["kernel_share_of_iteration"], "step", o["roofline"]["launch_ms_avg"], "move", o["roofline"]["move_kernel"]["launch_ms_avg"], "frac", o["roofline"]["frac"], "rnd", o["roofline"]["random_access"]["frac"])
o = a["c2_one_game"]; print("c2", o["sims_per_s"], "share", o["roofline"]["kernel_share_of_iteration"])
print("c2 dropin", a["c2_one_self_play_dropin"]["sims_per_s"], a["c2_one_self_play_dropin"]["seconds"])
print("precision", {k: v["sims_per_s"] for k, v in a["network_precision"].items()})
print("public api", a["public_api_whole_games"]["sims_per_s"])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref_arm.err; echo "rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
echo "== dedup probes"
timeout 600 python tools/dedup_probe.py --workload c4 --every 25 > gpurun_out/r02_dedup_probe_c4.json 2> gpurun_out/dedup_c4.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02_dedup_probe_c4.json'));print({k:d[k] for k in ('probes','leaves_probed','distinct','duplicate_fraction','by_game_phase')})"
timeout 300 python tools/dedup_probe.py --workload c3 --every 20 > gpurun_out/r02_dedup_probe_c3.json 2> gpurun_out/dedup_c3.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02_dedup_probe_c3.json'));print({k:d[k] for k in ('probes','leaves_probed','distinct','duplicate_fraction','by_game_phase')})"
echo "== OTH_DEBUG build: production-path tests with arena index assertions"
OTH_B200_DEBUG=1 timeout 1500 python -m pytest tests/test_production_path_gpu.py -q --timeout 900 > gpurun_out/r02_debug_build_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_debug_build_tests.log

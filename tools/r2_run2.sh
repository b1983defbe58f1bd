#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/debug_stuck.py 2>&1 | grep -c "finished=True"
SKIP_REST=1 bash tools/r2_tests.sh
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_native_host_gpu.py tests/test_arena_gpu.py -q -x --timeout 300 2>&1 | tail -5
one() {
  timeout 400 python bench.py --workload $1 --lanes $2 --move-launch $3 --steps 4 --warmup 3 --no-aux --no-cpu-baseline \
      > gpurun_out/sweep_$1_l$2_m$3.json 2> gpurun_out/sweep_$1_l$2_m$3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$1_l$2_m$3.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$1 lanes $2 move_launch $3: %.4g sims/s  e2e %.4g  step %.1f us avg / %.1f med  move %.1f avg / %.1f med  share %.3f  sm_mhz %s" % (
        d["value"], d["e2e"]["value"], 1e3 * r["launch_ms_avg"], 1e3 * r["launch_ms_median"], 1e3 * r["move_kernel"]["launch_ms_avg"],
        1e3 * r["move_kernel"]["launch_ms_median"], r["kernel_share_of_iteration"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1 lanes $2 move_launch $3: FAILED", e)
    print(open("gpurun_out/sweep_$1_l$2_m$3.err").read()[-800:])
PY
}
one c3 32 0; one c3 32 1; one c2 32 0; one c2 32 1
# launch lists (no graph) of the small-architecture configs: where does an iteration go?
for wl in c3 c2; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches_${wl}.csv \
    python bench.py --workload $wl --steps 1 --warmup 3 --iters-per-step 40 --no-aux --no-cpu-baseline --no-graph > gpurun_out/ncu_list_${wl}.log 2>&1
  echo "launch list $wl rc=$?"
done

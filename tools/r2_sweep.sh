#!/bin/bash
# Round-2 measurement sweep (run under gpurun, one B200): smoke, GPU test suite, then the lane-width / move-launch
# matrix of the small-architecture configs.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== production-path tests"; timeout 1500 python -m pytest tests/test_production_path_gpu.py -x -q 2>&1 | tail -15
echo "== rest of the gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_production_path_gpu.py 2>&1 | tail -8
one() {  # workload lanes move_launch tag
  timeout 400 python bench.py --workload $1 --lanes $2 --move-launch $3 --steps 4 --warmup 3 --no-aux --no-cpu-baseline \
      > gpurun_out/sweep_$1_l$2_m$3.json 2> gpurun_out/sweep_$1_l$2_m$3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$1_l$2_m$3.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$1 lanes $2 move_launch $3: %.3g sims/s  e2e %.3g  step %.1f us avg / %.1f med  move %.1f avg / %.1f med  share %.3f  sm_mhz %s" % (
        d["value"], d["e2e"]["value"], 1e3 * r["launch_ms_avg"], 1e3 * r["launch_ms_median"], 1e3 * r["move_kernel"]["launch_ms_avg"],
        1e3 * r["move_kernel"]["launch_ms_median"], r["kernel_share_of_iteration"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1 lanes $2 move_launch $3: FAILED", e)
    print(open("gpurun_out/sweep_$1_l$2_m$3.err").read()[-800:])
PY
}
for l in 8 16 32; do for m in 0 1; do one c3 $l $m; done; done
for l in 8 32; do for m in 0 1; do one c2 $l $m; done; done
one c4 8 0; one c4 8 1

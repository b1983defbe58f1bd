#!/bin/bash
# ncu launch list only (no full capture): tools/launch_list.sh <workload> [skip] [count]
WL=${1:-c3}
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --iters-per-step 40 --no-aux --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain_${WL}.log 2>&1 || { echo plain run failed; tail -3 gpurun_out/plain_${WL}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s ${2:-3000} -c ${3:-600} --csv --log-file gpurun_out/launches_${WL}.csv $CMD > gpurun_out/ncu_list_${WL}.log 2>&1
echo "rc=$?"

#!/bin/bash
# Profiling recipe (run under gpurun on one B200): plain run, then the ncu launch list and
# one --set full capture of the MCTS kernel for the same command. Outputs land in gpurun_out/.
#   tools/gpu_profile.sh [lanes] [workload]
set -u
LANES=${1:-32}
WL=${2:-c4}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --iters-per-step 40 --lanes $LANES --no-aux --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain_${WL}_l${LANES}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${WL}_l${LANES}.log; exit 1; }
tail -1 gpurun_out/plain_${WL}_l${LANES}.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 600 --csv \
    --log-file gpurun_out/launches_${WL}_l${LANES}.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_mcts_step -s 130 -c 2 \
    -o gpurun_out/prof_mcts_${WL}_l${LANES} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log

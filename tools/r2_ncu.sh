#!/bin/bash
# ncu --set full captures asked for by the round-1 review: stepwise env kernels, the move kernel in the launch where every
# slot re-roots, the C3 step kernel.  Each program has exited 0 without ncu before (tools/r2_run*.sh).
set -u
mkdir -p gpurun_out
echo "== env API"
ncu --set full --clock-control none --import-source on -k regex:'k_legal_moves|k_step' -s 24 -c 4 -o gpurun_out/prof_env_api -f python tools/env_profile.py > gpurun_out/ncu_env.log 2>&1; echo "rc=$?"
echo "== C3 step kernel (lanes 32)"
CMD3="python bench.py --workload c3 --steps 1 --warmup 3 --iters-per-step 40 --no-aux --no-cpu-baseline --no-graph"
ncu --set full --clock-control none --import-source on -k regex:k_mcts_step_fused -s 130 -c 2 -o gpurun_out/prof_mcts_c3_l32 -f $CMD3 > gpurun_out/ncu_c3.log 2>&1; echo "rc=$?"
echo "== C4 move kernel: launches around the one where all 16384 slots re-root"
CMD4="python bench.py --workload c4 --steps 1 --warmup 3 --iters-per-step 140 --no-aux --no-cpu-baseline --no-graph --move-launch ${ML:-1}"
ncu --set full --clock-control none --import-source on -k regex:k_mcts_move -s 398 -c 8 -o gpurun_out/prof_mcts_move_c4 -f $CMD4 > gpurun_out/ncu_move.log 2>&1; echo "rc=$?"
echo "== C4 step kernel (lanes 8), 2 launches mid-search"
ncu --set full --clock-control none --import-source on -k regex:k_mcts_step_fused -s 130 -c 2 -o gpurun_out/prof_mcts_c4_l8 -f $CMD4 > gpurun_out/ncu_c4.log 2>&1; echo "rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 600 --csv --log-file gpurun_out/launches_c4.csv $CMD4 > gpurun_out/ncu_list_c4.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# After the move kernel's register budget changed: parity subset, then the --set full capture of the all-slots re-root launch.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_production_path_gpu.py tests/test_mcts_gpu.py -q --timeout 600 -k "split or full_size or big-96 or small-64 or self_play or deep_tree or randomised" 2>&1 | tail -3
CMD4="python bench.py --workload c4 --steps 1 --warmup 3 --iters-per-step 140 --no-aux --no-cpu-baseline --no-graph --dedup 0"
ncu --set full --clock-control none --import-source on -k regex:k_mcts_move -s 398 -c 8 -o gpurun_out/prof_mcts_move_c4 -f $CMD4 > gpurun_out/ncu_move.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# Extended differential fuzz (150 draws = 3 600 games, kernel variant randomised) and the -DOTH_DEBUG build over the
# production-path suite.
set -u
mkdir -p gpurun_out
OTH_FUZZ_SEEDS=150 timeout 1500 python -m pytest tests/test_mcts_gpu.py -q -k randomised --timeout 300 2>&1 | tail -3 | tee gpurun_out/r02_fuzz150.txt
OTH_B200_DEBUG=1 timeout 1500 python -m pytest tests/test_production_path_gpu.py tests/test_mcts_gpu.py -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/r02_debug_build_tests.txt

#!/bin/bash
# GPU test pass with logs kept (run under gpurun): smoke, the production-path file, then the rest of the -m gpu suite
# with a per-test timeout so that a hang names its test.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== production-path tests"
timeout 2400 python -m pytest tests/test_production_path_gpu.py -q -s -v --durations=30 --timeout 900 ${PROD_ARGS:-} > gpurun_out/prod_tests.log 2>&1; echo "rc=$?"
grep -E "PASSED|FAILED|ERROR|row invariance|passed|failed|Timeout" gpurun_out/prod_tests.log | cut -c1-220 | tail -40
if [ -z "${SKIP_REST:-}" ]; then
echo "== rest of the gpu suite"
timeout 2400 python -m pytest tests -q -v -m gpu --deselect tests/test_production_path_gpu.py --durations=25 --timeout 600 > gpurun_out/rest_tests.log 2>&1; echo "rc=$?"
grep -E "FAILED|ERROR|passed|failed|Timeout" gpurun_out/rest_tests.log | cut -c1-220 | tail -30
grep -A30 "slowest" gpurun_out/rest_tests.log | head -32
fi

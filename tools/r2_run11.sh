#!/bin/bash
# Measured-time bucket policy: parity tests, whole C4 games (one and two per slot).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_production_path_gpu.py -q --timeout 600 -k "dedup" 2>&1 | tail -3
timeout 400 python tools/full_games.py c4 1 1 > gpurun_out/r02_full_games_c4_dedup1.json 2> gpurun_out/fg_1.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02_full_games_c4_dedup1.json'));print({k:d[k] for k in ('evaluation_dedup','seconds','sims_per_s','iterations','iterations_by_bucket')}); print(d['interval_sims_per_s'])"
timeout 600 python tools/full_games.py c4 2 1 > gpurun_out/r02_full_games_c4_two_per_slot_dedup1.json 2> gpurun_out/fg_2.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02_full_games_c4_two_per_slot_dedup1.json'));print({k:d[k] for k in ('evaluation_dedup','seconds','sims_per_s','iterations','iterations_by_bucket')}); print(d['interval_sims_per_s'])"

#!/bin/bash
set -u
mkdir -p gpurun_out
for dd in 1 0; do
  timeout 400 python tools/full_games.py c4 1 $dd > gpurun_out/r02_full_games_c4_dedup$dd.json 2> gpurun_out/fg_$dd.err; echo "rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02_full_games_c4_dedup$dd.json'));print({k:d[k] for k in ('evaluation_dedup','seconds','sims','sims_per_s','positions_per_s','iterations','iterations_by_bucket','arena_high_water','max_depth')}); print(d['interval_sims_per_s'])"
done
# steady state: two games per slot (the second game starts desynchronised)
timeout 600 python tools/full_games.py c4 2 1 > gpurun_out/r02_full_games_c4_two_per_slot_dedup1.json 2> gpurun_out/fg_2.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02_full_games_c4_two_per_slot_dedup1.json'));print({k:d[k] for k in ('evaluation_dedup','seconds','sims','sims_per_s','iterations_by_bucket')}); print(d['interval_sims_per_s'])"

#!/bin/bash
# Does the step kernel's in-pipeline time depend on the arena footprint (TLB reach / DRAM row locality)?
for nc in 20224 12288 8192; do
  timeout 200 python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline --node-cap $nc 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('node_cap $nc', d['config']['l2'][:28], 'value %.0f | step kernel ms avg %.4f median %.4f p90 %.4f' % (d['value'], r['launch_ms_avg'], r['launch_ms_median'], r['launch_ms_p90']))"
done

#!/bin/bash
# A/B of the dead-activation L2 discard (Models.FoldedNet.l2_discard): off | tail | all
for m in off tail all; do
  OTH_L2_DISCARD=$m timeout 200 python bench.py --steps 3 --warmup 3 --no-aux --no-cpu-baseline ${1:+--workload $1} 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$m', 'value %.0f e2e %.0f ms/step %.1f | step kernel ms avg %.4f median %.4f p90 %.4f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['launch_ms_avg'], r['launch_ms_median'], r['launch_ms_p90'], r['frac']))"
done

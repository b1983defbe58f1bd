#!/bin/bash
# Does an idle gap between the network and the step kernel change the step kernel's in-pipeline time?
for us in 0 20 200 2000; do
  OTH_BENCH_GAP_US=$us timeout 200 python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('gap ${us} us: step kernel ms avg %.4f median %.4f p90 %.4f | move median %.4f' % (r['launch_ms_avg'], r['launch_ms_median'], r['launch_ms_p90'], r['move_kernel']['launch_ms_median']))"
done

#!/bin/bash
# bench.py variants still run end to end: other workloads as the headline (aux on), eager mode, de-duplication off.
set -u
mkdir -p gpurun_out
chk() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "OK value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), "aux keys", sorted(d.get("aux") or {})[:20], "cpu", (d.get("cpu_baseline") or {}).get("kind"))
except Exception as e:
    print(sys.argv[1], "FAILED", repr(e)); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
timeout 600 python bench.py --workload c3 --steps 4 --warmup 3 > gpurun_out/v_c3.json 2> gpurun_out/v_c3.err; chk gpurun_out/v_c3.json
timeout 600 python bench.py --workload c2 --steps 4 --warmup 3 > gpurun_out/v_c2.json 2> gpurun_out/v_c2.err; chk gpurun_out/v_c2.json
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-graph --no-aux --no-cpu-baseline > gpurun_out/v_c4_eager.json 2> gpurun_out/v_c4_eager.err; chk gpurun_out/v_c4_eager.json
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --dedup 0 --move-launch 0 --lanes 16 --no-aux --no-cpu-baseline > gpurun_out/v_c4_l16.json 2> gpurun_out/v_c4_l16.err; chk gpurun_out/v_c4_l16.json

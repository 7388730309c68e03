#!/bin/bash
# gpurun job: the primary-wave refill threshold on config 3
mkdir -p gpurun_out
for fp in 8 16 24 32; do
  RT_B200_FETCH_PRIMARY=$fp timeout 300 python bench.py --workload config3 --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/sweep.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep.json")); k=d["config"]["kernel_ms_per_step"]
print("fetch_primary $fp  %.2f ms  trace %.2f logic %.2f"%(d["ms_per_step"],k["k_trace_wave"],k["k_logic"]))
PY
done

#!/bin/bash
mkdir -p gpurun_out
for w in config2 config4; do for fp in 16 24 28; do
  RT_B200_FETCH_PRIMARY=$fp timeout 400 python bench.py --workload $w --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/sweep.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep.json")); k=d["config"]["kernel_ms_per_step"]
print("$w fetch_primary $fp  %.2f ms  trace %.2f logic %.2f"%(d["ms_per_step"],k["k_trace_wave"],k["k_logic"]))
PY
done; done

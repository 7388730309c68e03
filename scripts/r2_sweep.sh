#!/bin/bash
# gpurun job: sweep of the trace kernel's scheduling knobs on config 3 (kernel ms per step)
mkdir -p gpurun_out
for lw in 8 12 16 20; do for fm in 8 12 16 24; do
  RT_B200_LEAF_WAIT=$lw RT_B200_FETCH_MIN=$fm RT_B200_FETCH_SHADOW=$fm timeout 300 python bench.py --workload ${1:-config3} --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/sweep.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep.json")); k=d["config"]["kernel_ms_per_step"]
print("leaf_wait $lw fetch_min $fm  %.2f ms  trace %.2f logic %.2f"%(d["ms_per_step"],k["k_trace_wave"],k["k_logic"]))
PY
done; done

#!/bin/bash
# gpurun job: pool size (path slots in flight) vs frame time
mkdir -p gpurun_out
for pool in 33554432 67108864 134217728 268435456; do
  for w in ${1:-config3 config4}; do
    RT_B200_POOL=$pool timeout 600 python bench.py --workload $w --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/pool.json 2> gpurun_out/pool.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pool.json")); k=d["config"]["kernel_ms_per_step"]
    print("pool $pool $w  %.2f ms  %.0f Mrays/s  trace %.2f logic %.2f waves %d"%(d["ms_per_step"],d["value"],k["k_trace_wave"],k["k_logic"],d["config"]["waves_per_step"]))
except Exception as e:
    print("pool $pool $w failed", open("gpurun_out/pool.err").read()[-400:])
PY
  done
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv

#!/bin/bash
# gpurun job: bench lines (kernel times) per child-bound variant: scripts/r2_bounds.sh "qbox box ..." "config2 config3 config4" [pytest]
mkdir -p gpurun_out
if [ -n "$3" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "child_bound or seeded or trace" 2>&1 | tail -3; fi
for v in $1; do
  for w in $2; do
    RT_B200_BOUNDS=$v timeout 600 python bench.py --workload $w --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_bounds_${v}_${w}.json 2> gpurun_out/r2_bounds_${v}_${w}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bounds_${v}_${w}.json"))
    k=d["config"]["kernel_ms_per_step"]
    print("%-10s %-8s %8.0f Mrays/s %9.2f ms  trace %8.2f  logic %8.2f"%("$v","$w",d["value"],d["ms_per_step"],k["k_trace_wave"],k["k_logic"]))
except Exception as e:
    print("$v $w failed", e); print(open("gpurun_out/r2_bounds_${v}_${w}.err").read()[-1500:])
PY
  done
done

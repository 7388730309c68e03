#!/bin/bash
# gpurun job: bench lines of config 2 / 3 / 4 (kernel times) for library variants: scripts/r2_variants.sh "<name> <name> ..." [workloads]
mkdir -p gpurun_out
WL=${2:-"config2 config3 config4"}
for v in $1; do
  lib=par_raytracer_b200/librt_b200_$v.so; [ "$v" = default ] && lib=par_raytracer_b200/librt_b200.so
  for w in $WL; do
    RT_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload $w --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_var_${v}_${w}.json 2> gpurun_out/r2_var_${v}_${w}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_var_${v}_${w}.json"))
    k=d["config"]["kernel_ms_per_step"]
    print("%-10s %-8s %8.0f Mrays/s %9.2f ms  trace %8.2f  logic %8.2f"%("$v","$w",d["value"],d["ms_per_step"],k["k_trace_wave"],k["k_logic"]))
except Exception as e:
    print("$v $w failed", e); print(open("gpurun_out/r2_var_${v}_${w}.err").read()[-1500:])
PY
  done
done

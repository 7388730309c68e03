#!/bin/bash
# gpurun job: 4-wide traversal -- parity tests under RT_B200_BOUNDS=qbox4, then config 2/3/4 trace times for qbox vs qbox4
mkdir -p gpurun_out
RT_B200_BOUNDS=qbox4 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_combine.py -m gpu -x -q > gpurun_out/r2_pytest_qbox4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_qbox4.log
tail -5 gpurun_out/r2_pytest_qbox4.log
for b in qbox qbox4; do
  for w in config3 config2 config4; do
    RT_B200_BOUNDS=$b timeout 600 python bench.py --workload $w --also none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_${b}_${w}.json 2> gpurun_out/r2_${b}_${w}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_${b}_${w}.json"))
    print("$b $w", round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms", d["config"]["kernel_ms_per_step"], "nodes", d["config"]["hierarchy_nodes"])
except Exception as e:
    print("$b $w failed", e); print(open("gpurun_out/r2_${b}_${w}.err").read()[-2000:])
PY
  done
done

#!/bin/bash
# gpurun job: the whole GPU test-suite (verbose enough to see the adaptive pixel-difference counts) + config 1 bench line + per-wave log of config 3
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s --durations=8 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
grep -E "pixels|passed|failed|rc=|Error|error" gpurun_out/r2_pytest_gpu.log | tail -20; tail -12 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --workload config1 --also none --steps 5 --warmup 3 > gpurun_out/r2_bench_config1.json 2> gpurun_out/r2_bench_config1.err; echo "bench config1 rc=$?"; tail -c 1500 gpurun_out/r2_bench_config1.json; tail -3 gpurun_out/r2_bench_config1.err
RT_B200_WAVE_LOG=1 timeout 600 python bench.py --workload config3 --also none --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2_wavelog_config3.json 2> gpurun_out/r2_wavelog_config3.err; echo "wavelog rc=$?"

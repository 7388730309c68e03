#!/bin/bash
# gpurun job: parity tests of the default library (RT_POOL_AOS=1), then A/B against the structure-of-arrays pool
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_aos.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_aos.log; tail -3 gpurun_out/r2_pytest_gpu_aos.log
bash scripts/r2_variants.sh "soa default" "config3 config2 config4"

#!/bin/bash
# gpurun job: parity tests of the default library, then A/B of the two k_logic changes (staged queue entries, frames popped with their last child)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_ab.log; tail -3 gpurun_out/r2_pytest_gpu_ab.log
bash scripts/r2_variants.sh "base default stage pop" "config3 config2 config4"

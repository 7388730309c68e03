#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_combine.py -m gpu -x -q 2>&1 | tail -3
bash scripts/r2_variants.sh "default wchunk128 wchunk512" "config2 config3 config4"

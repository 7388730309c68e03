#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
C="32:16,16:16,12:16,8:16,4:16,2:16,12:8,8:8,4:8,8:24,8:12"
for s in c2 707 2236; do
  python scripts/sweep2.py $s $C
  RT_B200_LIB=scripts/variants/librt_prefetch.so python scripts/sweep2.py $s 32:16,8:16
done
RT_B200_WAVE_LOG=1 python scripts/profile_step.py 64 1 2>&1 | tail -40

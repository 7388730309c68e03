#!/bin/bash
RT_B200_POOL_K=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
RT_B200_POOL_K=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
C="4:16,8:16,12:16,16:16,24:16,8:8,8:24,12:24"
for s in c2 707 2236; do
  for k in 2 3; do
    echo "K=$k"; RT_B200_POOL_K=$k python scripts/sweep2.py $s $C
  done
done

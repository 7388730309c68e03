#!/bin/bash
export RT_B200_POOL_K=2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
C="8:16,16:16,24:16,32:16,48:16,24:24"
for s in c2 707 2236; do
  python scripts/sweep2.py $s $C
  RT_B200_LIB=scripts/variants/librt_minb6.so python scripts/sweep2.py $s $C
done
RT_B200_PLOC_AREA=1 RT_B200_POOL_K=1 python scripts/sweep2.py 2236 12:16

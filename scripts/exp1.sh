#!/bin/bash
# experiment: box vs sphere+slab child bounds, PLOC cost metric
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for b in box sphere; do
  echo "== bounds=$b config2 64spp"; RT_B200_BOUNDS=$b python scripts/profile_step.py 64 2 | tail -1
  echo "== bounds=$b 10M 16spp"; RT_B200_BOUNDS=$b RT_B200_WAVE_LOG=1 python scripts/profile_big.py 2236 16 2>&1 | tail -12
  echo "== bounds=$b 1M 16spp"; RT_B200_BOUNDS=$b python scripts/profile_big.py 707 16 2>&1 | tail -1
done
echo "== area cost"
RT_B200_PLOC_AREA=1 python scripts/profile_step.py 64 2 | tail -1
RT_B200_PLOC_AREA=1 python scripts/profile_big.py 2236 16 2>&1 | tail -1
RT_B200_PLOC_AREA=1 python scripts/profile_big.py 707 16 2>&1 | tail -1

#!/bin/bash
for lib in default scripts/variants/librt_tmb7.so scripts/variants/librt_unroll2.so scripts/variants/librt_unroll2_tmb7.so; do
  for s in c2 2236; do
    if [ $lib = default ]; then python scripts/sweep2.py $s 12:16; else RT_B200_LIB=$lib python scripts/sweep2.py $s 12:16,8:16,16:16; fi
  done
done

#!/bin/bash
for lib in scripts/variants/librt_unroll3.so scripts/variants/librt_unroll4.so scripts/variants/librt_unroll4_tmb7.so; do
  for s in c2 707 2236; do RT_B200_LIB=$lib python scripts/sweep2.py $s 12:16,16:16; done
done

#!/bin/bash
for s in c2 707 2236; do for mb in 10 12; do RT_B200_LIB=scripts/variants/librt_minb$mb.so python scripts/sweep2.py $s 12:16; done; done

#!/bin/bash
for pe in 1 0; do for s in 707 2236; do echo "persist=$pe"; RT_B200_L2_PERSIST=$pe RT_B200_WAVE_LOG=1 python scripts/sweep2.py $s 12:16 2>&1 | grep -v "^\[wave" | tail -2; done; done

#!/bin/bash
for p in 4194304 8388608 16777216 33554432 67108864; do echo "pool=$p"; RT_B200_POOL=$p python scripts/profile_step.py 64 2 | tail -1; done

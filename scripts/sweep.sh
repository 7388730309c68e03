#!/bin/bash
for cfg in "16 16 16" "16 32 16" "16 16 32" "16 32 32" "16 24 24" "12 32 24" "20 32 32" "16 8 8"; do set -- $cfg; echo "bounce=$1 primary=$2 shadow=$3"; RT_B200_FETCH_MIN=$1 RT_B200_FETCH_PRIMARY=$2 RT_B200_FETCH_SHADOW=$3 python scripts/profile_step.py 64 2 | tail -1; done

#!/bin/bash
export RT_B200_POOL_K=2 RT_B200_LEAF_WAIT=12
ncu --set full --import-source on --clock-control none -k regex:k_trace -c 3 -o gpurun_out/r1_pool2_big python scripts/profile_big.py 2236 4 > gpurun_out/ncu_pool2_big.log 2>&1
export RT_B200_POOL_K=1 RT_B200_LEAF_WAIT=12
ncu --set full --import-source on --clock-control none -k regex:k_trace -c 3 -o gpurun_out/r1_box12_big python scripts/profile_big.py 2236 4 > gpurun_out/ncu_box12_big.log 2>&1

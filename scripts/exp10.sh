#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:k_trace -c 3 -o gpurun_out/r1_qbox_big python scripts/profile_big.py 2236 4 > gpurun_out/ncu_qbox_big.log 2>&1

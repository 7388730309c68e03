#!/bin/bash
# ncu --set full captures for profiles/: the three densest waves of a config-2 frame (trace + logic), and waves 0-2 of the 10 M-triangle scene
ncu --set full --import-source on --clock-control none -k regex:"k_trace_wave|k_logic" --launch-skip 22 -c 6 -o gpurun_out/r1_final_dense python scripts/profile_step.py 64 1 > gpurun_out/ncu_final_dense.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_wave -c 3 -o gpurun_out/r1_final_big python scripts/profile_big.py 2236 4 > gpurun_out/ncu_final_big.log 2>&1

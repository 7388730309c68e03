#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:k_logic --launch-skip 11 -c 3 -o gpurun_out/r1_logic_dense python scripts/profile_step.py 64 1 > gpurun_out/ncu_logic_dense.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_wave --launch-skip 11 -c 3 -o gpurun_out/r1_trace_dense_q python scripts/profile_step.py 64 1 > gpurun_out/ncu_trace_dense_q.log 2>&1

#!/bin/bash
# gpurun job: ncu --set full of waves 0-2 (trace + logic) of a config-3 frame at 16 spp, final round-2 kernels
mkdir -p gpurun_out
python scripts/profile_c3.py 16 1 > gpurun_out/r2_c3_plain_final.log 2>&1 && \
timeout 130 ncu --set full --import-source on --clock-control none -k regex:"k_trace_wave|k_logic" -c 6 -o gpurun_out/r2_c3_final2 -f python scripts/profile_c3.py 16 1 > gpurun_out/r2_c3_ncu_final.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r2_c3_plain_final.log; tail -2 gpurun_out/r2_c3_ncu_final.log

#!/bin/bash
# gpurun job: ncu --set full of waves 0-2 (trace + logic) of the second config-3 frame at 16 spp
mkdir -p gpurun_out
python scripts/profile_c3.py 16 2 > gpurun_out/r2_c3_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"k_trace_wave|k_logic" --launch-skip ${1:-24} -c 6 -o gpurun_out/r2_c3_dense -f python scripts/profile_c3.py 16 2 > gpurun_out/r2_c3_ncu.log 2>&1
cat gpurun_out/r2_c3_plain.log; tail -5 gpurun_out/r2_c3_ncu.log

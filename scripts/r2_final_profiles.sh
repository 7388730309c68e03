#!/bin/bash
# gpurun job: the ncu evidence of round 2 (final kernels): launch list of the bench command, DRAM traffic of every launch of one config-3 frame,
# --set full capture of waves 0-2 (trace + logic) of a config-3 frame at 16 spp. Each command first runs without ncu.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --also none --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --also none --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts/profile_c3.py 128 1 > gpurun_out/r2_plain_c3_128.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:"k_trace_wave|k_logic|k_resolve|k_finalize" --csv --log-file gpurun_out/r2_traffic_c3.csv python scripts/profile_c3.py 128 1 > gpurun_out/r2_ncu_traffic.log 2>&1
echo "traffic rc=$?"
python scripts/profile_c3.py 16 2 > gpurun_out/r2_c3_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"k_trace_wave|k_logic" --launch-skip 26 -c 6 -o gpurun_out/r2_c3_final -f python scripts/profile_c3.py 16 2 > gpurun_out/r2_c3_ncu.log 2>&1
echo "full rc=$?"; cat gpurun_out/r2_c3_plain.log gpurun_out/r2_plain_c3_128.log; tail -2 gpurun_out/r2_c3_ncu.log

#!/bin/bash
# gpurun job: ncu evidence for the final round-2 kernels: launch list of the bench command, DRAM / L2 traffic of every launch of one config-3 frame.
# Each command first runs without ncu.
mkdir -p gpurun_out
python scripts/profile_c3.py 128 1 > gpurun_out/r2_plain_c3_128.log 2>&1 && \
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:"k_trace_wave|k_logic|k_resolve|k_finalize" --csv --log-file gpurun_out/r2_traffic_c3.csv python scripts/profile_c3.py 128 1 > gpurun_out/r2_ncu_traffic.log 2>&1
echo "traffic rc=$?"; cat gpurun_out/r2_plain_c3_128.log
python bench.py --steps 2 --warmup 1 --also none --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --also none --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"

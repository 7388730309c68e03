#!/bin/bash
# one-GPU validation pass: GPU test-suite, the bench line of every single-GPU workload, launch list + traffic for profiles/
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
python bench.py --workload config3 --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err
python bench.py --workload config4 --steps 2 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 300 gpurun_out/bench_c4.err
for s in c2 707 2236; do python scripts/sweep2.py $s 12:16,8:16,16:16; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v4.csv python scripts/profile_step.py 64 1 > gpurun_out/ncu_l4.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_traffic_v4.csv python scripts/profile_step.py 64 1 > gpurun_out/ncu_t4.log 2>&1

#!/bin/bash
export RT_B200_LEAF_WAIT=12
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:k_trace_wave --csv --log-file gpurun_out/sort_exp.csv python scripts/exp_sort.py 2236 > gpurun_out/sort_exp.log 2>&1

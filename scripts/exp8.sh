#!/bin/bash
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8
for lib in default scripts/variants/librt_ploc32.so scripts/variants/librt_ploc64.so; do
  for s in c2 707 2236; do
    if [ $lib = default ]; then python scripts/sweep2.py $s 12:16; else RT_B200_LIB=$lib python scripts/sweep2.py $s 12:16; fi
  done
done
python bench.py > gpurun_out/bench_box.json 2> gpurun_out/bench_box.err; tail -c 3000 gpurun_out/bench_box.json

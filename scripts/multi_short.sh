#!/bin/bash
N=$(nvidia-smi -L | wc -l)
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N "$@" 2> gpurun_out/bench_n${N}_$name.err | tail -1 > gpurun_out/bench_n${N}_$name.json; }
run samples --steps 5 --warmup 3 --partition samples
run config4_tiles --steps 2 --warmup 3 --partition tiles --workload config4
run tiles --steps 5 --warmup 3 --partition tiles

#!/bin/bash
# N-GPU pass (N = number of visible GPUs): NCCL tests + the scaling bench lines
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N "$@" 2> gpurun_out/bench_n${N}_$name.err | tail -1 > gpurun_out/bench_n${N}_$name.json; tail -c 200 gpurun_out/bench_n${N}_$name.err | tail -2; }
run samples --steps 5 --warmup 3 --partition samples
run tiles --steps 5 --warmup 3 --partition tiles
run config4_tiles --steps 2 --warmup 3 --partition tiles --workload config4
run config5_samples --steps 2 --warmup 3 --partition samples --workload config5

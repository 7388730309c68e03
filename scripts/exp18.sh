#!/bin/bash
# 2 GPUs: NCCL tests + scaling bench lines
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for part in samples tiles; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --partition $part 2> gpurun_out/bench_n2_$part.err | tail -1 > gpurun_out/bench_n2_$part.json
  tail -c 300 gpurun_out/bench_n2_$part.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 3 --partition tiles --workload config4 2> gpurun_out/bench_n2_c4.err | tail -1 > gpurun_out/bench_n2_c4.json
tail -c 300 gpurun_out/bench_n2_c4.err

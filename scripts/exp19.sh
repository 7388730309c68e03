#!/bin/bash
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 --partition samples 2> gpurun_out/bench_n8_samples.err | tail -1 > gpurun_out/bench_n8_samples.json
tail -c 200 gpurun_out/bench_n8_samples.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 2 --warmup 3 --partition tiles --workload config4 2> gpurun_out/bench_n8_c4.err | tail -1 > gpurun_out/bench_n8_c4.json
tail -c 200 gpurun_out/bench_n8_c4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 8 --steps 5 --warmup 3 --partition tiles 2> gpurun_out/bench_n8_tiles.err | tail -1 > gpurun_out/bench_n8_tiles.json

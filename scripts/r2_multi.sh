#!/bin/bash
# gpurun --gpus N job: multi-GPU tests (NCCL ranks, single-process peer memory / NCCL fallback) + bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
if [ "$2" != "nopytest" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_combine.py -m gpu -x -q > gpurun_out/r2_pytest_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi_n$N.log
tail -15 gpurun_out/r2_pytest_multi_n$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err

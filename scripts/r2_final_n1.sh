#!/bin/bash
# gpurun job: what the driver runs at round end on one GPU -- GPU tests, smoke, reference arm (short), bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err; echo "ref rc=$?"; tail -c 700 gpurun_out/r2_bench_ref_n1.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err

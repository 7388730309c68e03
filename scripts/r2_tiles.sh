#!/bin/bash
# gpurun --gpus N job: tile edge of the strong-scaling partition
N=${1:-8}
mkdir -p gpurun_out
for t in 8 16 32 64; do
  RT_BENCH_TILE=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 10 --warmup 3 --also none > gpurun_out/tiles_$t.json 2>/dev/null
  python - <<PY
import json
for l in open("gpurun_out/tiles_$t.json"):
    if l.startswith("{"):
        d=json.loads(l); print("tile $t: %.2f ms/frame %.0f Mrays/s parity %s %s"%(d["ms_per_step"],d["value"],d["combine_parity"],{k:round(v,2) for k,v in d["per_step_ms"].items() if k!="note"}))
PY
done

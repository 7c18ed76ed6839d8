#!/bin/bash
mkdir -p gpurun_out
run() {  # $1 tag, env in front
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_$1.json 2> gpurun_out/r2_bench_n2_$1.err
  echo "rc=$? $1"; tail -c 400 gpurun_out/r2_bench_n2_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n2_$1.json").read().strip().splitlines()[-1])
    print("$1", d["value"], d["ms_per_step"], d.get("ranks_in_sync"))
except Exception as e:
    print("$1 no line", e)
PY
}
run seg
RSR_GRAPH_NCCL=1 run graphnccl

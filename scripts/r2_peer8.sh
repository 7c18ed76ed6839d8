#!/bin/bash
# 8-GPU lines with the peer-memory all-reduce (the NCCL lines of the same tree: scripts/r2_bench_n8.sh)
mkdir -p gpurun_out
N=${N:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/gpu_peer_check.py > gpurun_out/r2_peer_check_n$N.txt 2> gpurun_out/r2_peer_check_n$N.err
echo "peer_check rc=$?"; tail -n 3 gpurun_out/r2_peer_check_n$N.txt; tail -n 5 gpurun_out/r2_peer_check_n$N.err
for spec in "cfg2 f16" "cfg5 f16" "cfg4 f16"; do
  set -- $spec
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $1 --dtype $2 --steps 20 --warmup 5 > gpurun_out/r2_bench_$1_$2_n${N}_peer1.json 2> gpurun_out/r2_bench_$1_$2_n${N}_peer1.err
  echo "== $1 $2 N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_$1_$2_n${N}_peer1.json").read().strip().splitlines()[-1])
    print("$1 $2", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "sync", d.get("ranks_in_sync"), d.get("allreduce"), d.get("allreduce_error"))
except Exception as e:
    print("$1 $2 no line", e)
PY
done

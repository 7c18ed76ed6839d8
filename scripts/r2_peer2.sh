#!/bin/bash
# peer-memory all-reduce on N GPUs: protocol tests, check against NCCL + timing, bench lines with and without it
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_peer_allreduce_gpu.py -q -m gpu -x > gpurun_out/r2_peer_tests.log 2>&1
tail -n 15 gpurun_out/r2_peer_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/gpu_peer_check.py > gpurun_out/r2_peer_check_n$N.txt 2> gpurun_out/r2_peer_check_n$N.err
echo "peer_check rc=$?"; tail -n 3 gpurun_out/r2_peer_check_n$N.txt; tail -n 5 gpurun_out/r2_peer_check_n$N.err
for mode in 1 0; do
  RSR_PEER_ALLREDUCE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$mode bench.py --gpus $N --config cfg2 --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2_n${N}_peer$mode.json 2> gpurun_out/r2_bench_cfg2_n${N}_peer$mode.err
  echo "== peer=$mode N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_cfg2_n${N}_peer$mode.json").read().strip().splitlines()[-1])
    print("value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "sync", d.get("ranks_in_sync"), d.get("allreduce"))
except Exception as e:
    print("no line", e)
PY
  tail -n 3 gpurun_out/r2_bench_cfg2_n${N}_peer$mode.err
done

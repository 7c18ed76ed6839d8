#!/bin/bash
# final tree on N GPUs: peer protocol tests, then bench lines of cfg-2 / cfg-5 (/ cfg-4) with the peer-memory all-reduce
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_peer_allreduce_gpu.py -q -m gpu -x > gpurun_out/r2f_peer_tests_n$N.log 2>&1
tail -n 3 gpurun_out/r2f_peer_tests_n$N.log
for spec in ${SPECS:-cfg2:f16 cfg5:f16}; do
  cfg=${spec%%:*}; dt=${spec##*:}
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $cfg --dtype $dt --steps 10 --warmup 3 > gpurun_out/r2_bench_${cfg}_${dt}_n${N}_v3.json 2> gpurun_out/r2_bench_${cfg}_${dt}_n${N}_v3.err
  echo "== $cfg $dt N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${cfg}_${dt}_n${N}_v3.json").read().strip().splitlines()[-1])
    print("$cfg $dt", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "sync", d.get("ranks_in_sync"), d.get("allreduce"))
except Exception as e:
    print("$cfg $dt no line", e)
PY
  tail -n 2 gpurun_out/r2_bench_${cfg}_${dt}_n${N}_v3.err
done

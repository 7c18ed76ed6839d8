#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
for spec in "cfg2 f16" "cfg5 f16" "cfg5 bf16" "cfg4 f16"; do
  set -- $spec
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $1 --dtype $2 --steps 10 --warmup 3 > gpurun_out/r2_bench_$1_$2_n$N.json 2> gpurun_out/r2_bench_$1_$2_n$N.err
  echo "== $1 $2 N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_$1_$2_n$N.json").read().strip().splitlines()[-1])
    print("$1 $2", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "sync", d.get("ranks_in_sync"))
except Exception as e:
    print("$1 $2 no line", e)
PY
done

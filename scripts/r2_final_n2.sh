#!/bin/bash
# last tree on 2 GPUs, launched as the driver launches it
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2_f16_n2_v7.json 2> gpurun_out/r2_bench_cfg2_f16_n2_v7.err
echo rc=$?
python - <<'PY'
import json
d = json.loads([x for x in open("gpurun_out/r2_bench_cfg2_f16_n2_v7.json") if x.startswith("{")][-1])
print("value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "sync", d.get("ranks_in_sync"), d.get("allreduce"), d.get("clocks"))
PY
tail -n 3 gpurun_out/r2_bench_cfg2_f16_n2_v7.err

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r2_bench_cfg2_f16_n1_v7.json 2> gpurun_out/r2_bench_v7.err; echo rc=$?
timeout 300 python bench.py --config cfg4 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg4_f16_n1_v7.json 2>> gpurun_out/r2_bench_v7.err; echo rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cfg2_f16_n1_v7.json", "gpurun_out/r2_bench_cfg4_f16_n1_v7.json"):
    d=json.loads([x for x in open(f) if x.startswith("{")][-1])
    print(round(d["value"]), round(d["ms_per_step"],3), d["step_roofline"], d["roofline"]["frac"], d["roofline"]["frac_burst"])
PY
tail -n 5 gpurun_out/r2_bench_v7.err

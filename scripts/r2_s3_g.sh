#!/bin/bash
mkdir -p gpurun_out
timeout 80 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg2_f16_n1_v8.json 2> gpurun_out/r2_bench_v8.err; echo rc=$?
timeout 60 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_cfg2_reference_v8.json 2>> gpurun_out/r2_bench_v8.err; echo rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cfg2_f16_n1_v8.json", "gpurun_out/r2_bench_cfg2_reference_v8.json"):
    try:
        d = json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(round(d["value"]), round(d["ms_per_step"], 3), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 gpurun_out/r2_bench_v8.err

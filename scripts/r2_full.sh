#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1
tail -n 25 gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
tail -c 600 gpurun_out/r2_bench_cfg2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_cfg2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "roofline", "clocks") if k in d})
for k, v in sorted(d.get("kernel_shares", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])[:12]:
    print("%-26s %5.0f calls %8.3f ms" % (k, v["calls_per_step"], v["ms_per_step"]))
PY

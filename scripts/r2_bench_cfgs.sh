#!/bin/bash
mkdir -p gpurun_out
for spec in "cfg2 f16" "cfg5 f16" "cfg5 bf16" "cfg4 f16" "cfgR f16" "cfgP f16"; do
  set -- $spec
  timeout 900 python bench.py --config $1 --dtype $2 --steps 10 --warmup 3 > gpurun_out/r2_bench_$1_$2.json 2> gpurun_out/r2_bench_$1_$2.err
  echo "== $1 $2 rc=$?"; tail -c 300 gpurun_out/r2_bench_$1_$2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_$1_$2.json").read().strip().splitlines()[-1])
    print("$1 $2", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"), "step frac %.3f" % d["step_roofline"]["frac_of_sustained_bf16"])
except Exception as e:
    print("$1 $2 no line", e)
PY
done

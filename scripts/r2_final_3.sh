#!/bin/bash
# last tree of the round, one GPU: the whole -m gpu suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_gpu_tests.log 2>&1
tail -n 3 gpurun_out/r2h_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -n 1 gpurun_out/r2h_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_cfg2_f16_n1_v6.json 2> gpurun_out/r2_bench_v6.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2_bench_cfg2_f16_n1_v6.json") if x.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
PY

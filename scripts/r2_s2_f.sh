#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s2_gpu_tests.log 2>&1
tail -n 8 gpurun_out/r2s2_gpu_tests.log
for cfg in cfg2 cfg5 cfgP; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s2_bench_${cfg}_w.json 2> gpurun_out/r2s2_bench_${cfg}_w.err
RSR_NO_WAVE=1 timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s2_bench_${cfg}_nw.json 2>> gpurun_out/r2s2_bench_${cfg}_w.err
done
python - <<'PY'
import json
for c in ("cfg2","cfg5","cfgP"):
  for v in ("w","nw"):
    f="gpurun_out/r2s2_bench_%s_%s.json"%(c,v)
    try:
        d=json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2s2_bench_cfg5_w.err

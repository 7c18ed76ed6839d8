#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_kernels_gpu.py tests/test_batchnorm_gpu.py -x -q -m gpu > gpurun_out/r2s2_h_tests.log 2>&1
tail -n 3 gpurun_out/r2s2_h_tests.log
for cfg in cfg2 cfg5; do
timeout 300 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_${cfg}_h.json 2> gpurun_out/r2s2_bench_${cfg}_h.err
done
python - <<'PY'
import json
for c in ("cfg2","cfg5"):
    f="gpurun_out/r2s2_bench_%s_h.json"%c
    try:
        d=json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY

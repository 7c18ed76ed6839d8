#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_frame_models_gpu.py tests/test_gan_frame.py tests/test_batchnorm_gpu.py -x -q -m gpu > gpurun_out/r2s2_k_tests.log 2>&1
tail -n 2 gpurun_out/r2s2_k_tests.log
for m in 0 1; do
RSR_NO_AUX_STREAM=$m timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_k$m.json 2> gpurun_out/r2s2_bench_cfg2_k.err
done
python - <<'PY'
import json
for m in (0,1):
    d=json.loads([x for x in open("gpurun_out/r2s2_bench_cfg2_k%d.json"%m) if x.startswith("{")][-1])
    print("no_aux=%d"%m, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
PY

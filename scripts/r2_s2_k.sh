#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gan_gpu.py -x -q -m gpu -k "fc1 or golden or T100" > gpurun_out/r2s2_k_tests.log 2>&1
tail -n 2 gpurun_out/r2s2_k_tests.log
timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_k.json 2> gpurun_out/r2s2_bench_cfg2_k.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2s2_bench_cfg2_k.json") if x.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["kernel_shares"].get("rsr_fc1_head"))
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/r2s2_j_tests.log 2>&1
tail -n 2 gpurun_out/r2s2_j_tests.log
timeout 300 python scripts/gpu_bench_gemm.py > gpurun_out/r2_gemm_shapes_v1.txt 2>&1
cat gpurun_out/r2_gemm_shapes_v1.txt
timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_j.json 2> gpurun_out/r2s2_bench_cfg2_j.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2s2_bench_cfg2_j.json") if x.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "recurrence" > gpurun_out/r2s2_rec_tests.log 2>&1
tail -n 3 gpurun_out/r2s2_rec_tests.log
RSR_REC_SHAPES="32,100,1024;64,100,1024;64,200,1024;8,100,768;64,100,768" timeout 300 python scripts/gpu_bench_rec.py f16 > gpurun_out/r2_rec_steps_big_v3.txt 2>&1
cat gpurun_out/r2_rec_steps_big_v3.txt
timeout 600 python bench.py --config cfg5 --dtype f16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg5_f16_n1.json 2> gpurun_out/r2s2_bench_cfg5.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2s2_bench_cfg5_f16_n1.json") if x.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3))
PY

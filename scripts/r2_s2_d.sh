#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/r2s2_gan_tests.log 2>&1
tail -n 8 gpurun_out/r2s2_gan_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_wave.json 2> gpurun_out/r2s2_bench_cfg2_wave.err
python - <<'PY'
import json
for f in ["gpurun_out/r2s2_bench_cfg2_wave.json"]:
    d=json.loads([x for x in open(f) if x.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"]); print({k:round(v["ms_per_step"],3) for k,v in d["kernel_shares"].items()})
PY
RSR_WAVE_DSTEP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_wave_dstep.json 2>> gpurun_out/r2s2_bench_cfg2_wave.err
RSR_NO_WAVE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_nowave.json 2>> gpurun_out/r2s2_bench_cfg2_wave.err
python - <<'PY'
import json
for f in ["gpurun_out/r2s2_bench_cfg2_wave_dstep.json","gpurun_out/r2s2_bench_cfg2_nowave.json"]:
    d=json.loads([x for x in open(f) if x.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
tail -5 gpurun_out/r2s2_bench_cfg2_wave.err

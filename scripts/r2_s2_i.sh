#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gan_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/r2s2_i_tests.log 2>&1
tail -n 2 gpurun_out/r2s2_i_tests.log
timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_i.json 2> gpurun_out/r2s2_bench_cfg2_i.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2s2_bench_cfg2_i.json") if x.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
PY
L=$PWD/rsrgan_b200/librsrgan_trace.so
RSR_LIB=$L RSR_WAVE_NBP=48 timeout 60 python scripts/gpu_trace_rec.py 128 512 wavebwd > gpurun_out/r2_wave_trace_bwd_v1.txt 2>&1
cat gpurun_out/r2_wave_trace_bwd_v1.txt
timeout 300 python scripts/gpu_timeline_graph.py cfg2 > gpurun_out/r2_timeline_graph_cfg2_v5.txt 2> /dev/null
head -1 gpurun_out/r2_timeline_graph_cfg2_v5.txt

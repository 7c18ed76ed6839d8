#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_gan_gpu.py tests/test_gan_frame.py tests/test_batchnorm_gpu.py tests/test_frame_models_gpu.py -x -q -m gpu > gpurun_out/r2s2_h_tests.log 2>&1
tail -n 4 gpurun_out/r2s2_h_tests.log
for cfg in cfg2 cfg4; do
timeout 300 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_${cfg}_h.json 2> gpurun_out/r2s2_bench_${cfg}_h.err
done
RSR_NO_HEAD_FUSION=1 timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_bench_cfg2_nohead.json 2>> gpurun_out/r2s2_bench_cfg2_h.err
python - <<'PY'
import json
for c in ("cfg2_h","cfg4_h","cfg2_nohead"):
    f="gpurun_out/r2s2_bench_%s.json"%c
    try:
        d=json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 300 python scripts/gpu_timeline_graph.py cfg2 > gpurun_out/r2_timeline_graph_cfg2_v4.txt 2> /dev/null
head -1 gpurun_out/r2_timeline_graph_cfg2_v4.txt

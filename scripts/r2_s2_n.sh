#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batchnorm_gpu.py tests/test_gan_frame.py tests/test_frame_models_gpu.py -x -q -m gpu > gpurun_out/r2s2_n_tests.log 2>&1
tail -n 5 gpurun_out/r2s2_n_tests.log
timeout 200 python scripts/gpu_bench_bn.py > gpurun_out/r2_bn_bench_v3.jsonl 2> gpurun_out/r2s2_n.err
cat gpurun_out/r2_bn_bench_v3.jsonl

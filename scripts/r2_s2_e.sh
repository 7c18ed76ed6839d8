#!/bin/bash
mkdir -p gpurun_out
RSR_DEBUG=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "wave or recurrence_fwd_bwd" > gpurun_out/r2s2_wave_tests.log 2>&1
tail -n 25 gpurun_out/r2s2_wave_tests.log
timeout 120 python scripts/gpu_bench_wave.py f16 > gpurun_out/r2_wave_steps_v7.txt 2>&1
cat gpurun_out/r2_wave_steps_v7.txt

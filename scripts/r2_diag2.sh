#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_forward or recurrence" > gpurun_out/r2_pair_tests.log 2>&1
tail -n 15 gpurun_out/r2_pair_tests.log
timeout 300 python scripts/gpu_bench_rec.py f16 > gpurun_out/r2_rec_steps_v1.txt 2>&1
cat gpurun_out/r2_rec_steps_v1.txt
for B in 32 128; do
  RSR_LIB=$PWD/rsrgan_b200/librsrgan_trace.so timeout 120 python scripts/gpu_trace_rec.py $B 512 pfwd
done > gpurun_out/r2_trace_v1.txt 2>&1
cat gpurun_out/r2_trace_v1.txt

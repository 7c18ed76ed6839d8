#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_forward or recurrence" 2>&1 | tail -n 15
timeout 300 python scripts/gpu_bench_rec.py f16 2>&1
for w in pfwd bwd; do
  RSR_LIB=$PWD/rsrgan_b200/librsrgan_trace.so timeout 120 python scripts/gpu_trace_rec.py 128 512 $w
done
} > gpurun_out/r2_pair_v5.txt 2>&1
cat gpurun_out/r2_pair_v5.txt

#!/bin/bash
mkdir -p gpurun_out
for nch in ${CHAINS:-1}; do
  export RSR_PAIR_CHAINS=$nch
  echo "=== RSR_PAIR_CHAINS=$nch"
  timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_forward" 2>&1 | tail -n 3
  timeout 300 python scripts/gpu_bench_rec.py f16 2>&1 | grep -E "fused|^ +(32|128) +(100|200) +512|^ +128 +100 +256"
  for B in 128; do
    RSR_LIB=$PWD/rsrgan_b200/librsrgan_trace.so timeout 120 python scripts/gpu_trace_rec.py $B 512 pfwd
  done
done > gpurun_out/r2_pair_v4.txt 2>&1
cat gpurun_out/r2_pair_v4.txt

#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_forward or recurrence" 2>&1 | tail -n 5
timeout 300 python scripts/gpu_bench_rec.py f16 2>&1 | grep -E "fused|^ +(32|128) +(100|200) +512|^ +128 +100 +256"
RSR_FAST_GATES=1 timeout 300 python scripts/gpu_bench_rec.py f16 2>&1 | grep -E "^ +(128) +(100) +512"
RSR_LIB=$PWD/rsrgan_b200/librsrgan_trace.so timeout 120 python scripts/gpu_trace_rec.py 128 512 bwd
RSR_FAST_GATES=1 RSR_LIB=$PWD/rsrgan_b200/librsrgan_trace.so timeout 120 python scripts/gpu_trace_rec.py 128 512 pfwd
} > gpurun_out/r2_pair_v6.txt 2>&1
cat gpurun_out/r2_pair_v6.txt
{
for dt in f16 bf16; do for fast in 0 1; do
  RSR_FAST_GATES=$fast timeout 900 python scripts/gpu_measure_parity.py $dt all 2>&1 | grep -v Warning
done; done
} > gpurun_out/r2_parity_measured_v0.jsonl 2>&1
cat gpurun_out/r2_parity_measured_v0.jsonl

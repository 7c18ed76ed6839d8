#!/bin/bash
# round-2 diagnostics, call 1: MMA pair probe, DSMEM all-gather, GEMM shapes vs cuBLAS, recurrence step times + phase trace
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_diag1_gpu.txt 2>&1
timeout 120 ./scripts/micro/mma_pair_probe > gpurun_out/r2_mma_pair_probe.txt 2>&1
timeout 120 ./scripts/micro/dsmem_bw > gpurun_out/r2_dsmem_bw.txt 2>&1
timeout 300 python scripts/gpu_bench_gemm.py f16 > gpurun_out/r2_gemm_shapes_v0.txt 2>&1
timeout 300 python scripts/gpu_bench_rec.py f16 > gpurun_out/r2_rec_steps_v0.txt 2>&1
for B in 16 32; do for w in fwd bwd; do
  RSR_LIB=$PWD/rsrgan_b200/librsrgan_trace.so timeout 120 python scripts/gpu_trace_rec.py $B 512 $w
done; done > gpurun_out/r2_trace_v0.txt 2>&1
tail -n 30 gpurun_out/r2_mma_pair_probe.txt

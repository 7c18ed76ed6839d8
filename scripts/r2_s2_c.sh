#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rsrgan_b200/librsrgan_trace.so
{
RSR_LIB=$L timeout 60 python scripts/gpu_trace_rec.py 96 512 bwd
RSR_LIB=$L RSR_WAVE_NBP=32 timeout 60 python scripts/gpu_trace_rec.py 96 512 wavebwd
RSR_LIB=$L RSR_WAVE_NBP=48 timeout 60 python scripts/gpu_trace_rec.py 96 512 wavebwd
} > gpurun_out/r2_wave_trace_bwd_v0.txt 2>&1
cat gpurun_out/r2_wave_trace_bwd_v0.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/micro/red_probe > gpurun_out/r2_red_probe.txt 2>&1
cat gpurun_out/r2_red_probe.txt
RSR_REC_SHAPES="16,100,1024;32,100,1024;64,100,1024;64,200,1024;8,100,768;16,100,768;32,100,768;64,100,768" timeout 300 python scripts/gpu_bench_rec.py f16 > gpurun_out/r2_rec_steps_big_v0.txt 2>&1
cat gpurun_out/r2_rec_steps_big_v0.txt

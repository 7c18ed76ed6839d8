#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/gpu_measure_long_T.py > gpurun_out/r2_long_T_v1.jsonl 2> gpurun_out/r2_long_T_v1.err
cat gpurun_out/r2_long_T_v1.jsonl; tail -n 5 gpurun_out/r2_long_T_v1.err

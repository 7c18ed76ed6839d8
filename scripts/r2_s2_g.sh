#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gan_gpu.py -x -q -m gpu -k "gauss or golden or graph or wave" > gpurun_out/r2s2_g_tests.log 2>&1
tail -n 4 gpurun_out/r2s2_g_tests.log
timeout 300 python scripts/gpu_timeline_graph.py cfg2 > gpurun_out/r2_timeline_graph_cfg2_v2.txt 2> gpurun_out/r2s2_g.err
head -3 gpurun_out/r2_timeline_graph_cfg2_v2.txt; tail -3 gpurun_out/r2s2_g.err

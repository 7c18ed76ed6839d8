#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_reference_graph_gpu.py -q -m gpu > gpurun_out/r2s3_f_tests.log 2>&1
tail -n 3 gpurun_out/r2s3_f_tests.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_reference_graph_gpu.py -q -m gpu -s > gpurun_out/r2s3_d_tests.log 2>&1
grep -a "reference-graph deviations\|passed\|failed\|Error\|assert" gpurun_out/r2s3_d_tests.log | head -30

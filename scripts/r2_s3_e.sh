#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_reference_graph_gpu.py -q -m gpu -s -k training_loop > gpurun_out/r2s3_e_tests.log 2>&1
grep -a "reference-loop deviations\|passed\|failed\|Error\|assert\|^E " gpurun_out/r2s3_e_tests.log | head -30

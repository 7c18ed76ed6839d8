#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_ops_gpu.py tests/test_kernels_gpu.py tests/test_batchnorm_gpu.py tests/test_abi.py -q -m gpu > gpurun_out/r2_t1.log 2>&1
tail -n 40 gpurun_out/r2_t1.log

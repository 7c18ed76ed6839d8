#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_batchnorm_gpu.py tests/test_frame_models_gpu.py tests/test_abi.py -q -m gpu -x > gpurun_out/r2_t3.log 2>&1
tail -n 40 gpurun_out/r2_t3.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gan_gpu.py -q -m gpu -k "empty_utterances or long_utterance" > gpurun_out/r2s3_b_tests.log 2>&1
tail -n 30 gpurun_out/r2s3_b_tests.log

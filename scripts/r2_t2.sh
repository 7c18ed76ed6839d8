#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_frame_models_gpu.py tests/test_abi.py -q -m gpu > gpurun_out/r2_t2.log 2>&1
tail -n 40 gpurun_out/r2_t2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_timeline_graph.py cfg2 > gpurun_out/r2_timeline_graph_cfg2_v1.txt 2>gpurun_out/r2_timeline.err
head -3 gpurun_out/r2_timeline_graph_cfg2_v1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_v1.csv python scripts/ncu_one_step.py cfg2 2 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstmp_fwd_pair -c 1 -f -o gpurun_out/r2_recfwd_pair python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstmp_bwd_pair -c 1 -f -o gpurun_out/r2_recbwd_pair python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"lsgan_mse|fc1_fwd|fc1_bwd|stage_input|unstage|clip_update|seg_sumsq_kernel|colsum|fill32|transpose16" -c 40 -f -o gpurun_out/r2_hbm_kernels python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -n 3 gpurun_out/ncu2.log gpurun_out/ncu4.log

#!/bin/bash
# round-2 evidence refresh: fixed 2-D RCED test, ncu DRAM table of the stream kernels, launch list, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frame_models_gpu.py -q -m gpu -k splice11 > gpurun_out/r2_t4.log 2>&1
tail -n 5 gpurun_out/r2_t4.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:"lsgan_mse|fc1_fwd|fc1_bwd|stage_input|unstage|clip_update|seg_sumsq_kernel|colsum|fill32|transpose16|add_cast|cmvn" -c 200 --csv --log-file gpurun_out/r2_hbm_cfg2.csv python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu5.log 2>&1
python scripts/ncu_kernel_table.py gpurun_out/r2_hbm_cfg2.csv > gpurun_out/r2_hbm_kernels.csv
cat gpurun_out/r2_hbm_kernels.csv
timeout 600 ncu --metrics $M --clock-control none -k regex:"bn_|affine_act|conv_" -c 300 --csv --log-file gpurun_out/r2_hbm_rcedbn.csv python scripts/ncu_rced_bn_step.py > gpurun_out/ncu6.log 2>&1
python scripts/ncu_kernel_table.py gpurun_out/r2_hbm_rcedbn.csv > gpurun_out/r2_hbm_kernels_rced_bn.csv
cat gpurun_out/r2_hbm_kernels_rced_bn.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_v2.csv python scripts/ncu_one_step.py cfg2 2 > gpurun_out/ncu7.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2_f16_n1_v2.json 2> gpurun_out/r2_bench_v2.err
tail -c 1500 gpurun_out/r2_bench_cfg2_f16_n1_v2.json

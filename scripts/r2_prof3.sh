#!/bin/bash
# round-2 evidence, second session: launch list, ncu --set full of the wavefront kernels / the backward pair kernel, ncu
# sections of the GEMM launches, graph timeline, bench lines of every config (each with cpu_baseline).  Reports are turned
# into raw csv pages on the box (gpurun brings back at most 64 MiB).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_v3.csv python scripts/ncu_one_step.py cfg2 2 > gpurun_out/ncu_p3_1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstmp_wave_fwd -c 1 -f -o gpurun_out/r2_wave_fwd python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu_p3_2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:lstmp_bwd_pair -c 1 -f -o gpurun_out/r2_recbwd_pair_v2 python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu_p3_3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:lstmp_wave_bwd -c 1 -f -o gpurun_out/r2_wave_bwd python scripts/ncu_one_step.py cfgP 1 > gpurun_out/ncu_p3_4.log 2>&1
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k regex:gemm -c 90 -f -o gpurun_out/r2_gemm python scripts/ncu_one_step.py cfg2 1 > gpurun_out/ncu_p3_5.log 2>&1
for r in r2_wave_fwd r2_recbwd_pair_v2 r2_wave_bwd r2_gemm; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2> /dev/null
  ls -la gpurun_out/$r.ncu-rep gpurun_out/${r}_raw.csv
done
rm -f gpurun_out/r2_gemm.ncu-rep gpurun_out/r2_recbwd_pair_v2.ncu-rep gpurun_out/r2_wave_bwd.ncu-rep
timeout 300 python scripts/gpu_timeline_graph.py cfg2 > gpurun_out/r2_timeline_graph_cfg2_v3.txt 2> gpurun_out/r2_p3_tl.err
head -1 gpurun_out/r2_timeline_graph_cfg2_v3.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2_f16_n1_v3.json 2> gpurun_out/r2_bench_v3.err
for cfg in cfg5 cfg4 cfgP cfgR; do
timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r2_bench_${cfg}_f16_n1_v3.json 2>> gpurun_out/r2_bench_v3.err
done
timeout 600 python bench.py --config cfg5 --dtype bf16 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_bf16_n1_v3.json 2>> gpurun_out/r2_bench_v3.err
timeout 600 python -m pytest tests/test_gan_gpu.py -x -q -m gpu -k "wavefront" > gpurun_out/r2_p3_tests.log 2>&1
tail -n 3 gpurun_out/r2_p3_tests.log
du -sh gpurun_out
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*_n1_v3.json")):
    try:
        d=json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY

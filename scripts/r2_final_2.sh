#!/bin/bash
# final tree of the round (statistics epilogue in), one GPU: the whole -m gpu suite, smoke(), default bench line, reference arm,
# ncu duration / DRAM table of the two batch_norm statistics forms, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_gpu_tests.log 2>&1
tail -n 3 gpurun_out/r2g_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -n 1 gpurun_out/r2g_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_cfg2_f16_n1_v5.json 2> gpurun_out/r2_bench_v5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg2_reference_v5.json 2>> gpurun_out/r2_bench_v5.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm|bn_" -c 60 --csv --log-file gpurun_out/r2_bn_epilogue_raw.csv python scripts/ncu_bn_epilogue.py > gpurun_out/ncu_g1.log 2>&1
python scripts/ncu_kernel_table.py gpurun_out/r2_bn_epilogue_raw.csv > gpurun_out/r2_bn_epilogue_ncu.csv
cat gpurun_out/r2_bn_epilogue_ncu.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_v5.csv python scripts/ncu_one_step.py cfg2 2 > gpurun_out/ncu_g2.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cfg2_f16_n1_v5.json","gpurun_out/r2_bench_cfg2_reference_v5.json"):
    try:
        d=json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(d["value"]), round(d.get("ms_per_step",0),3), d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("frac_burst"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY

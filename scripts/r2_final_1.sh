#!/bin/bash
# final tree, one GPU: the whole -m gpu suite, smoke(), the default bench line, launch list and graph timeline
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_gpu_tests.log 2>&1
tail -n 3 gpurun_out/r2f_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -n 1 gpurun_out/r2f_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_cfg2_f16_n1_v4.json 2> gpurun_out/r2_bench_v4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg2_reference_v4.json 2>> gpurun_out/r2_bench_v4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_v4.csv python scripts/ncu_one_step.py cfg2 2 > gpurun_out/ncu_f1.log 2>&1
timeout 300 python scripts/gpu_timeline_graph.py cfg2 > gpurun_out/r2_timeline_graph_cfg2_v6.txt 2> /dev/null
head -1 gpurun_out/r2_timeline_graph_cfg2_v6.txt
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cfg2_f16_n1_v4.json","gpurun_out/r2_bench_cfg2_reference_v4.json"):
    try:
        d=json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(d["value"]), round(d.get("ms_per_step",0),3), d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY

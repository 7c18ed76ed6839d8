#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_parity_measured_v1.jsonl
for dt in f16 bf16; do
  RSR_FAST_GATES=1 timeout 600 python scripts/gpu_measure_parity.py $dt all >> gpurun_out/r2_parity_measured_v1.jsonl 2>> gpurun_out/r2_parity_v1.err
done
wc -l gpurun_out/r2_parity_measured_v1.jsonl; tail -n 3 gpurun_out/r2_parity_v1.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_parity_measured_v1.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print({k:(round(v,6) if isinstance(v,float) else v) for k,v in d.items() if not isinstance(v,(dict,list))})
PY

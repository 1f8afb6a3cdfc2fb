#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload batch512 --steps 10 > gpurun_out/r2q_batch512.json 2> gpurun_out/r2q_batch512.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q_batch512.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","batch_stroke"): print(k, d.get(k))
PY
tail -5 gpurun_out/r2q_batch512.err

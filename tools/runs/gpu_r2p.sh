#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload strokes4k --steps 10 > gpurun_out/r2p_strokes4k.json 2> gpurun_out/r2p_strokes4k.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2p_strokes4k.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","batch_stroke","plotter_stroke"): print(k, d.get(k))
PY
tail -5 gpurun_out/r2p_strokes4k.err

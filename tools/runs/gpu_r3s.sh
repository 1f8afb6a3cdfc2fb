#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "analytic or heptagram or polygons or layers or fig_kats or negative or rows_above" > gpurun_out/r3s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3s_pytest.log; tail -2 gpurun_out/r3s_pytest.log
for args in "" "--format graya8p"; do
  timeout 300 python bench.py $args --steps 30 --kernel-only > gpurun_out/r3s_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3s_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),4))
PY
done

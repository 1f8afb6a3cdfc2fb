#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rgba or analytic or heptagram or layers or config1 or config3 or polygons" > gpurun_out/r3l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3l_pytest.log; tail -2 gpurun_out/r3l_pytest.log
for args in "--format rgba8p" "--workload strokes4k" "--workload batch512" ""; do
  timeout 600 python bench.py $args --steps 20 --kernel-only > gpurun_out/r3l_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3l_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "tfrac", r.get("traffic_frac"))
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
for args in "--format rgba8p" "--format graya8p" "--workload strokes4k" ""; do
  timeout 600 python bench.py $args --steps 20 --kernel-only > gpurun_out/r2m_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2m_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram matte8", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3), "tfrac", r.get("traffic_frac"))
PY
done

#!/bin/bash
# composite formats after a tile-kernel change: parity subset, then the kernel-only legs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "rgba or analytic or polygons or layers or config3 or config1 or fishy or heptagram or curved or wide or srgb or small" > gpurun_out/r2s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2s_pytest.log
tail -3 gpurun_out/r2s_pytest.log
for args in "--format rgba8p" "--format graya8p" "--workload strokes4k" ""; do
  timeout 600 python bench.py $args --steps 20 --kernel-only > gpurun_out/r2s_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2s_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram matte8", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3), "tfrac", r.get("traffic_frac"))
PY
done

#!/bin/bash
# raster_bins window width: 64 against 128 columns
mkdir -p gpurun_out
FTL_BIN_WC=64 timeout 1200 python -m pytest tests -m gpu -x -q -k "config4 or config5_many or wide or polygons or config3 or curved or layers" > gpurun_out/r2x_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2x_pytest.log
tail -3 gpurun_out/r2x_pytest.log
for wc in 128 64; do
  for args in "--workload fishy256" "--workload strokes4k" "--workload batch512" "--workload bigraster"; do
    FTL_BIN_WC=$wc timeout 600 python bench.py $args --steps 10 --kernel-only > gpurun_out/r2x_tmp.json 2>/dev/null
    python - "$wc $args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2x_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1], {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3))
PY
  done
done

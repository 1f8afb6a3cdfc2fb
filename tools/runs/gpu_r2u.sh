#!/bin/bash
# where should jobs of 9..64 edges go: raster_tiles (direct) or raster_bins?
mkdir -p gpurun_out
for dm in 64 32 16 8; do
  for args in "--workload fishy256" "--workload strokes4k" "--workload batch512"; do
    FTL_DIRECT_MAX=$dm timeout 600 python bench.py $args --steps 10 --kernel-only > gpurun_out/r2u_tmp.json 2>/dev/null
    python - "$dm $args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2u_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1], {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3))
PY
  done
done

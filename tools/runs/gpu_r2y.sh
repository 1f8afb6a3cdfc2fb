#!/bin/bash
mkdir -p gpurun_out
for wc in 64 32; do
  for args in "--workload fishy256" "--workload batch512"; do
    FTL_BIN_WC=$wc timeout 600 python bench.py $args --steps 10 --kernel-only > gpurun_out/r2y_tmp.json 2>/dev/null
    python - "$wc $args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2y_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1], {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3))
PY
  done
done

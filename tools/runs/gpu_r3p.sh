#!/bin/bash
mkdir -p gpurun_out
for args in "" "--format rgba8p" "--workload batch512"; do
  timeout 600 python bench.py $args --steps 30 --kernel-only > gpurun_out/r3p_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3p_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),4))
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rgba or analytic or heptagram or layers or config1 or config3" > gpurun_out/r3i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3i_pytest.log; tail -2 gpurun_out/r3i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for args in "--format rgba8p" "--workload strokes4k"; do
  timeout 600 python bench.py $args --steps 20 --kernel-only > gpurun_out/r3i_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3i_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1], {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "tfrac", r.get("traffic_frac"))
PY
done

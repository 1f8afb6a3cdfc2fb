#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "config4 or config5_many or wide or polygons or curved or config3" > gpurun_out/r3o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3o_pytest.log; tail -2 gpurun_out/r3o_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/r3o_launches_batch512.csv python bench.py --workload batch512 --steps 3 --warmup 3 --kernel-only > gpurun_out/r3o_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3o_launches_batch512.csv 2>&1 | grep -E "bin_fill|edge_build|topkey"
for args in "--workload batch512" "--workload fishy256"; do
  timeout 600 python bench.py $args --steps 20 --kernel-only > gpurun_out/r3o_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3o_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1], {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4))
PY
done

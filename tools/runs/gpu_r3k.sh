#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3k_pytest.log
tail -3 gpurun_out/r3k_pytest.log
for args in "--workload bigraster" "--workload batch512" "--workload fishy256"; do
  timeout 600 python bench.py $args --steps 10 --kernel-only > gpurun_out/r3k_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3k_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1], {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4))
PY
done
FTL_BENCH_BAND=3/8 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r3k_launches_band.csv python bench.py --workload bigraster --steps 3 --warmup 3 --kernel-only > gpurun_out/r3k_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3k_launches_band.csv 2>&1 | grep -E "cull|flatten"
FTL_BENCH_BAND=3/8 python bench.py --workload bigraster --steps 5 --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('band 3/8', {k:round(d.get(k),4) for k in ('value','ms_per_step')}, d['roofline']['avg_launch_ms'])"
for args in "--format rgba8p" "--workload strokes4k" ""; do
  timeout 600 python bench.py $args --steps 20 --kernel-only > gpurun_out/r3k_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3k_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "tfrac", r.get("traffic_frac"))
PY
done

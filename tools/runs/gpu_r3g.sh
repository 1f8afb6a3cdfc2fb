#!/bin/bash
# after the bin_fill pipelining and the new stroke tests: full suite, smoke, config 5 / config 4 timings, launch list of config 5
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3g_pytest.log
tail -3 gpurun_out/r3g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for args in "--workload bigraster" "--workload batch512" "--workload strokes4k"; do
  timeout 600 python bench.py $args --steps 10 --kernel-only > gpurun_out/r3g_tmp.json 2>/dev/null
  python - "$args" <<PY
import json,sys
d=json.loads(open("gpurun_out/r3g_tmp.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(sys.argv[1] or "heptagram", {k:round(d.get(k),4) for k in ("value","ms_per_step")}, "tile_ms", round(r.get("avg_launch_ms",0),4), "frac", round(r.get("frac",0),3))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/r3g_launches_bigraster.csv python bench.py --workload bigraster --steps 3 --warmup 3 --kernel-only > gpurun_out/r3g_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3g_launches_bigraster.csv 2>&1 | grep -E "bin_fill|edge_build|raster_bins"
for wc in 64 128; do
  FTL_BIN_WC=$wc timeout 600 python bench.py --workload batch512 --format rgba8p --steps 10 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('batch512 rgba8p wc=$wc', {k:round(d.get(k),4) for k in ('value','ms_per_step')}, 'tile_ms', round(r.get('avg_launch_ms',0),4))"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "config5 or ctx or bands" > gpurun_out/r3f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3f_pytest.log
tail -3 gpurun_out/r3f_pytest.log
FTL_BENCH_BAND=3/8 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r3f_launches_band.csv python bench.py --workload bigraster --steps 3 --warmup 3 --kernel-only > gpurun_out/r3f_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3f_launches_band.csv 2>&1 | grep -E "cull|bin_fill|edge_build"
for b in 3/8 0/8 7/8; do FTL_BENCH_BAND=$b python bench.py --workload bigraster --steps 5 --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$b', {k:round(d.get(k),4) for k in ('value','ms_per_step')}, d['roofline']['avg_launch_ms'])"; done

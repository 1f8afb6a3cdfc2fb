#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "flatten or config4 or curved or config1" > gpurun_out/r2z_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2z_pytest.log
tail -3 gpurun_out/r2z_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/r2z_launches_batch512.csv python bench.py --workload batch512 --steps 3 --warmup 3 --kernel-only > gpurun_out/r2z_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r2z_launches_batch512.csv 2>&1 | head -30
timeout 600 python bench.py --workload batch512 --steps 10 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:round(d.get(k),4) for k in ('value','ms_per_step')})"

#!/bin/bash
# launch list of one rank of an 8-band split of config 5, played on one GPU
mkdir -p gpurun_out
FTL_BENCH_BAND=3/8 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r3e_launches_band.csv python bench.py --workload bigraster --steps 3 --warmup 3 --kernel-only > gpurun_out/r3e_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3e_launches_band.csv 2>&1 | head -40
FTL_BENCH_BAND=3/8 python bench.py --workload bigraster --steps 5 --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:round(d.get(k),4) for k in ('value','ms_per_step')}, d['roofline']['avg_launch_ms'])"

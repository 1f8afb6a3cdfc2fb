#!/bin/bash
mkdir -p gpurun_out
for wl in batch512 bigraster; do
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/r3c_launches_$wl.csv python bench.py --workload $wl --steps 3 --warmup 3 --kernel-only > gpurun_out/r3c_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3c_launches_$wl.csv 2>&1 | head -30
done

#!/bin/bash
mkdir -p gpurun_out
python tools/stroke_probe.py 4096 5 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file gpurun_out/r3d_launches_stroke.csv env FTL_DEVICE_STROKE=1 python tools/stroke_probe.py 4096 1 > gpurun_out/r3d_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r3d_launches_stroke.csv 2>&1 | head -50

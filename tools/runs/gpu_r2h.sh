#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/small_diag.py > gpurun_out/r2h_diag.log 2>&1
cat gpurun_out/r2h_diag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_fill -s 10 -c 1 -o gpurun_out/r2h_small python tools/one_fill_probe.py 20 > gpurun_out/r2h_ncu.log 2>&1
echo "ncu exit $?"

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2k_bins_b512 \
   python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only > gpurun_out/r2k_full_b512.log 2>&1
echo "full b512 exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2k_launches_batch512.csv \
     python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only > gpurun_out/r2k_ll_batch512.log 2>&1
tail -2 gpurun_out/r2k_ll_batch512.log

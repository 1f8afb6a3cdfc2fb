#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2d_bins_b512 \
   python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only > gpurun_out/r2d_full_b512.log 2>&1
echo "full b512 exit $?"
FTL_BIN_WC=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2d_bins_big128 \
   python bench.py --workload bigraster --steps 2 --warmup 3 --kernel-only > gpurun_out/r2d_full_big.log 2>&1
echo "full big exit $?"

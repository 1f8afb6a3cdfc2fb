#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2j_bins_b512 \
   python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only > gpurun_out/r2j_full_b512.log 2>&1
echo "full b512 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2j_bins_big \
   python bench.py --workload bigraster --steps 2 --warmup 3 --kernel-only > gpurun_out/r2j_full_big.log 2>&1
echo "full big exit $?"
for wl in bigraster batch512; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2j_launches_$wl.csv \
     python bench.py --workload $wl --steps 2 --warmup 3 --kernel-only > gpurun_out/r2j_ll_$wl.log 2>&1
done

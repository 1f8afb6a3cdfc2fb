#!/bin/bash
# profiles of the binned kernel: launch lists + one full capture per workload
mkdir -p gpurun_out
for wl in bigraster batch512 fishy256; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_$wl.csv \
     python bench.py --workload $wl --steps 2 --warmup 3 --kernel-only > gpurun_out/r2b_ll_$wl.log 2>&1
  echo "$wl launch list exit $?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2b_bins_big \
   python bench.py --workload bigraster --steps 2 --warmup 3 --kernel-only > gpurun_out/r2b_full_big.log 2>&1
echo "full big exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2b_bins_b512 \
   python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only > gpurun_out/r2b_full_b512.log 2>&1
echo "full b512 exit $?"
ls -la gpurun_out/r2b_*

#!/bin/bash
# source-level instruction counts of raster_bins on the two Matte8 batch workloads
mkdir -p gpurun_out
for wl in batch512 fishy256; do
  ncu --set full --clock-control none --import-source on -k regex:raster_bins -s 3 -c 1 -o gpurun_out/r2w_$wl python bench.py --workload $wl --steps 2 --warmup 3 --kernel-only > gpurun_out/r2w_$wl.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2w_$wl.ncu-rep > gpurun_out/r2w_${wl}_summary.txt 2>&1
  ncu -i gpurun_out/r2w_$wl.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r2w_${wl}_source.csv.gz
  rm -f gpurun_out/r2w_$wl.ncu-rep
done
ls -la gpurun_out/r2w_*

#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:raster_tiles -s 4 -c 1 -o gpurun_out/r2t_rgba python bench.py --format rgba8p --steps 2 --warmup 3 --kernel-only > gpurun_out/r2t.log 2>&1
python tools/ncu_summary.py gpurun_out/r2t_rgba.ncu-rep > gpurun_out/r2t_rgba_summary.txt 2>&1
ncu -i gpurun_out/r2t_rgba.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r2t_rgba_source.csv.gz
rm -f gpurun_out/r2t_rgba.ncu-rep
cat gpurun_out/r2t_rgba_summary.txt

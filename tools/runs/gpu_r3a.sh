#!/bin/bash
mkdir -p gpurun_out
for k in edge_build bin_fill flatten_ops; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r3a_$k python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only > gpurun_out/r3a_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/r3a_$k.ncu-rep > gpurun_out/r3a_${k}_summary.txt 2>&1
  ncu -i gpurun_out/r3a_$k.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r3a_${k}_source.csv.gz
  rm -f gpurun_out/r3a_$k.ncu-rep
done

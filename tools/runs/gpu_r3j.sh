#!/bin/bash
mkdir -p gpurun_out
FTL_BENCH_BAND=3/8 ncu --set full --clock-control none --import-source on -k regex:cull_op_extents -s 1 -c 1 -o gpurun_out/r3j_cull python bench.py --workload bigraster --steps 2 --warmup 2 --kernel-only > gpurun_out/r3j.log 2>&1
python tools/ncu_summary.py gpurun_out/r3j_cull.ncu-rep > gpurun_out/r3j_cull_summary.txt 2>&1
ncu -i gpurun_out/r3j_cull.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r3j_cull_source.csv.gz
ncu -i gpurun_out/r3j_cull.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r3j_cull_raw.csv
rm -f gpurun_out/r3j_cull.ncu-rep
cat gpurun_out/r3j_cull_summary.txt

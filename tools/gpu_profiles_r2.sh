#!/bin/bash
# Round-2 evidence in one call: bench lines (both arms), sustained run, ncu launch lists and one full capture per workload.
mkdir -p gpurun_out
O=gpurun_out
python bench.py > $O/p2_bench_heptagram.json 2> $O/p2_bench_heptagram.err
python bench.py --impl reference > $O/p2_bench_heptagram_reference_arm.json 2>> $O/p2_bench_heptagram.err
python bench.py --steps 3000 --no-secondary > $O/p2_bench_heptagram_sustained.json 2>> $O/p2_bench_heptagram.err
for wl in batch512 fishy256 strokes4k bigraster latency; do
  python bench.py --workload $wl --steps 10 > $O/p2_bench_$wl.json 2> $O/p2_bench_$wl.err
done
python bench.py --workload bigraster --impl reference --steps 1 > $O/p2_bench_bigraster_reference_arm.json 2>> $O/p2_bench_bigraster.err
python bench.py --workload batch512 --impl reference --steps 5 > $O/p2_bench_batch512_reference_arm.json 2>> $O/p2_bench_batch512.err
python bench.py --format rgba8p > $O/p2_bench_heptagram_rgba8p.json 2> $O/p2_bench_heptagram_rgba8p.err
python bench.py --format graya8p > $O/p2_bench_heptagram_graya8p.json 2>> $O/p2_bench_heptagram_rgba8p.err
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/p2_smi.csv 2>&1
# launch lists (kernel-only legs)
for wl in heptagram batch512 bigraster strokes4k fishy256; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file $O/p2_launches_$wl.csv \
     python bench.py --workload $wl --steps 3 --warmup 3 --kernel-only > $O/p2_ll_$wl.log 2>&1
done
# one full capture of the dominant kernel per workload; the raw metric page and the (gzipped) source page are exported here so
# that what travels back stays under gpurun's 64 MiB, the reports themselves are dropped
cap() {  # name kernel-regex skip cmd...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/p2_full_$name "$@" >> $O/p2_full.log 2>&1
  ncu -i $O/p2_full_$name.ncu-rep --page raw --csv > $O/p2_raw_$name.csv 2>> $O/p2_full.log
  ncu -i $O/p2_full_$name.ncu-rep --page source --csv 2>> $O/p2_full.log | gzip -9 > $O/p2_source_$name.csv.gz
  rm -f $O/p2_full_$name.ncu-rep
}
cap heptagram raster_tiles 4 python bench.py --steps 2 --warmup 3 --kernel-only
cap heptagram_rgba8p raster_tiles 4 python bench.py --format rgba8p --steps 2 --warmup 3 --kernel-only
cap batch512 raster_bins 3 python bench.py --workload batch512 --steps 2 --warmup 3 --kernel-only
cap bigraster raster_bins 3 python bench.py --workload bigraster --steps 2 --warmup 3 --kernel-only
cap strokes4k raster_bins 3 python bench.py --workload strokes4k --steps 2 --warmup 3 --kernel-only
cap fishy256 raster_bins 3 python bench.py --workload fishy256 --steps 2 --warmup 3 --kernel-only
cap small small_fill 20 python tools/one_fill_probe.py 30
cap stroke_segments stroke_segments 2 env FTL_DEVICE_STROKE=1 python tools/stroke_probe.py 4096 1
# the device stroker: kernel list of one batched stroke call, and device against host wall clock
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file $O/p2_launches_stroke.csv env FTL_DEVICE_STROKE=1 python tools/stroke_probe.py 4096 1 > $O/p2_ll_stroke.log 2>&1
python tools/stroke_probe.py 4096 5 > $O/p2_stroke_probe.txt 2>&1
du -sh $O; ls -la $O/p2_* | awk '{print $5, $9}'

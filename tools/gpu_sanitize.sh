#!/bin/bash
# compute-sanitizer over the kernels new in round 2 (binned tiles with look-back and 64-column windows, small fills, culling,
# sRGB, area probe, the device stroker, the cp.async staging of Rgba8p alpha-0 spans)
mkdir -p gpurun_out
SEL='small_fill or random_polygons or curved_paths or wide_dense or config5_many or fill_layers or debug_area or srgb or ctx_fill or wide_raster or clear_after or speculative or batch_stroke or overlapping or fig_kats or device_stroker_outline or device_stroker_degenerate or analytic_rows or config4_batch_matches or strict_vid_mode_a'
for tool in memcheck racecheck synccheck; do
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/san_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/san_$tool.log | tail -3
done
# the 4K Rgba8p heptagram drives the staged alpha-0 spans (long rows); memcheck + racecheck on a short run
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python bench.py --format rgba8p --steps 1 --warmup 1 --kernel-only --batch 2 > gpurun_out/san_rgba_$tool.log 2>&1
  echo "rgba8p $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san_rgba_$tool.log | tail -2
done

#!/bin/bash
# compute-sanitizer over the kernels new in round 2 (binned tiles with look-back, small fills, culling, sRGB, area probe)
mkdir -p gpurun_out
SEL='small_fill or random_polygons or curved_paths or wide_dense or config5_many or fill_layers or debug_area or srgb or ctx_fill or wide_raster or clear_after or speculative or batch_stroke or overlapping or fig_kats'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/san_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/san_$tool.log | tail -3
done

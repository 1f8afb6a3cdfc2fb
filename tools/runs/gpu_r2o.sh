#!/bin/bash
# device stroker: parity tests, then the whole suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "stroke or stroker or libm" > gpurun_out/r2o_stroke.log 2>&1
echo "exit $?" >> gpurun_out/r2o_stroke.log
tail -30 gpurun_out/r2o_stroke.log

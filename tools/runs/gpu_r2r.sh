#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "strict" > gpurun_out/r2r.log 2>&1
echo "exit $?" >> gpurun_out/r2r.log
tail -30 gpurun_out/r2r.log

#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
bash tools/gpu_r2d.sh

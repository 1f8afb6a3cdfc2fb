#!/bin/bash
# Re-entry baseline: the GPU suite and the default bench line on the restored tree.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2n_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
tail -25 gpurun_out/r2n_pytest.log
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -c 3000 gpurun_out/r2n_bench.json

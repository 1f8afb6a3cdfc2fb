#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -12 gpurun_out/r2g_pytest.log
timeout 300 python tools/latency_probe.py > gpurun_out/r2g_latency.json 2> gpurun_out/r2g_latency.err
cat gpurun_out/r2g_latency.json; tail -3 gpurun_out/r2g_latency.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2g_launches_small.csv python tools/one_fill_probe.py 20 > gpurun_out/r2g_onefill.log 2>&1
tail -2 gpurun_out/r2g_onefill.log

#!/bin/bash
# round-2 GPU check: parity suite, then the benches of the configs the binned kernel serves
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
for wl in bigraster batch512 strokes4k fishy256; do
  timeout 600 python bench.py --workload $wl --steps 5 > gpurun_out/r2a_$wl.json 2> gpurun_out/r2a_$wl.err
  echo "$wl exit $?"; tail -c 600 gpurun_out/r2a_$wl.json
done
timeout 600 python bench.py > gpurun_out/r2a_heptagram.json 2> gpurun_out/r2a_heptagram.err
tail -c 400 gpurun_out/r2a_heptagram.json

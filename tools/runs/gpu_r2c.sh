#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
for wc in 128 256; do
for wl in bigraster batch512 strokes4k fishy256; do
  FTL_BIN_WC=$wc timeout 600 python bench.py --workload $wl --steps 5 --kernel-only > gpurun_out/r2f_${wl}_wc$wc.json 2> gpurun_out/r2f_${wl}_wc$wc.err
  echo "$wl wc=$wc exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_${wl}_wc$wc.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","tile_ms_per_launch","gpu_launches")}, (d.get("roofline") or {}).get("avg_launch_ms"))
except Exception as e:
    print("ERR",e)
PY
done
done

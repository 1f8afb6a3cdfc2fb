#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2i_default.json 2> gpurun_out/r2i_default.err ) 2>&1 | grep real
tail -c 1800 gpurun_out/r2i_default.json; tail -3 gpurun_out/r2i_default.err
timeout 600 python bench.py --workload strokes4k --steps 10 > gpurun_out/r2i_strokes4k.json 2> gpurun_out/r2i_strokes4k.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_strokes4k.json").read().strip().splitlines()[-1])
print("strokes4k", {k:d.get(k) for k in ("value","ms_per_step")}, d.get("e2e"), d.get("plotter_stroke"))
PY
tail -3 gpurun_out/r2i_strokes4k.err
timeout 900 python bench.py --workload bigraster --steps 3 > gpurun_out/r2i_bigraster.json 2> gpurun_out/r2i_bigraster.err
tail -c 1500 gpurun_out/r2i_bigraster.json; tail -3 gpurun_out/r2i_bigraster.err

#!/bin/bash
# final sanity of the round: GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest.log
tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-secondary > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("metric","value","unit","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
PY

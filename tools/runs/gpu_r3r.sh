#!/bin/bash
# e2e against the number of host threads that expand the packed read-back
mkdir -p gpurun_out
nproc
for t in 12 8 10 12 8; do
  FTL_HOST_THREADS=$t timeout 300 python bench.py --no-secondary --steps 10 --cpu-fills 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads $t e2e', round(d['e2e']['value'],1), 'Gpx/s', round(d['e2e']['ms_per_step'],2), 'ms')"
done

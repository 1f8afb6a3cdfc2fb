#!/bin/bash
# usage: gpu_scale.sh N  (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2s_default_n$N.json 2> gpurun_out/r2s_default_n$N.err
echo "default n=$N exit $?"; tail -c 1600 gpurun_out/r2s_default_n$N.json; tail -2 gpurun_out/r2s_default_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --workload bigraster --kernel-only > gpurun_out/r2s_bigraster_n$N.json 2> gpurun_out/r2s_bigraster_n$N.err
echo "bigraster n=$N exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_bigraster_n$N.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, d["roofline"]["avg_launch_ms"])
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/r2s_bigraster_n$N.err

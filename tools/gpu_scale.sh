#!/bin/bash
# usage: gpu_scale.sh N  (run under gpurun --gpus N): the default bench line at N GPUs; its `secondary` object carries configs 4 and 5
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2s_default_n$N.json 2> gpurun_out/r2s_default_n$N.err
echo "default n=$N exit $?"; tail -c 2600 gpurun_out/r2s_default_n$N.json; tail -2 gpurun_out/r2s_default_n$N.err

#!/bin/bash
# usage: gpu_gather.sh N (under gpurun --gpus N): the 2-GPU NCCL band test of the suite and the gather timing
N=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/gather_probe.py 2>/dev/null | tail -1 | tee gpurun_out/gather_n$N.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/fuzz_parity.py 424242 8 > gpurun_out/r3q_fuzz.log 2>&1
echo "fuzz exit $?"; tail -4 gpurun_out/r3q_fuzz.log

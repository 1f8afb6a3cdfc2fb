#!/bin/bash
# racecheck on the multi-handle test alone (three engines on one device): does it finish, and does the look-back matter?
mkdir -p gpurun_out
for lb in default 0; do
  if [ $lb = default ]; then unset FTL_BIN_LOOKBACK; else export FTL_BIN_LOOKBACK=$lb; fi
  SECONDS=0; timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ctx_fill" > gpurun_out/r3h_race_$lb.log 2>&1
  echo "lookback=$lb exit $? after ${SECONDS}s"; grep -E "RACECHECK SUMMARY|passed|failed|Potential" gpurun_out/r3h_race_$lb.log | sort | uniq -c | tail -4; tail -1 gpurun_out/r3h_race_$lb.log
done

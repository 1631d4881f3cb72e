#!/bin/bash
# compute-sanitizer over the hand-synchronised kernels; logs under gpurun_out/<tag>_sanitize_*.log
tag=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_cases.py $SANITIZE_CASES > gpurun_out/${tag}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/${tag}_sanitize_${tool}.log
done

#!/bin/bash
# compute-sanitizer over tiny shapes of every kernel (tools/sanitize_target.py); logs -> gpurun_out/sanitize_*.log, to be copied to profiles/.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for part in eval auc exchange retrieval; do
    log=gpurun_out/sanitize_${tool}_${part}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py --part $part > $log 2>&1
    echo "exit $?" >> $log
    echo "== $tool $part: $(grep -E 'part .* ok|ERROR SUMMARY|RACECHECK SUMMARY|exit ' $log | tr '\n' ' ')"
  done
done

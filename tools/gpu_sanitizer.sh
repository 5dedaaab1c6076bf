#!/bin/bash
# compute-sanitizer over one small launch of every kernel (tools/sanitizer_cases.py), tool by tool and case by case
# (separate processes and timeouts: a hang or a fault in one case must not take the others with it).
# Summaries -> gpurun_out/sanitizer_<tool>_<case>.log ; one-line verdicts -> gpurun_out/sanitizer_summary.txt
mkdir -p gpurun_out
: > gpurun_out/sanitizer_summary.txt
timeout 120 python tools/sanitizer_cases.py > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; }
TOOLS=${TOOLS:-"memcheck synccheck racecheck"}
CASES=${CASES:-"gemm conv lstm attention integers frontend model"}
for tool in $TOOLS; do
  for c in $CASES; do
    log=gpurun_out/sanitizer_${tool}_${c}.log
    timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_cases.py $c > $log 2>&1
    rc=$?
    echo "$tool $c rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a gpurun_out/sanitizer_summary.txt
  done
done

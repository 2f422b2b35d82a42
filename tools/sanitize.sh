#!/bin/bash
# compute-sanitizer passes over small launches of every kernel family (run on the GPU box): logs under gpurun_out/.
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python tools/sanitizer_cases.py \
    > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log >> gpurun_out/sanitizer_summary.txt
done
cat gpurun_out/sanitizer_summary.txt

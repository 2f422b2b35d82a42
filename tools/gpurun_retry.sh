#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy/transient)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
cat /tmp/gpurun_last.txt
exit $rc

#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   - retries while the pod answers busy / transient (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient\|no box or slot\|another call" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then
    sleep 90; continue
  fi
  cat /tmp/gpurun_last.txt; exit $rc
done
cat /tmp/gpurun_last.txt; exit 3

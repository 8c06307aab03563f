#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   -- retries while the pod answers busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 120; continue; fi
  echo "$out" | tail -25; exit $rc
done
echo "gave up"; exit 3

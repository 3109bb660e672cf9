#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <command...>   -- retries while the pod answers busy/transient (nothing charged)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 120; continue; fi
  break
done

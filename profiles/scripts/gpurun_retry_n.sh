#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <timeout_s> <command...>   -- like gpurun_retry.sh on N GPUs of one box
G=$1; T=$2; shift 2
for i in $(seq 1 15); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 150; continue; fi
  break
done

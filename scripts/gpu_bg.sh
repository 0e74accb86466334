#!/bin/bash
# usage: scripts/gpu_bg.sh <logfile> <timeout> <command...> -- retries gpurun while the pod answers busy (exit 3 / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|retry in a few minutes" $log; then sleep 90; continue; fi
  break
done
echo "[gpu_bg done]" >> $log

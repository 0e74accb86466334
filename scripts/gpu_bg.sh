#!/bin/bash
# usage: scripts/gpu_bg.sh <logfile> <timeout> [--gpus N] <command...> -- retries gpurun while the pod answers busy
log=$1; shift; to=$1; shift
extra=""
if [ "$1" == "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
for i in $(seq 1 40); do
  gpurun --timeout $to $extra -- "$@" > $log 2>&1
  if grep -q "status=transient\|retry in a few minutes\|status=busy" $log; then sleep 90; continue; fi
  break
done
echo "[gpu_bg done]" >> $log

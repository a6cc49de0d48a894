#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> [--gpus N] <command>   -- retries while the pod answers "transient / busy" (nothing charged)
log=$1; to=$2; shift 2
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift 2; fi
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  grep -q "status=transient\|status=busy\|rc=3" $log || break
  sleep 120
done

#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "transient / busy" (nothing charged)
log=$1; to=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  grep -q "status=transient\|status=busy" $log || break
  sleep 120
done

#!/bin/bash
# usage: gpurun_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers "transient / busy"
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy\|exit code 3" $log; then sleep 45; continue; fi
  break
done

#!/bin/bash
# developer aid: gpurun with retries while the pod answers "busy" (exit 3).  usage: tools/gpurun_retry.sh LOG TIMEOUT 'command'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  echo "attempt $i rc=$rc" >> $log.attempts
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done

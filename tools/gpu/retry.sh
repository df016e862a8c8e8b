#!/bin/bash
# local: keep asking for a GPU box until the call is accepted (exit code 3 = no box free, nothing charged)
# usage: tools/gpu/retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3

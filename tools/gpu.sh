#!/bin/bash
# usage: tools/gpu.sh <timeout_s> <logname> <command...>   (retries while the pod answers "busy")
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/$LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

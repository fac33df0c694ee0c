#!/bin/bash
# usage: [GPUS=n] tools/gpu.sh <timeout_s> <logname> <command...>   (retries while the pod answers "busy")
T=$1; LOG=$2; shift 2
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > gpurun_out/$LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

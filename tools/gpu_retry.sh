#!/bin/bash
# gpu_retry.sh <timeout_s> <command...>: call gpurun, retrying while the pool answers "busy" (exit code 3).
# GPUS=N in the environment asks for an N-GPU box.
T=$1; shift
G=${GPUS:+--gpus $GPUS}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

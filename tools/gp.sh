#!/bin/bash
# Developer helper: run a gpurun command, retrying while the pod answers "busy / draining" (exit code 3).
# usage: tools/gp.sh LOGFILE [--gpus N] TIMEOUT 'command'
log=$1; shift
gpus=""
if [ "$1" == "--gpus" ]; then gpus="--gpus $2"; shift 2; fi
to=$1; shift
for attempt in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $gpus --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
exit 3

#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT 'command': gpurun, retried while the pod answers "no slot right now" (exit code 3)
t=$1; shift
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout $t -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 45
done
exit 3

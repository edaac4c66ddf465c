#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT 'command': gpurun, retried while the pod answers "no slot right now" (exit code 3) or a
# previous call of this repo is still registered (exit code 2 with that message)
t=$1; shift
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1); rc=$?
    echo "$out"
    if [ $rc -eq 3 ]; then sleep 45; continue; fi
    if [ $rc -eq 2 ] && echo "$out" | grep -q "already running"; then sleep 30; continue; fi
    exit $rc
done
exit 3

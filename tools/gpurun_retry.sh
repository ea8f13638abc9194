#!/bin/bash
# Re-issue a gpurun call while the pod answers "no box / slot free right now" (exit code 3: nothing charged).
#   tools/gpurun_retry.sh <log file> <gpurun args...>
log="$1"; shift
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then echo "[gpurun_retry] attempt $attempt rc=$rc" >> "$log"; exit $rc; fi
    sleep 90
done
echo "[gpurun_retry] gave up" >> "$log"
exit 3

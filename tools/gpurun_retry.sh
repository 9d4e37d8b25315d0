#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'   (retries while the pod answers "transient / busy")
for attempt in 1 2 3 4 5 6 7 8; do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    echo "$out" | tail -60
    if echo "$out" | grep -q "status=transient\|status=busy\|no box or slot"; then
        echo "[retry] attempt $attempt was transient; sleeping 90 s"
        sleep 90
        continue
    fi
    break
done

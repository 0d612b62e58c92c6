#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3; nothing is charged): tools/gpu_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
    if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpu_retry] busy (attempt $i), sleeping 60 s"
    sleep 60
done
exit 3

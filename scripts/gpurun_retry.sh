#!/usr/bin/env bash
# usage: scripts/gpurun_retry.sh <timeout_s> <script> [gpus] — retries while the pod answers busy (exit code 3)
t=$1; s=$2; g=${3:-1}
for k in $(seq 1 30); do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "bash $s"; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "bash $s"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3

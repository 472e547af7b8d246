#!/bin/bash
# usage: retry.sh <timeout-seconds> <script> [gpus]   -- re-submits while the pod answers "busy / transient" (nothing charged)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] busy"; sleep 120; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3

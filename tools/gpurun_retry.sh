#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...] -- 'command'
# Re-submits while the pod answers "transient"/busy (nothing is charged for those).
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|no box or slot\|status=busy"; then
    echo "[retry $i] pod busy, sleeping 120 s"; sleep 120; continue
  fi
  break
done

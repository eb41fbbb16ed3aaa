#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy/transient
log=$1; to=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" "$log" && ! grep -q "status=ok" "$log"; then sleep 90; continue; fi
  break
done
tail -40 "$log"

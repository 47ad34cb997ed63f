#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <logfile> <command...>   -- retries while the pod answers "busy" (rc 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then echo "gpurun rc=$rc (attempt $i)"; tail -40 "$LOG"; exit $rc; fi
  sleep 120
done
echo "gave up after 40 attempts"; exit 3

#!/bin/bash
# resubmit a gpurun call until the pod takes it (exit code 3 / "transient" = nothing charged): tools/gpurun_retry.sh OUTFILE [gpurun args...]
out=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if ! grep -q "status=transient" "$out"; then exit 0; fi
  sleep 45
done

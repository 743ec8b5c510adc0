#!/bin/bash
# ncu evidence for round 1 (run under gpurun on ONE GPU; numbers printed under ncu are never bench values)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
# (1) launch list with device times (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
# (2) full capture of the dominant kernel and the pull-push kernel
ncu --set full --clock-control none --import-source on -k regex:final_gather -s 3 -c 1 -o gpurun_out/fg_${TAG} $CMD > gpurun_out/ncu_fg_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pullpush -s 3 -c 1 -o gpurun_out/pp_${TAG} $CMD > gpurun_out/ncu_pp_${TAG}.log 2>&1
ls -la gpurun_out

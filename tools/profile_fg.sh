#!/bin/bash
# ncu full capture of the gathering kernel only (1 GPU). usage: profile_fg.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:final_gather -s 3 -c 1 -o gpurun_out/fg_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fg_${TAG}.log 2>&1

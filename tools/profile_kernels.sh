#!/bin/bash
# usage: profile_kernels.sh TAG regex   -- ncu --set full on the kernels matching regex (first frame after warm-up)
TAG=$1; RE=$2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$RE -s ${3:-4} -c ${4:-2} -o gpurun_out/k_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k_${TAG}.log 2>&1

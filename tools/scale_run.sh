#!/bin/bash
# usage (under gpurun --gpus 8): tools/scale_run.sh TAG  -> gpurun_out/scale_TAG_*.json (4K/4096 at N=8,4,2 fused; 8K/16384 at N=8)
TAG=${1:-r1}
run() { # n workload exchange name
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $1 --workload $2 --exchange $3 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${TAG}_$4.log 2> gpurun_out/scale_${TAG}_$4.err
  grep "^{" gpurun_out/scale_${TAG}_$4.log | tail -1 > gpurun_out/scale_${TAG}_$4.json
  python -c "
import json; d=json.load(open('gpurun_out/scale_${TAG}_$4.json')); print('$4', d['n_gpus'], 'ms', round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['stage_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],3))"
}
run 8 4k_4096 fused n8_fused
run 4 4k_4096 fused n4_fused
run 2 4k_4096 fused n2_fused



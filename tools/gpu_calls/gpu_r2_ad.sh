#!/bin/bash
# round 2, GPU call AD (8 GPUs): end to end at 8 / 4 GPUs with and without binding each rank to its GPU's NUMA node
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi topo -m | grep -E "^GPU[0-7]" | awk '{print $1, $(NF-2), $(NF-1), $NF}' | head -8
nproc
run() {
  n=$1; label=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err
  python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_ad.json') if l.startswith('{')][-1]); print('n=$n $label: e2e ms', round(d['e2e']['ms_per_step'],4), 'frame', round(d['ms_per_step'],4), 'bound cores', d['e2e'].get('host_cores_bound_to_gpu_numa_node'), 'match', d['matches_single_gpu_image'])
except Exception as e: print('n=$n $label unreadable', e)
" | tee -a gpurun_out/ab_ad.log
}
run 8 "bind" X=1
run 8 "no bind" MLGI_BENCH_BIND=0
run 8 "bind" X=1
run 4 "bind" X=1
run 4 "no bind" MLGI_BENCH_BIND=0

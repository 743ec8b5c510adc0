#!/bin/bash
# round 2, GPU call T (8 GPUs): end-to-end A/B of the sharded host frame at 8 and 4 GPUs: graph replay on/off, phase counts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() {  # n, label, env...
  n=$1; label=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
  python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_t.json') if l.startswith('{')][-1]); print('n=$n $label: e2e ms', round(d['e2e']['ms_per_step'],4), 'frame', round(d['ms_per_step'],4), 'match', d['matches_single_gpu_image'])
except Exception as e: print('n=$n $label unreadable', e)
" | tee -a gpurun_out/ab_t.log
  grep "mlgi trace" gpurun_out/bench_t.err | tail -n 2 | cut -c1-400
}
run 8 "graph default(on)" X=1
run 8 "graph off" MLGI_HOST_GRAPH=0
run 8 "graph on, 2 phases" MLGI_HOST_SLABS=2
run 8 "graph off, trace" MLGI_HOST_GRAPH=0 MLGI_HOST_TRACE=1
run 4 "graph default(on)" X=1
run 4 "graph off" MLGI_HOST_GRAPH=0
run 4 "graph on, 4 phases" MLGI_HOST_SLABS=4
run 2 "graph default(on)" X=1

#!/bin/bash
# round 2, GPU call Y (1 GPU): slab-count sweep of the single-GPU host pipeline after sorting / compact RSM; timeline at the best counts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for k in 8 10 12 16; do
  MLGI_HOST_SLABS=$k python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('slabs $k: e2e ms', round(d['e2e']['ms_per_step'],4), 'frame', round(d['ms_per_step'],4))" | tee -a gpurun_out/ab_y.log
done
MLGI_HOST_TRACE=1 MLGI_HOST_SLABS=8 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scale-8k 2>&1 >/dev/null | grep "mlgi trace" | tail -n 2 | cut -c1-900 | tee -a gpurun_out/ab_y.log

#!/bin/bash
# round 2, GPU call Z2 (1 GPU): occupancy cap of the gathering in the host pipeline x slab plans; timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for occ in 8 7 10; do for k in shaped 8 10; do
  if [ $k = shaped ]; then unset MLGI_HOST_SLABS; else export MLGI_HOST_SLABS=$k; fi
  MLGI_FG_HOST_OCC=$occ python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('occ $occ slabs $k: e2e ms', round(d['e2e']['ms_per_step'],4), 'frame', round(d['ms_per_step'],4))" | tee -a gpurun_out/ab_z2.log
done; done
unset MLGI_HOST_SLABS
MLGI_FG_HOST_OCC=8 MLGI_HOST_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scale-8k 2>&1 >/dev/null | grep "mlgi trace" | tail -n 1 | cut -c1-900 | tee -a gpurun_out/ab_z2.log
MLGI_HOST_SLABS=8 MLGI_FG_HOST_OCC=8 MLGI_HOST_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scale-8k 2>&1 >/dev/null | grep "mlgi trace" | tail -n 1 | cut -c1-900 | tee -a gpurun_out/ab_z2.log

#!/bin/bash
# round 2, GPU call Z (1 GPU): shaped slab plan of the host pipeline: parity (host-frame tests) + end to end vs uniform slabs + timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "host or frame or packed or full_size" > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu_z.log
for k in shaped 8 10 shaped 8 10; do
  if [ $k = shaped ]; then unset MLGI_HOST_SLABS; else export MLGI_HOST_SLABS=$k; fi
  python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('slabs $k: e2e ms', round(d['e2e']['ms_per_step'],4), 'frame', round(d['ms_per_step'],4))" | tee -a gpurun_out/ab_z.log
done
unset MLGI_HOST_SLABS
MLGI_HOST_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scale-8k 2>&1 >/dev/null | grep "mlgi trace" | tail -n 2 | cut -c1-900 | tee -a gpurun_out/ab_z.log

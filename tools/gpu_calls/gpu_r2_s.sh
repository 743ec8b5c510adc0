#!/bin/bash
# round 2, GPU call S (1 GPU): parity with the host-frame graphs; end-to-end A/B (graph on/off, slab counts); pipeline option test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu_s.log
for hg in 1 0; do for k in 4 6 8; do
  MLGI_HOST_GRAPH=$hg MLGI_HOST_SLABS=$k python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('host graph $hg slabs $k: e2e ms', round(d['e2e']['ms_per_step'],4), 'frame', round(d['ms_per_step'],4))" | tee -a gpurun_out/ab_s.log
done; done

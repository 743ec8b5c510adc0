#!/bin/bash
# round 2, GPU call AF (1 GPU): culling started after the splat (MLGI_CULL_LATE) vs beside point generation
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in 0 1 0 1; do
  MLGI_CULL_LATE=$v python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cull_late=$v', 'frame %.4f ms' % d['ms_per_step'], d['image_sha256'][:12])" | tee -a gpurun_out/ab_af.log
done
MLGI_CULL_LATE=1 python -m pytest tests -m gpu -x -q -k "frame or graph" 2>&1 | tail -n 2

#!/bin/bash
# round 2, GPU call F: blur variants, stream-priority A/B (full frame and 1/8 shard), blur parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "blur or composite or frame" > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_f.log
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_f.log
for pr in 1 0; do
  MLGI_PRIO=$pr python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('MLGI_PRIO=$pr', 'frame %.4f ms' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
  MLGI_PRIO=$pr python tools/shard_sim.py --world 8 --rank 0
  MLGI_PRIO=$pr python tools/shard_sim.py --world 2 --rank 0
done 2>&1 | tee gpurun_out/prio_f.log

#!/bin/bash
# round 2, GPU call V (1 GPU): parity after the compact RSM upload (host frames)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu_v.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e ms', round(d['e2e']['ms_per_step'],4), 'h2d', d['e2e']['h2d_bytes_per_step'], 'frame', round(d['ms_per_step'],4))"

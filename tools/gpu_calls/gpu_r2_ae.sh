#!/bin/bash
# round 2, GPU call AE (1 GPU): final parity run after the host-pointer guard
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ae.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu_ae.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('frame', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"

#!/bin/bash
# round 2, GPU call AB (1 GPU): pull-push after the unchecked exact divisions + warp-aggregated list append: parity, timing, instruction count
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_ab.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('frame %.4f ms' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()}, d['image_sha256'][:12])"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:pullpush -s 4 -c 2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | grep -E "gpu__time|smsp__inst" | tail -n 4

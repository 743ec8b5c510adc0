#!/bin/bash
# round 2, GPU call AC (1 GPU): gathering occupancy variants (9 / 10 / 11 / 12 CTAs per SM) on the round-2 kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_ac.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default (10)', 'frame %.3f ms' % d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms'].items()})" | tee -a gpurun_out/ab_ac.log

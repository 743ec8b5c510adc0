#!/bin/bash
# round 2, GPU call E: A/B of the bulk-async variants (FG record staging, pull-push TMA tile copies), parity suite, memcheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu_e.log
for cfg in "MLGI_FG_BULK=1 MLGI_PP_TMA=1" "MLGI_FG_BULK=0 MLGI_PP_TMA=1" "MLGI_FG_BULK=1 MLGI_PP_TMA=0" "MLGI_FG_BULK=0 MLGI_PP_TMA=0"; do
  for rep in 1 2; do
  env $cfg python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$cfg', 'frame %.4f ms' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
  done
done 2>&1 | tee gpurun_out/ab_bulk_e.log
for w in 8; do MLGI_FG_BULK=1 python tools/shard_sim.py --world $w --rank 0; MLGI_FG_BULK=0 python tools/shard_sim.py --world $w --rank 0; done 2>&1 | tee gpurun_out/shard_sim_e.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/sanitizer_memcheck.full.log 2>&1; grep -E "ERROR SUMMARY|sanitize run done" gpurun_out/sanitizer_memcheck.full.log

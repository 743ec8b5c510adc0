#!/bin/bash
# round 2, GPU call J: two-ranks-on-one-GPU test vs hardware queue count; parity after the cull split / strided slabs; e2e timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for c in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k two_sequential_shards > gpurun_out/pytest_j_conn$c.log 2>&1; echo "connections=$c rc=$?"; tail -n 2 gpurun_out/pytest_j_conn$c.log
done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu_j.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default lib', 'frame %.4f ms' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()}, d['image_sha256'][:12])" | tee gpurun_out/ab_j.log
for k in 4 6 8 12; do
  MLGI_HOST_TRACE=1 MLGI_HOST_SLABS=$k python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_j_slabs$k.json 2> gpurun_out/trace_j_slabs$k.log
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_j_slabs$k.json') if l.startswith('{')][-1]); print('slabs $k e2e ms', round(d['e2e']['ms_per_step'],4))" | tee -a gpurun_out/ab_j.log
  tail -n 1 gpurun_out/trace_j_slabs$k.log
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
ncu --set full --clock-control none --import-source on -k regex:"tile_bounds|tile_lists" -c 2 -o gpurun_out/r2j_cull $CMD > gpurun_out/ncu_r2j_cull.log 2>&1; echo "ncu cull rc=$?"

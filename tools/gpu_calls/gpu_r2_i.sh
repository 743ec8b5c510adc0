#!/bin/bash
# round 2, GPU call I: parity after the three-way frame fork / new blur / host pipeline streams; e2e timeline per slab count;
# ncu of the culling and blur kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu_i.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default lib', 'frame %.4f ms' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()}, d['image_sha256'][:12])" | tee gpurun_out/ab_i.log
for k in 6 8 10 12 16; do
  MLGI_HOST_TRACE=1 MLGI_HOST_SLABS=$k python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_i_slabs$k.json 2> gpurun_out/trace_i_slabs$k.log
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_i_slabs$k.json') if l.startswith('{')][-1]); print('slabs $k e2e ms', round(d['e2e']['ms_per_step'],4))" | tee -a gpurun_out/ab_i.log
  tail -n 1 gpurun_out/trace_i_slabs$k.log
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
for k in cull_tiled gi_blur; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/r2i_$k $CMD > gpurun_out/ncu_r2i_$k.log 2>&1; echo "ncu $k rc=$?"
done

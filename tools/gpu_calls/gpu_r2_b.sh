#!/bin/bash
# round 2, GPU call B: parity suite + quick bench + ncu of pull-push (TAG = $1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
TAG=${1:-b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -n 12 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
ncu --set full --clock-control none --import-source on -k regex:pullpush -s 3 -c 1 -o gpurun_out/pp_$TAG $CMD > gpurun_out/ncu_pp_$TAG.log 2>&1; echo "ncu rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'stages', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'enq', d.get('host_enqueue_us_per_frame'))
print('roofline', d['roofline']['frac'], d['roofline_hbm']['frac'])
"
tail -n 5 gpurun_out/bench_$TAG.err

#!/bin/bash
# round 2, GPU call D: parity suite (n3, blur, exchange), bench lines (default, clustered, anim64), compute-sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu_d.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --cull clustered --no-cpu-baseline --no-scale-8k > gpurun_out/bench_d_clustered.json 2> gpurun_out/bench_d_clustered.err; echo "bench clustered rc=$?"
python bench.py --steps 64 --warmup 64 --workload anim64 --no-cpu-baseline --no-scale-8k --no-e2e > gpurun_out/bench_d_anim64.json 2> gpurun_out/bench_d_anim64.err; echo "bench anim rc=$?"
python -c "
import json
for f in ('bench_d','bench_d_clustered','bench_d_anim64'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, 'ms', round(d['ms_per_step'],4), 'stages', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'enq', round(d.get('host_enqueue_us_per_frame'),2), 'graph', d.get('graph'), 'blur', (d.get('roofline_blur') or {}).get('frac'))
    except Exception as e:
        print(f, 'unreadable', e)
"
tail -n 3 gpurun_out/bench_d*.err
bash tools/run_sanitizer.sh

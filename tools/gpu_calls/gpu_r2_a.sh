#!/bin/bash
# round 2, GPU call A: parity suite, bench line, A/B of graph replay vs direct launches, ncu of the new pull-push kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
MLGI_GRAPH=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k > gpurun_out/bench_a_nograph.json 2> gpurun_out/bench_a_nograph.err; echo "bench nograph rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
ncu --set full --clock-control none --import-source on -k regex:pullpush -s 3 -c 1 -o gpurun_out/pp_r2a $CMD > gpurun_out/ncu_pp_r2a.log 2>&1; echo "ncu rc=$?"
python -c "
import json
for f in ('bench_a','bench_a_nograph'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', d['ms_per_step'], 'stages', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'enq', d.get('host_enqueue_us_per_frame'), 'clk', d.get('clocks'), '8k', d.get('scale_8k'))
    except Exception as e:
        print(f, 'unreadable', e)
"
tail -n 5 gpurun_out/bench_a.err

#!/bin/bash
# round 2, GPU call O (1 GPU): full GPU parity suite (phased sharded host frame included), pipe micro-benchmark, ISM-fetch A/B,
# bench records for the clustered and animated configurations, ncu captures of blur / pull-push / gathering
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu_o.log
timeout 120 python tools/debug_sharded_host.py 2>&1 | tail -n 14
./tools/ubench/pipes | tee gpurun_out/ubench_pipes.txt
python tools/texf_study.py 4k_4096 texi 2>&1 | tail -n 3 | tee gpurun_out/texi_study.txt
bash tools/ab_env.sh MLGI_ISM_FETCH tex texi none tex texi none | tee gpurun_out/ab_o.log
python bench.py --cull clustered --steps 50 --warmup 5 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_o_clustered.json 2>/dev/null; echo "clustered rc=$?"
python bench.py --workload anim64 --steps 64 --warmup 5 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_o_anim64.json 2>/dev/null; echo "anim64 rc=$?"
python -c "
import json
for n in ('clustered','anim64'):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_o_%s.json'%n) if l.startswith('{')][-1]); print(n, 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), d.get('unit'), (d.get('e2e') or {}).get('ms_per_step'))
    except Exception as e: print(n, 'unreadable', e)
"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
for k in gi_blur pullpush final_gather; do
  skip=4; [ $k = gi_blur ] && skip=0
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2o_$k $CMD > gpurun_out/ncu_r2o_$k.log 2>&1; echo "ncu $k rc=$?"
done

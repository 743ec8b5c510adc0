#!/bin/bash
# round 2, GPU call C: pull-push variants (register cap), shard simulation, parity suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pullpush or full_size or frame" > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_c.log
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_c.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default lib', 'frame %.3f ms' % d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms'].items()})" | tee -a gpurun_out/ab_c.log
for r in 0 3 7; do python tools/shard_sim.py --world 8 --rank $r; done 2>&1 | tee gpurun_out/shard_sim_c.log
python tools/shard_sim.py --world 2 --rank 0 2>&1 | tee -a gpurun_out/shard_sim_c.log
python tools/shard_sim.py --world 8 --rank 0 --workload 8k_16384 --steps 10 2>&1 | tee -a gpurun_out/shard_sim_c.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
ncu --set full --clock-control none --import-source on -k regex:pullpush -s 3 -c 1 -o gpurun_out/pp_c $CMD > gpurun_out/ncu_pp_c.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# round 2, GPU call H: parity after the cull phase-1 / gathering tweaks, A/B against the previous library, FFMA2 issue micro-benchmark,
# timeline of the host-input frame, ncu of the blur kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_h.log
tools/ubench/ffma2 2>&1 | tee gpurun_out/ubench_ffma2.log
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_h.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default lib', 'frame %.4f ms' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()}, d['image_sha256'][:12])" | tee -a gpurun_out/ab_h.log
MLGI_HOST_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_h_trace.json 2> gpurun_out/trace_h.log; tail -n 4 gpurun_out/trace_h.log
ncu --set full --clock-control none --import-source on -k regex:gi_blur -c 1 -o gpurun_out/r2_gi_blur python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k > gpurun_out/ncu_r2_gi_blur.log 2>&1; echo "ncu blur rc=$?"

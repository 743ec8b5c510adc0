#!/bin/bash
# round 2, GPU call R (1 GPU): the committed state -- full GPU parity suite, smoke, default bench line + reference arm, launch list,
# ncu summaries of the three main kernels, compute-sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu_r.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py > gpurun_out/bench_r_n1.json 2> gpurun_out/bench_r_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r_n1.json') if l.startswith('{')][-1])
print('frame', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'stages', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'fg frac', round(d['roofline']['frac'],4), 'pp frac', round(d['roofline_hbm']['frac'],4), 'blur frac', round(d['roofline_blur']['frac'],4), '8k', d['scale_8k']['ms_per_frame'], d['image_sha256'][:12], 'clk', d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r_ref.json 2> gpurun_out/bench_r_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-scale-8k > gpurun_out/ncu_r.log 2>&1; echo "ncu list rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
for k in gi_blur pullpush final_gather; do
  skip=4; [ $k = gi_blur ] && skip=0
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2r_$k $CMD > gpurun_out/ncu_r2r_$k.log 2>&1; echo "ncu $k rc=$?"
done
bash tools/run_sanitizer.sh > gpurun_out/sanitizer_r.log 2>&1; echo "sanitizer rc=$?"; tail -n 12 gpurun_out/sanitizer_r.log

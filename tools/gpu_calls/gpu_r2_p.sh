#!/bin/bash
# round 2, GPU call P (N GPUs, default 8): multi-GPU parity tests (first 2 GPUs), full bench lines at N, N/2, 2 (driver-style launch), timeline of the host frame
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
NG=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi_p.log 2>&1; echo "pytest multi rc=$?"; tail -n 4 gpurun_out/pytest_multi_p.log
for n in $NG $((NG/2)) 2; do
  [ $n -lt 2 ] && continue
  MLGI_HOST_TRACE=${TRACE:-0} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_p_n$n.json 2> gpurun_out/bench_p_n$n.err; echo "bench n=$n rc=$?"
  python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_p_n$n.json') if l.startswith('{')][-1])
    print('n=$n', 'ms', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'stages', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'e2e', round((d.get('e2e') or {}).get('ms_per_step',0),4), 'match', d['matches_single_gpu_image'], d['image_sha256'][:12], '8k', (d.get('scale_8k') or {}).get('ms_per_frame'), (d.get('scale_8k') or {}).get('matches_single_gpu_image'), 'enq', d.get('host_enqueue_us_per_frame'), 'clk', d['clocks']['samples'], 'fgfrac', round(d['roofline']['frac'],3))
except Exception as e:
    print('n=$n unreadable', e)
"
  tail -n 2 gpurun_out/bench_p_n$n.err | cut -c1-300
done

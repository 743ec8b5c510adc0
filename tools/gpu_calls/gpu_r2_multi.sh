#!/bin/bash
# round 2, multi-GPU call (N = number of GPUs of the box): multi-GPU parity tests + bench at N with the three exchange modes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 15 gpurun_out/pytest_multi.log
for x in pull push nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --exchange $x > gpurun_out/bench_n${N}_$x.json 2> gpurun_out/bench_n${N}_$x.err; echo "bench $x rc=$?"
  python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n${N}_$x.json') if l.startswith('{')][-1])
    print('$x', 'ms', round(d['ms_per_step'],4), 'stages', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'match', d['matches_single_gpu_image'], d['image_sha256'][:12], '8k', (d.get('scale_8k') or {}).get('ms_per_frame'), 'enq', d.get('host_enqueue_us_per_frame'))
except Exception as e:
    print('$x unreadable', e)
"
  tail -n 3 gpurun_out/bench_n${N}_$x.err
done

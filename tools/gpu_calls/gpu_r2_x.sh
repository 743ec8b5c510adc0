#!/bin/bash
# round 2, GPU call X (2 GPUs): multi-GPU parity tests with the final defaults (push exchange), recorded bench line at 2 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "push-True or host" > gpurun_out/pytest_multi_x.log 2>&1; echo "pytest multi rc=$?"; tail -n 4 gpurun_out/pytest_multi_x.log
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2991$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_x_n$n.json 2> gpurun_out/bench_x_n$n.err; echo "bench n=$n rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_x_n$n.json') if l.startswith('{')][-1])
print('n=$n', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'match', d['matches_single_gpu_image'], d['image_sha256'][:12], '8k', d['scale_8k']['ms_per_frame'], d['config']['parallelism'])"

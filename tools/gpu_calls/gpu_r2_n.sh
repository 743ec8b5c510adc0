#!/bin/bash
# round 2, GPU call N (2 GPUs): timeline of mlgi_frame_sharded_host at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
MLGI_HOST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-scale-8k > gpurun_out/bench_n_n2.json 2> gpurun_out/bench_n_n2.err; echo "bench rc=$?"
grep "mlgi trace" gpurun_out/bench_n_n2.err | tail -n 12
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n_n2.json') if l.startswith('{')][-1]); print('e2e ms', d['e2e']['ms_per_step'], 'frame', d['ms_per_step'])"
nvidia-smi topo -m | head -8

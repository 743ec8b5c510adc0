#!/bin/bash
# round 2, GPU call W (8 GPUs): the recorded bench lines at 8 / 4 / 2 GPUs with the library defaults (driver-style launch), pull vs push at 2 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2981$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_w_n$n.json 2> gpurun_out/bench_w_n$n.err; echo "bench n=$n rc=$?"
  python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_w_n$n.json') if l.startswith('{')][-1])
    print('n=$n', 'ms', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round((d.get('e2e') or {}).get('ms_per_step',0),4), 'match', d['matches_single_gpu_image'], d['image_sha256'][:12], '8k', (d.get('scale_8k') or {}).get('ms_per_frame'), (d.get('scale_8k') or {}).get('matches_single_gpu_image'), d['config']['parallelism'], 'fgfrac', round(d['roofline']['frac'],3))
except Exception as e:
    print('n=$n unreadable', e)
"
  tail -n 2 gpurun_out/bench_w_n$n.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29722 tools/sweep_host_frame.py 4::pull 4::push 2>/dev/null | grep "^N=" | tee gpurun_out/sweep_w.log

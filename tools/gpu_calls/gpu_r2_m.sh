#!/bin/bash
# round 2, GPU call M (2 GPUs): multi-GPU parity tests, debug of the one-GPU two-rank host frame, bench at N=2 (default exchange)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python tools/debug_sharded_host.py > gpurun_out/debug_sharded_host.log 2>&1; echo "debug rc=$?"; tail -n 30 gpurun_out/debug_sharded_host.log
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi_m.log 2>&1; echo "pytest multi rc=$?"; tail -n 15 gpurun_out/pytest_multi_m.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_m_n2.json 2> gpurun_out/bench_m_n2.err; echo "bench rc=$?"
tail -c 5000 gpurun_out/bench_m_n2.json; tail -n 5 gpurun_out/bench_m_n2.err

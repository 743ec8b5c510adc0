#!/bin/bash
# round 2, GPU call U (8 GPUs): multi-GPU parity after the compact RSM upload; sweep of the sharded host frame's phase count at 8 / 4 / 2 GPUs;
# pull vs push exchange on the device-resident frame
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -k "host or pull-True" > gpurun_out/pytest_multi_u.log 2>&1; echo "pytest multi rc=$?"; tail -n 3 gpurun_out/pytest_multi_u.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 8 --master-port 29718 tools/sweep_host_frame.py 1 2 3 4 2:0 1::push 2>/dev/null | grep "^N=" | tee gpurun_out/sweep_u.log
timeout 300 $T --nproc-per-node 4 --master-port 29714 tools/sweep_host_frame.py 1 2 3 4 1::push 2>/dev/null | grep "^N=" | tee -a gpurun_out/sweep_u.log
timeout 300 $T --nproc-per-node 2 --master-port 29712 tools/sweep_host_frame.py 2 4 6 8 2>/dev/null | grep "^N=" | tee -a gpurun_out/sweep_u.log

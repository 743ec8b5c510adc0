#!/bin/bash
# round 2, GPU call AA (1 GPU): the one-GPU multi-rank tests with four shards (exchange, host frame), repeated for flakiness
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for i in 1 2 3; do
  python -m pytest tests -m gpu -q -k "sequential_shards or sharded_host" > gpurun_out/pytest_gpu_aa$i.log 2>&1; echo "run $i rc=$?"; tail -n 3 gpurun_out/pytest_gpu_aa$i.log
done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_aa.log 2>&1; echo "full rc=$?"; tail -n 3 gpurun_out/pytest_gpu_aa.log

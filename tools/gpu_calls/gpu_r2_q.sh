#!/bin/bash
# round 2, GPU call Q (1 GPU): parity after the barrier-free record ring + the pull-push / gathering pipeline; A/B of the ring depth
# (2 / 3 / 4 buffers) and of the number of pipeline groups
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu_q.log
bash tools/ab_env.sh MLGI_PP_PARTS 1 2 4 8 1 4 2>&1 | tee gpurun_out/ab_q.log
MLGI_PP_PARTS=1 bash tools/ab.sh 2>&1 | tee -a gpurun_out/ab_q.log
bash tools/ab.sh 2>&1 | tee -a gpurun_out/ab_q.log

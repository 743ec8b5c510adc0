#!/bin/bash
# round 2, GPU call O: full GPU parity suite after the test fixes; pipe micro-benchmark
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu_o.log
./tools/ubench/pipes | tee gpurun_out/ubench_pipes.txt
python tools/texf_study.py 4k_4096 texi 2>&1 | tail -n 3 | tee gpurun_out/texi_study.txt
bash tools/ab_env.sh MLGI_ISM_FETCH tex texi none tex texi none | tee gpurun_out/ab_o.log

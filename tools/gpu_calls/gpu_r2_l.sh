#!/bin/bash
# round 2, GPU call L: full GPU parity suite (no -x), sanitizer logs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_gpu_l.log
cat gpurun_out/parity_stats.json 2>/dev/null | head -c 1500
bash tools/run_sanitizer.sh > gpurun_out/sanitizer_l.log 2>&1; echo "sanitizer rc=$?"; tail -n 20 gpurun_out/sanitizer_l.log

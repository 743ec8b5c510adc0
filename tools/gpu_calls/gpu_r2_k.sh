#!/bin/bash
# round 2, GPU call K (re-entry): full GPU parity suite, smoke, the default bench line, the reference arm, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu_k.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
python bench.py > gpurun_out/bench_k_n1.json 2> gpurun_out/bench_k_n1.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_k_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_k_ref.json 2> gpurun_out/bench_k_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_k_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-scale-8k > gpurun_out/ncu_k.log 2>&1; echo "ncu rc=$?"

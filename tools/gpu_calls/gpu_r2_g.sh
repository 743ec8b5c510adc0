#!/bin/bash
# round 2, GPU call G: full parity suite, default bench line, ncu launch list + --set full captures of every frame kernel (r2 evidence)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_g.log
python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scale-8k"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu list rc=$?"
for k in final_gather pullpush cull_tiled gi_blur splat_kernel point_generate; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/r2_$k $CMD > gpurun_out/ncu_r2_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out | tail -n 12

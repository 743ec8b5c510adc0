#!/bin/bash
# A/B timing of CUDA library variants in many-lights-gi_b200/lib/variants/*.so (FG + frame ms)
for so in many-lights-gi_b200/lib/variants/*.so; do
  MLGI_LIB_PATH=$PWD/$so python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$so', 'frame %.3f ms' % d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms'].items()})"
done

#!/bin/bash
# A/B timing of CUDA library variants in many-lights-gi_b200/lib/variants/*.so (FG + frame ms); pass "test" to also run the radiance tests per variant
for so in many-lights-gi_b200/lib/variants/*.so; do
  MLGI_LIB_PATH=$PWD/$so python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$so', 'frame %.3f ms' % d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms'].items()})"
  if [ "$1" = "test" ]; then MLGI_LIB_PATH=$PWD/$so python -m pytest tests -m gpu -x -q -k "radiance or frame" 2>&1 | tail -1; fi
done

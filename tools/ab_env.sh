#!/bin/bash
# A/B timing of one environment switch: tools/ab_env.sh VAR v1 v2 ...   (FG + frame ms per value)
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-scale-8k 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$var=$v', 'frame %.3f ms' % d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms'].items()})"
done

#!/bin/bash
# A/B on the GPU box by environment: bash tools/ab_env.sh name1 "ENV=.." name2 "ENV=.." ...
while [ $# -ge 2 ]; do
  env $2 python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), {k: round(v,3) for k,v in d['stages_ms_per_step'].items()})"
  shift 2
done

#!/bin/bash
# e2e sweep by environment on the GPU box: bash tools/ab_e2e.sh "ENV=.. ENV2=.." ...
for e in "$@"; do
  env $e python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$e', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['runs_ms_per_step'])"
done

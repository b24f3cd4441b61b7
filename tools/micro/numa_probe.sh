#!/bin/bash
nvidia-smi topo -m 2>/dev/null | head -20
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" | head -12
python - <<'PY'
import os
print("affinity", sorted(os.sched_getaffinity(0))[:8], "... n =", len(os.sched_getaffinity(0)))
PY
for i in 1 2 3; do
python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e', round(e['value']), 'ceiling', round(e['copy_ceiling']['value']), 'h2d GB/s', round(e['copy_ceiling']['h2d_GBs_aggregate'],1), e['runs_ms_per_step'])"
done

#!/bin/bash
# bench.py --no-extras at N ranks with an environment: bash tools/scale_env.sh <N> "ENV=.."
N=$1; shift
for e in "$@"; do
env $e python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port $((29800 + N)) bench.py --gpus $N --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$e', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['copy_ceiling']['value']))"
done

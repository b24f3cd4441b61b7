#!/bin/bash
# bench.py at N ranks only (run under `gpurun --gpus N`): bash tools/scale_one.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port $((29700 + N)) bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
echo "bench N=$N rc=$?"
python - gpurun_out/bench_n${N}_$TAG.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value %.0f e2e %.0f ceiling %.0f" % (d["value"], d["e2e"]["value"], d["e2e"]["copy_ceiling"]["value"]))
    print(json.dumps(d.get("matching_sharded"))[:600])
except Exception as e:
    print("parse failed", e)
PY

#!/bin/bash
# One multi-GPU session on the box (run under `gpurun --gpus N`): the host->device ceiling at 1..N ranks, then bench.py at N.
# Usage: bash tools/scale_round.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  for flags in "" "--d2h" "--wc --d2h"; do
    $TR --nproc-per-node $n --master-port $((29600 + n)) tools/micro/h2d_all.py $flags 2>/dev/null | tail -1
  done
done
for n in 2 4 8; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port $((29700 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_$TAG.json 2> gpurun_out/bench_n${n}_$TAG.err
  echo "bench N=$n rc=$?"
  python - gpurun_out/bench_n${n}_$TAG.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]))
    print(json.dumps(d.get("matching_sharded"), indent=1))
except Exception as e:
    print("parse failed", e)
PY
done

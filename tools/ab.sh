#!/bin/bash
# A/B on the GPU box: bench.py --no-extras with the in-tree library and with every gpurun_variants/libvsg_*.so given.
# Usage (under gpurun): bash tools/ab.sh tag [variant ...]
TAG=$1; shift
mkdir -p gpurun_out
run() { # name libpath
  VSG_LIB_PATH=$2 python bench.py --no-extras --steps 10 --warmup 3 > gpurun_out/ab_${TAG}_$1.json 2> gpurun_out/ab_${TAG}_$1.err
  python - "$1" gpurun_out/ab_${TAG}_$1.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print("%-10s value %.0f  e2e %.0f  stages %s" % (sys.argv[1], d["value"], d["e2e"]["value"], {k: round(v,3) for k,v in d["stages_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base ""
for v in "$@"; do run $v $PWD/gpurun_variants/libvsg_$v.so; done

#!/bin/bash
# One GPU-box session: smoke, bench, ncu launch list and a full capture of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'fast_blur_kernel|fast_kernel|blur_kernel|blur_tc_kernel|octree_kernel|resize_kernel|resize_tma_kernel|describe_kernel|slot_kernel' \
    -s 33 -c 11 -o $OUT/prof_$TAG python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT

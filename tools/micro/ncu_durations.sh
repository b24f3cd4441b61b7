#!/bin/bash
# kernel durations of one command under ncu (cold-cache, serialised): bash tools/micro/ncu_durations.sh <skip> <count> <cmd...>
S=$1; C=$2; shift 2
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c $C --csv --log-file gpurun_out/_dur.csv "$@" > /dev/null 2>&1
python - <<'PY'
import csv
rows = [l for l in open("gpurun_out/_dur.csv") if not l.startswith("==")]
for r in csv.DictReader(rows):
    print("%-40s %10s %s" % (r["Kernel Name"][:40], r["Metric Value"], r["Metric Unit"]))
PY

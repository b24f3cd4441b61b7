#!/usr/bin/env python3
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export: share of executed warp
instructions and of stall samples per CUDA source line (the profile behind DESIGN.md's instruction budgets).
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv;  python tools/ncu_lines.py src.csv [file.cu] [min_pct]
"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    cur, agg = None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if len(r) > 8 and r[0].isdigit():
            try:
                k = r.index("-")
            except ValueError:
                continue
            if r[k + 1] != "-":
                continue
            vals = r[k + 2:]
            try:
                inst, stall = int(vals[3]), int(vals[0])
            except ValueError:
                continue
            agg.append((cur.split("/")[-1], int(r[0]), ",".join(r[1:k]), inst, stall))
    return agg


def main():
    agg = load(sys.argv[1])
    only = sys.argv[2] if len(sys.argv) > 2 else None
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
    tot = sum(a[3] for a in agg)
    stot = sum(a[4] for a in agg) or 1
    print("total warp instructions %d, stall samples %d" % (tot, stot))
    byfile = {}
    for f, _, _, i, st in agg:
        byfile.setdefault(f, [0, 0])
        byfile[f][0] += i
        byfile[f][1] += st
    for f, v in byfile.items():
        print("  %-32s %6.2f%% inst %6.2f%% stall" % (f, 100 * v[0] / tot, 100 * v[1] / stot))
    for f, l, s, i, st in sorted(agg, key=lambda a: (a[0], a[1])):
        if (only is None or f == only) and 100 * i / tot >= min_pct:
            print("%-18s %4d %6.2f%% inst %6.2f%% stall  %s" % (f, l, 100 * i / tot, 100 * st / stot, s[:110]))


if __name__ == "__main__":
    main()

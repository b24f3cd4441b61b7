#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/summarize_ncu.py <tag>
reads  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_<tag>.ncu-rep   (ncu --set full capture of the top kernels), if present
writes profiles/<tag>_launches.md, profiles/<tag>_kernels.csv
"""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

launch_csv = os.path.join(root, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(launch_csv):
    lines = [l for l in open(launch_csv) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % tag)
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py "
                "--steps 2 --warmup 1 --no-extras` (512 frames per device-resident launch, 128 per launch in the end-to-end part; per-launch times are cold-cache and serialised: "
                "compare SHARES).\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, n, us, 100 * us / total))
    print("wrote launches summary,", len(rows), "rows")

rep = os.path.join(root, "gpurun_out", "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
            "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(out_dir, "%s_kernels.csv" % tag), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([w for w, _ in idx])
        wr.writerow([units[i] for _, i in idx])
        for r in rows[2:]:
            wr.writerow([r[i].split("(")[0] if w == "Kernel Name" else r[i] for w, i in idx])
    print("wrote kernel metrics,", len(rows) - 2, "kernels")

    # DRAM traffic per launch of every captured kernel (dram__bytes_read.sum + dram__bytes_write.sum of its largest launch):
    # bench.py's roofline.traffic reads this file; the commit it was measured on is stamped into it.
    import json
    try:
        head = subprocess.run(["git", "-C", root, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:  # noqa: BLE001
        head = "unknown"
    col = {w: hdr.index(w) for w in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum") if w in hdr}
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {"source": "ncu --set full --clock-control none (tools/gpu_round.sh %s): dram__bytes_read.sum + dram__bytes_write.sum of the "
                         "kernel's longest captured launch" % tag,
               "head": head, "frames_per_launch": int(os.environ.get("VSG_PROFILE_BATCH", "512"))}
    best = {}
    extra_cols = {"issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                  "alu_pipe_active_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
                  "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                  "warp_instructions": "smsp__inst_executed.sum"}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("vsg::", "")
        rd = float(r[col["dram__bytes_read.sum"]]) * unit_scale.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr_ = float(r[col["dram__bytes_write.sum"]]) * unit_scale.get(units[col["dram__bytes_write.sum"]], 1.0)
        dur = float(r[col["gpu__time_duration.sum"]])
        if name not in best or dur > best[name][0]:
            ex = {}
            for k, c in extra_cols.items():
                hit = [i for i, hname in enumerate(hdr) if hname == c or hname.endswith("." + c)]
                if hit and r[hit[0]] not in ("", "n/a"):
                    ex[k] = round(float(r[hit[0]]), 2)
            best[name] = (dur, rd + wr_, ex)
    for name, (dur, b, ex) in best.items():
        traffic[name] = dict({"dram_bytes_per_launch": int(b)}, **ex)
    json.dump(traffic, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
    print("wrote ncu_traffic.json for", sorted(best))

#!/usr/bin/env python
"""Turns the raw ncu outputs a gpurun call left in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_profiles.py --round r01 --launches gpurun_out/launches_r01.csv --report gpurun_out/trace_r01.ncu-rep

* profiles/<round>_launches.csv          the launch list itself (ncu --metrics gpu__time_duration.sum)
* profiles/<round>_launch_shares.md      per-kernel share of the profiled region (cold-cache, serialised:
                                         compare SHARES with bench.py's kernel_ms_per_step, not absolutes)
* profiles/<round>_trace_closest.md      key metrics of the `ncu --set full` capture of the traversal kernel
* profiles/trace_closest_traffic.json    dram bytes per launch -> bench.py's roofline.traffic
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--round", default="r01")
ap.add_argument("--launches")
ap.add_argument("--report")
ap.add_argument("--command", default="")
a = ap.parse_args()
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("gtlc::", "")
    if "trace_kernel" in name:
        mode = {"0": "CLOSEST", "1": "ANY", "2": "SHADOW"}
        import re
        m = re.search(r"trace_kernel(_batched|_v2|_tma)?<(?:\(int\))?(\d)", name)
        if m:
            n = "trace_kernel%s<%s>" % (m.group(1) or "", mode.get(m.group(2), m.group(2)))
        if "(bool)1" in name or re.search(r"trace_kernel<\d, 1>", name):
            n += "+count"
    return n


if a.launches:
    rows = []
    with open(a.launches) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) * (1e-3 if r["Metric Unit"] == "ns" else 1.0)))
    shutil.copyfile(a.launches, os.path.join(out_dir, "%s_launches.csv" % a.round))
    agg = OrderedDict()
    for n, us in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + us)
    total = sum(t for _, t in agg.values())
    with open(os.path.join(out_dir, "%s_launch_shares.md" % a.round), "w") as f:
        f.write("# %s: launch list of `%s`\n\n" % (a.round, a.command or "python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline"))
        f.write("Source: `profiles/%s_launches.csv` (ncu --metrics gpu__time_duration.sum --clock-control none). Per-launch times under ncu are "
                "cold-cache and serialised: the SHARE per kernel is what must agree with bench.py's `kernel_ms_per_step`, not the absolute.\n\n" % a.round)
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n, c, t, 100.0 * t / total))
        f.write("| **all** | %d | %.1f | 100 %% |\n" % (len(rows), total))
    print("wrote launch shares:", {n: round(100 * t / total, 1) for n, (c, t) in agg.items()})

if a.report:
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.avg",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = []
    with open(os.path.join(out_dir, "%s_trace_closest.md" % a.round), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:trace_kernel` (%s)\n\n" % (a.round, os.path.basename(a.report)))
        f.write("One column per captured launch (consecutive bounces of one wave of config C2).\n\n| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        if ki is not None:
            f.write("| kernel | | " + " | ".join("`%s`" % short(r[ki]) for r in data) + " |\n")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in data)))
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in data:
            traffic.append(float(r[ir]) * to_bytes.get(units[ir], 1.0) + float(r[iw]) * to_bytes.get(units[iw], 1.0))
        f.write("\nDRAM traffic per launch (read + write): " + ", ".join("%.1f MB" % (t / 1e6) for t in traffic) + "\n")
    with open(os.path.join(out_dir, "trace_closest_traffic.json"), "w") as f:
        json.dump({"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "launches": traffic,
                   "mean": sum(traffic) / max(len(traffic), 1), "source": "profiles/%s_trace_closest.md" % a.round}, f)
    print("wrote trace summary; dram MB per launch:", [round(t / 1e6, 1) for t in traffic])

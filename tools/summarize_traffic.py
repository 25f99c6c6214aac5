#!/usr/bin/env python
"""Per-launch DRAM traffic of the per-bounce kernels over ONE whole wave (all eight bounces), from
`ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed`,
into profiles/<round>_traffic.json -- what bench.py reports as `roofline.traffic` next to the algorithmic bytes per launch.
The captured wave has the headline's wave size (1920x1080 x 8 samples = 16.6 M paths), so a launch here is a launch there.

  python tools/summarize_traffic.py --round r02 --config c3 --csv gpurun_out/r02_traffic_c3.csv
"""
import argparse
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--round", default="r02")
ap.add_argument("--config", required=True)
ap.add_argument("--csv", required=True)
a = ap.parse_args()
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def family(name):
    m = re.search(r"(trace_kernel\w*)<(?:\(int\))?(\d)", name)
    if m:
        return m.group(1) + {"0": "", "1": "_any", "2": "_shadow"}.get(m.group(2), "")
    return name.split("(")[0].replace("void ", "").replace("gtlc::", "").split("<")[0]


with open(a.csv) as f:
    lines = [ln for ln in f if ln.startswith('"')]
launches = {}
for r in csv.DictReader(lines):
    e = launches.setdefault(r["ID"], {"kernel": family(r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", ""))
    n, u = r["Metric Name"], r["Metric Unit"]
    if n.startswith("dram__bytes"):
        e["bytes"] = e.get("bytes", 0.0) + v * TO_BYTES.get(u, 1.0)
    elif n == "gpu__time_duration.sum":
        e["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
    elif n.startswith("gpu__dram_throughput"):
        e["pct"] = v
fam = {}
for e in launches.values():
    f = fam.setdefault(e["kernel"], {"bytes": [], "ms": [], "pct": []})
    f["bytes"].append(e.get("bytes", 0.0)); f["ms"].append(e.get("ms", 0.0)); f["pct"].append(e.get("pct", 0.0))
tp = os.path.join(ROOT, "profiles", "%s_traffic.json" % a.round)
traffic = {}
if os.path.exists(tp):
    with open(tp) as f:
        traffic = json.load(f)
cfg = {}
for k, v in fam.items():
    n = len(v["bytes"])
    cfg[k] = {"dram_bytes_per_launch": sum(v["bytes"]) / n, "dram_pct_of_peak": sum(v["pct"]) / n, "launches_captured": n,
              "ms_per_launch_under_ncu": sum(v["ms"]) / n, "dram_bytes_by_bounce": v["bytes"],
              "source": "profiles/%s_traffic_%s.csv (one 16.6 M-path wave, all bounces; ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum)" % (a.round, a.config)}
traffic[a.config] = cfg
with open(tp, "w") as f:
    json.dump(traffic, f, indent=1, sort_keys=True)
import shutil
shutil.copyfile(a.csv, os.path.join(ROOT, "profiles", "%s_traffic_%s.csv" % (a.round, a.config)))
print(a.config, {k: (round(v["dram_bytes_per_launch"] / 1e6, 1), v["launches_captured"]) for k, v in cfg.items()})

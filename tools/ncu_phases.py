#!/usr/bin/env python
"""Per-source-line aggregation of an `ncu --set full --import-source on` report (needs -lineinfo builds).

  python tools/ncu_phases.py REPORT.ncu-rep [--launch K] [--top N] [--file traverse.cuh] [--ranges name:file:lo-hi,...]

Prints, for the K-th profiled launch (default 0): warp instructions executed, average active lanes, the share of
stall samples per source line, and -- with --ranges -- the same aggregated over named line ranges (kernel phases).
"""
import argparse
import csv
import io
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--launch", type=int, default=0)
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--ranges", default="")
a = ap.parse_args()

txt = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the report concatenates launches: a new launch starts at every "Kernel Name"/"Function Name" block following a file list restart
launch = -1
cur = None
hdr = None
agg = []
seen_files = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        if f in seen_files or launch < 0:
            if f in seen_files or launch < 0:
                launch += 1
                seen_files = set()
        seen_files.add(f)
        cur = f
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name", "Kernel Name"):
        continue
    if hdr and launch == a.launch and len(r) > 8 and r[2] == "-":
        d = dict(zip(hdr, r))
        try:
            ie = int(d["Instructions Executed"]); te = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        if ie or smp:
            agg.append((cur, int(r[0]), ie, te, smp, r[1].strip()[:100]))
tot = sum(x[2] for x in agg) or 1
tott = sum(x[3] for x in agg)
tots = sum(x[4] for x in agg) or 1
print("launch %d: %d warp instructions, %.2f active lanes on average, %d stall samples" % (a.launch, tot, tott / tot, tots))
if a.ranges:
    for spec in a.ranges.split(","):
        name, f, lh = spec.split(":")
        lo, hi = (int(x) for x in lh.split("-"))
        s = [x for x in agg if x[0] == f and lo <= x[1] <= hi]
        ie = sum(x[2] for x in s); te = sum(x[3] for x in s); sm = sum(x[4] for x in s)
        print("  %-18s inst %5.1f %%   lanes %5.1f   samples %5.1f %%" % (name, 100.0 * ie / tot, te / max(ie, 1), 100.0 * sm / tots))
byfile = {}
for x in agg:
    e = byfile.setdefault(x[0], [0, 0, 0]); e[0] += x[2]; e[1] += x[3]; e[2] += x[4]
for f, e in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  file %-28s inst %5.1f %%   lanes %5.1f   samples %5.1f %%" % (f, 100.0 * e[0] / tot, e[1] / max(e[0], 1), 100.0 * e[2] / tots))
print("top lines by stall samples:")
for x in sorted(agg, key=lambda x: -x[4])[:a.top]:
    print("  %s:%d  inst %.1f %%  lanes %.1f  samples %.1f %%  | %s" % (x[0], x[1], 100.0 * x[2] / tot, x[3] / max(x[2], 1), 100.0 * x[4] / tots, x[5]))

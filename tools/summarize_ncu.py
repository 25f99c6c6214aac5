#!/usr/bin/env python
"""Summarises one `ncu --set full` report into profiles/<round>_<name>.md and records the per-launch DRAM traffic of
each kernel family in profiles/<round>_traffic.json (bench.py's `roofline.traffic` reads it).

  python tools/summarize_ncu.py --round r02 --config c3 --name trace_c3 --report gpurun_out/r02_trace_c3.ncu-rep [--command "..."]

The traffic file is {config: {family: {"dram_bytes_per_launch": mean, "dram_pct_of_peak": mean, "source": path}}} with
family one of trace_kernel_batched, trace_kernel_batched_shadow, shade_kernel, ...
"""
import argparse
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--round", default="r02")
ap.add_argument("--config", required=True)
ap.add_argument("--name", required=True)
ap.add_argument("--report", required=True)
ap.add_argument("--command", default="")
a = ap.parse_args()

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def family(name):
    n = name.split("(")[0].replace("void ", "").replace("gtlc::", "")
    m = re.search(r"(trace_kernel\w*)<(?:\(int\))?(\d)", name)
    if m:
        return m.group(1) + {"0": "", "1": "_any", "2": "_shadow"}.get(m.group(2), "")
    return n.split("<")[0]


if a.report.endswith(".csv"):          # the raw page exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv`) when the report itself was too large to bring back
    raw = open(a.report).read()
else:
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
md = os.path.join(out_dir, "%s_%s.md" % (a.round, a.name))
with open(md, "w") as f:
    f.write("# %s %s: `ncu --set full --clock-control none --import-source on` (%s)\n\n" % (a.round, a.name, os.path.basename(a.report)))
    if a.command:
        f.write("Command: `%s`\n\n" % a.command)
    f.write("One column per captured launch.\n\n| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |\n")
    f.write("|---|---|" + "---:|" * len(data) + "\n")
    f.write("| kernel | | " + " | ".join("`%s`" % family(r[ki]) for r in data) + " |\n")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in data)))
ir, iw, ip = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
fam = {}
for r in data:
    b = float(r[ir]) * TO_BYTES.get(units[ir], 1.0) + float(r[iw]) * TO_BYTES.get(units[iw], 1.0)
    e = fam.setdefault(family(r[ki]), [[], []])
    e[0].append(b); e[1].append(float(r[ip]))
tp = os.path.join(out_dir, "%s_traffic.json" % a.round)
traffic = {}
if os.path.exists(tp):
    with open(tp) as f:
        traffic = json.load(f)
cfg = traffic.setdefault(a.config, {})
for k, (b, p) in fam.items():
    cfg[k] = {"dram_bytes_per_launch": sum(b) / len(b), "dram_pct_of_peak": sum(p) / len(p), "launches_captured": len(b),
              "source": "profiles/%s_%s.md" % (a.round, a.name)}
with open(tp, "w") as f:
    json.dump(traffic, f, indent=1, sort_keys=True)
print("wrote", md, {k: round(v["dram_bytes_per_launch"] / 1e6, 1) for k, v in cfg.items()})

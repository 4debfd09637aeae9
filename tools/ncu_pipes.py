#!/usr/bin/env python
"""Pipe utilisation / issue metrics of every launch in an ncu report (raw page):
    python tools/ncu_pipes.py report.ncu-rep [label ...]"""
import csv
import subprocess
import sys

rep, labels = sys.argv[1], sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
KEYS = ("inst_executed_pipe_", "issue_active", "inst_executed.avg.per_cycle_active", "inst_issued.avg.per_cycle_active",
        "warps_eligible", "gpu__time_duration.sum", "cycles_elapsed.max", "thread_inst_executed_per_inst_executed")
for li, r in enumerate(rows[2:]):
    print("launch", li, labels[li] if li < len(labels) else "", r[h.index("Kernel Name")][-48:])
    for i, k in enumerate(h):
        if any(t in k for t in KEYS) and "pct" in k or k in ("gpu__time_duration.sum",) or "per_cycle_active" in k:
            if i < len(r) and r[i] not in ("", "0"):
                print("   %-95s %s" % (k, r[i]))

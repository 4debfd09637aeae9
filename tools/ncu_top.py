#!/usr/bin/env python
"""Every `pct_of_peak` metric above a threshold for each launch of an ncu report (what is close to saturation?):
    python tools/ncu_top.py report.ncu-rep [threshold=40]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for li, r in enumerate(rows[2:]):
    print("launch", li, r[h.index("Kernel Name")][-48:], "time", r[h.index("gpu__time_duration.sum")])
    hits = []
    for i, k in enumerate(h):
        if "pct" in k and i < len(r):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v >= thr and ".min." not in k and ".max." not in k and ".sum." not in k:
                hits.append((v, k))
    for v, k in sorted(hits, reverse=True)[:40]:
        print("   %6.1f  %s" % (v, k))

#!/usr/bin/env python
"""Summarise the SASS lines with the most warp-stall samples from an ncu report.
usage: tools/ncu_hot.py report.ncu-rep <launch index in report> [top N]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
# the csv holds one block per kernel launch, each starting with a "Kernel Name" line
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
rows = list(csv.reader(io.StringIO("\n".join(blocks[idx]))))
hdr = rows[0]
si = hdr.index("# Samples")
src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [(int(r[si] or 0), n, r) for n, r in enumerate(rows[1:]) if len(r) > si]
total = sum(d[0] for d in data)
print("total samples", total, "instructions", len(data))
for s, n, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  #%-5d %-70s %s" % (s, 100.0 * s / max(total, 1), n, r[src].strip()[:70], st))

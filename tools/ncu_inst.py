#!/usr/bin/env python
"""Executed warp-instructions per SASS basic block (runs of equal execution count) from the CSV of
    ncu -i report.ncu-rep --page source --csv --print-source sass > sass.csv
usage: tools/ncu_inst.py sass.csv [top N] [--dump FIRST LAST]   (--dump prints the SASS lines FIRST..LAST)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ie, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = [(n, r[isrc].strip(), int(r[ie] or 0), int(r[isamp] or 0)) for n, r in enumerate(rows[2:]) if len(r) > ie]
if "--dump" in sys.argv:
    i = sys.argv.index("--dump")
    for d in data[int(sys.argv[i + 1]):int(sys.argv[i + 2]) + 1]:
        print("%5d x %9d  %s" % (d[0], d[2], d[1]))
    sys.exit(0)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
tot = sum(d[2] for d in data)
print(rows[0][1][:120])
print("total warp instructions", tot)
blocks, cur = [], [data[0]]
for d in data[1:]:
    if d[2] == cur[-1][2]:
        cur.append(d)
    else:
        blocks.append(cur)
        cur = [d]
blocks.append(cur)
for b in sorted(blocks, key=lambda b: -b[0][2] * len(b))[:top]:
    ops = {}
    for d in b:
        o = d[1].split()[1] if d[1].startswith("@") else d[1].split()[0]
        ops[o] = ops.get(o, 0) + 1
    print("sass #%4d-%4d  len %3d  x %9d  %4.1f%% of all instr  stall samples %5d | %s" % (
        b[0][0], b[-1][0], len(b), b[0][2], 100.0 * b[0][2] * len(b) / tot, sum(d[3] for d in b),
        " ".join("%s=%d" % x for x in sorted(ops.items(), key=lambda x: -x[1])[:8])))

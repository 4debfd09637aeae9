#!/usr/bin/env python
"""Aggregate warp-stall samples by reason (and by SASS opcode) for one launch of an ncu report.
usage: tools/ncu_stalls.py report.ncu-rep <launch index>"""
import collections, csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
rows = list(csv.reader(io.StringIO("\n".join(blocks[idx]))))
hdr = rows[0]
si, src = hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_reason, by_op = collections.Counter(), collections.Counter()
ex = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
by_op_exec = collections.Counter()
for r in rows[1:]:
    if len(r) <= si: continue
    for i in stall_cols:
        by_reason[hdr[i]] += int(r[i] or 0)
    toks = r[src].split()
    op = next((t for t in toks if not t.startswith('@')), '?').split('.')[0]
    by_op[op] += int(r[si] or 0)
    if ex is not None: by_op_exec[op] += int(r[ex] or 0)
tot = sum(by_reason.values())
print("stall reasons:", [(k, v, "%.1f%%" % (100.0 * v / tot)) for k, v in by_reason.most_common(10)])
print("samples by opcode:", by_op.most_common(25))
if ex is not None:
    print("warp-instructions executed by opcode:", by_op_exec.most_common(30))

#!/usr/bin/env python
"""Condense an ncu report (--set full) into the per-launch metrics quoted in DESIGN.md / profiles/.
usage: tools/ncu_summary.py report.ncu-rep [layer names...] > profiles/rNN/xxx_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__cluster_dim_x", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]

rep = sys.argv[1]
names = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for n, r in enumerate(rows[2:]):
    print("launch %d %s" % (n, names[n] if n < len(names) else ""))
    for k in ["Kernel Name", "Grid Size", "Block Size"]:
        print("   %-70s %s" % (k, r[col[k]]))
    for k in KEYS:
        if k in col:
            print("   %-70s %s %s" % (k, r[col[k]], units[col[k]]))

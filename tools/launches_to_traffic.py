#!/usr/bin/env python
"""Condenses an ncu launch list (gpu__time_duration + dram bytes per launch, CSV) into the DRAM-traffic record
bench.py quotes as roofline.traffic:  python tools/launches_to_traffic.py launches_bf16.csv bf16 "<source text>" > out.json"""
import csv
import json
import sys

path, dtype, source = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]
iid, im, iv = h.index("ID"), h.index("Metric Name"), h.index("Metric Value")
d = {}
for r in rows[1:]:
    d.setdefault(int(r[iid]), {})[r[im]] = float(r[iv].replace(",", ""))
rd = sum(v.get("dram__bytes_read.sum", 0) for v in d.values())
wr = sum(v.get("dram__bytes_write.sum", 0) for v in d.values())
ns = sum(v.get("gpu__time_duration.sum", 0) for v in d.values())
json.dump({"dtype": dtype, "launches": len(d), "dram_bytes_read": rd, "dram_bytes_write": wr, "bytes_per_step": rd + wr,
           "sum_kernel_ms_under_ncu": ns / 1e6, "source": source}, sys.stdout, indent=1)

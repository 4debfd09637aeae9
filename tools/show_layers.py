#!/usr/bin/env python
"""print per-layer reports (bench.py --layer-report) side by side: tools/show_layers.py new.json [old.json]"""
import json, sys
new = json.load(open(sys.argv[1]))
old = {l['name']: l for l in json.load(open(sys.argv[2]))['layers']} if len(sys.argv) > 2 else {}
tot = told = 0
for l in new['layers']:
    o = old.get(l['name'])
    tot += l['ms'] * l['count']
    told += o['ms'] * o['count'] if o else 0
    print("%-10s %-8s x%d %.4f ms (was %s) tot %.3f %7.1f TF %7.1f GB/s %-6s floor %.4f ff %.3f" % (
        l['name'], l['algo'].split('_')[-1], l['count'], l['ms'], "%.4f" % o['ms'] if o else "-", l['ms'] * l['count'],
        l['tflops'], l['gbs'], l['bound'], l['floor_ms'], l['frac_of_floor']))
print("total %.3f ms (was %.3f)" % (tot, told))

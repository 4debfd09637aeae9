#!/usr/bin/env python
"""Pipeline-depth sensitivity: a ResNet-50 layer at forced stage counts (tools build, B200CB_STAGES), single CTA
and CTA pair.  If time falls with every added stage the main loop is bound by TMA round-trip latency.
    B200CB_LIB=megengine_b200/lib/libb200cb_tools.so python tools/exp_stages.py l3.bN.c2 l4.bN.c1"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megengine_b200 import _capi as C  # noqa: E402
from tests import _cases as K  # noqa: E402
from tools.sweep_resnet import time_plan  # noqa: E402

lib = C.load()
names = sys.argv[1:] or ["l3.bN.c2"]
for (name, ic, h, w, oc, f, s, pd, has_z, cnt) in K.RESNET50_LAYERS:
    if name not in names:
        continue
    d = K.make_desc("bf16", 256, ic, h, w, oc, f, stride=s, pad=pd, nonline=C.NL_RELU, has_z=bool(has_z))
    for pair in (0, 1):
        out = []
        for st in (2, 3, 4, 5, 6, 7, 8):
            os.environ["B200CB_PAIR"] = str(pair)
            os.environ["B200CB_STAGES"] = str(st)
            plan = C.Plan()
            if lib.b200cb_query(ctypes.byref(d), ctypes.byref(plan)) != 0:
                continue
            if plan.stages != st:
                continue
            ms = time_plan(lib, plan, d, 20)
            out.append("st%d %.4f (%.0f TF)" % (st, ms, plan.flops / ms / 1e9))
        print("%-9s bn %d pair %d | %s" % (name, plan.block_n, pair, " | ".join(out)), flush=True)

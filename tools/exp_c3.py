#!/usr/bin/env python
"""Why are the residual 1x1 layers (l3.c3, l4.c3) short of their HBM floor?  Same shape with / without z,
z-ring depth 4 / 8 (more pipeline stages), TMA / direct stores, per tile width (tools build knobs).
    B200CB_LIB=megengine_b200/lib/libb200cb_tools.so python tools/exp_c3.py l3.c3 l4.c3"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megengine_b200 import _capi as C  # noqa: E402
from tests import _cases as K  # noqa: E402
from tools.sweep_resnet import time_plan  # noqa: E402

lib = C.load()
names = sys.argv[1:] or ["l3.c3"]
KNOBS = ("B200CB_PAIR", "B200CB_CLUSTER", "B200CB_NZ", "B200CB_STAGES", "B200CB_DIRECT")
VARIANTS = [("base", True, {}), ("no_z", False, {}), ("nz4", True, {"B200CB_NZ": "4"}), ("nz8", True, {"B200CB_NZ": "8"}),
            ("nz4_direct", True, {"B200CB_NZ": "4", "B200CB_DIRECT": "1"}), ("direct", True, {"B200CB_DIRECT": "1"}),
            ("no_z_direct", False, {"B200CB_DIRECT": "1"})]
for (name, ic, h, w, oc, f, s, pd, has_z, cnt) in K.RESNET50_LAYERS:
    if name not in names:
        continue
    for vn, z, env in VARIANTS:
        d = K.make_desc("bf16", 256, ic, h, w, oc, f, stride=s, pad=pd, nonline=C.NL_RELU, has_z=z)
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        os.environ["B200CB_PAIR"] = "0"
        out = []
        for algo in range(lib.b200cb_algo_count()):
            nm = lib.b200cb_algo_name(algo).decode()
            if "IGEMM" not in nm:
                continue
            plan = C.Plan()
            if lib.b200cb_query_algo(ctypes.byref(d), algo, ctypes.byref(plan)) != 0:
                continue
            ms = time_plan(lib, plan, d, 20)
            out.append("%s st%d nz%d %.4f" % (nm.split("_")[-1], plan.stages, plan.reserved[5] >> 16, ms))
        print("%-8s %-12s %s" % (name, vn, " | ".join(out)), flush=True)

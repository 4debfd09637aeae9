#!/usr/bin/env python
"""L2 -> SM traffic experiment: one ResNet-50 layer under {single CTA, multicast cluster of 2 / 4, CTA pair}
(tools build knobs), 2 launches each -- run under `ncu --metrics ...` to read the crossbar / LTS bytes per
variant, or plain for event timings.
    B200CB_LIB=megengine_b200/lib/libb200cb_tools.so python tools/exp_cluster.py l3.bN.c2 l3.bN.c1"""
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
names = sys.argv[1:] or ["l3.bN.c2"]
reps = int(os.environ.get("EXP_REPS", "20"))
VARIANTS = [("single", {"B200CB_PAIR": "0", "B200CB_CLUSTER": "1"}), ("mc2", {"B200CB_PAIR": "0", "B200CB_CLUSTER": "2"}),
            ("mc4", {"B200CB_PAIR": "0", "B200CB_CLUSTER": "4"}), ("pair", {"B200CB_PAIR": "1"})]
for (name, ic, h, w, oc, f, s, pd, has_z, cnt) in K.RESNET50_LAYERS:
    if name not in names:
        continue
    d = K.make_desc("bf16", 256, ic, h, w, oc, f, stride=s, pad=pd, nonline=C.NL_RELU, has_z=bool(has_z))
    for vn, env in VARIANTS:
        for k in ("B200CB_PAIR", "B200CB_CLUSTER"):
            os.environ.pop(k, None)
        os.environ.update(env)
        plan = C.Plan()
        rc = lib.b200cb_query(ctypes.byref(d), ctypes.byref(plan))
        assert rc == 0
        ms = time_plan(lib, plan, d, reps)
        print("%-10s %-7s bn %3d cluster %d pair %d stages %d  %.4f ms  %.0f TF" % (
            name, vn, plan.block_n, plan.reserved[6], plan.reserved[7], plan.stages, ms, plan.flops / ms / 1e9), flush=True)

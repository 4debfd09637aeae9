#!/usr/bin/env python
"""time one dense 1x1 layer: tools/pw_one.py IC OC H W [dtype] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megengine_b200 import _capi as C
from tests import _cases as K
from tools.bench_layers import time_desc
ic, oc, h, w = map(int, sys.argv[1:5])
dt = sys.argv[5] if len(sys.argv) > 5 else "bf16"
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20
d = K.make_desc(dt, 256, ic, h, w, oc, 1, nonline=C.NL_RELU)
print(time_desc(C.load(), d, reps, {"hbm_gbs": 6544.3}))

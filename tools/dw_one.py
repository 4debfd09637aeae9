#!/usr/bin/env python
"""time one depthwise layer: tools/dw_one.py C H W stride [dtype] [reps]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megengine_b200 import _capi as C
from tests import _cases as K
from tools.bench_layers import time_desc
c, h, w, s = map(int, sys.argv[1:5])
dt = sys.argv[5] if len(sys.argv) > 5 else "bf16"
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20
d = K.dw_desc(dt, 256, c, h, w, s)
print(time_desc(C.load(), d, reps, {"hbm_gbs": 6544.3}))

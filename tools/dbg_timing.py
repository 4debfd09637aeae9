#!/usr/bin/env python
"""Where do the warp roles of igemm_kernel wait?  Needs the instrumented build
(make -C megengine_b200/csrc dbg) and B200CB_LIB pointing at it:

    B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so python tools/dbg_timing.py [bf16|qs8] [layer ...]
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
NC = 17  # 16 counters + the role's total
dbg = torch.zeros(148 * 3 * NC, dtype=torch.int64, device="cuda")
os.environ["B200CB_DBG_PTR"] = hex(dbg.data_ptr())
from megengine_b200 import _capi as C
from tests import _cases as K
lib = C.load()
kind = sys.argv[1] if len(sys.argv) > 1 else "bf16"
only = set(sys.argv[2:])
names = ["empty", "tempty", "full", "tfull", "zfull", "zempty", "bres", "stwait", "tmem_ld", "body", "store", "zstore_rd",
         "tma_issue", "x13", "x14", "x15"]
# "pw" as the only layer name: the MobileNet pointwise table (tools/bench_layers.py) instead of ResNet-50
if only == {"pw"}:
    from tools.bench_layers import MOBILENET_PW
    work = [(("pw_%dto%d_%d" % (ic, oc, h),), K.make_desc(kind, 256, ic, h, w, oc, 1, nonline=C.NL_RELU))
            for oc, ic, h, w in MOBILENET_PW]
else:
    work = [(layer, K.resnet50_desc(layer, kind, 256, ic_pad=8 if kind == "qs8" else 4))
            for layer in K.RESNET50_LAYERS if not only or layer[0] in only]
for layer, d in work:
    plan = C.Plan()
    assert lib.b200cb_query(ctypes.byref(d), ctypes.byref(plan)) == 0
    tdt = torch.int8 if kind == "qs8" else torch.bfloat16
    mk = lambda n: (torch.randint(-8, 9, (n,), dtype=torch.int8, device="cuda") if kind == "qs8"
                    else torch.randn(n, device="cuda").to(tdt))
    src, flt = mk(d.n * d.ih * d.iw * d.ic), mk(d.oc * d.ic * d.fh * d.fw)
    bias = torch.randint(-50, 50, (d.oc,), dtype=torch.int32, device="cuda") if kind == "qs8" else mk(d.oc)
    z = mk(d.n * d.oh * d.ow * d.oc) if d.has_z else None
    dst = torch.empty(d.n * d.oh * d.ow * d.oc, dtype=tdt, device="cuda")
    ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
    p = lambda t: t.data_ptr() if t is not None else None
    for _ in range(3):
        dbg.zero_()
        rc = lib.b200cb_launch(ctypes.byref(plan), p(src), p(flt), p(bias), p(z), p(dst), p(ws), ws.numel(),
                               torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
    t = dbg.view(148, 3, NC).double()
    used = t[:, 1, NC - 1] > 0.5 * t[:, 1, NC - 1].max()  # CTAs whose MMA warp ran (pair: the leaders)
    t = t[used].mean(0)
    tot = t[:, NC - 1]
    fmt = lambda role: " ".join("%s %4.1f%%" % (names[i], 100 * t[role, i] / tot[role]) for i in range(16) if t[role, i] > 0.005 * tot[role])
    print("%-10s mode %d bn %3d stages %2d  cycles %8.0f | producer: %s | mma: %s | epi(w4): %s" % (
        layer[0], plan.a_mode, plan.block_n, plan.stages, tot[1], fmt(0), fmt(1), fmt(2)))

#!/usr/bin/env python
"""Resident filter tile (IgemmParams::b_res) on the 1x1 stride-1 layers: per layer and tile width, the streamed
filter (product default, pairs as the heuristic picks them) against the resident one (single CTAs), with the z
ring at depth 4 / 8; every variant's dst is compared bit for bit with the streamed kernel's.
    B200CB_LIB=megengine_b200/lib/libb200cb_tools.so python tools/exp_bres.py [bf16 qs8] [--only l3.c3]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megengine_b200 import _capi as C  # noqa: E402
from tests import _cases as K  # noqa: E402
from tools.bench_layers import TORCH_DT, rand  # noqa: E402

lib = C.load()
args = sys.argv[1:]
only = ""
if "--only" in args:
    only = args[args.index("--only") + 1]
    args = [a for a in args if a not in ("--only", only)]
kinds = args or ["bf16", "qs8"]
KNOBS = ("B200CB_PAIR", "B200CB_BRES", "B200CB_NZ", "B200CB_STAGES")
REPS = 30


def run(plan, d, t):
    s = torch.cuda.current_stream().cuda_stream
    p = lambda x: x.data_ptr() if x is not None else None
    copies = len(t["src"])

    def launch(i):
        rc = lib.b200cb_launch(ctypes.byref(plan), p(t["src"][i]), p(t["flt"]), p(t["bias"]), p(t["z"][i]),
                               p(t["dst"][i]), None, 0, s)
        assert rc == 0, lib.b200cb_strerror(rc)

    for i in range(copies):
        launch(i)
    torch.cuda.synchronize()
    out = t["dst"][0].clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(REPS):
        launch(r % copies)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS, out


for kind in kinds:
    for (name, ic, h, w, oc, f, s, pd, has_z, cnt) in K.RESNET50_LAYERS:
        if f != 1 or s != 1 or (only and only not in name):
            continue
        d = K.make_desc(kind, 256, ic, h, w, oc, f, stride=s, pad=pd,
                        nonline=C.NL_IDENTITY if name.endswith(".ds") else C.NL_RELU, has_z=bool(has_z))
        sdt, ddt = TORCH_DT[d.src_dtype], TORCH_DT[d.dst_dtype]
        n_src, n_dst = d.n * d.ih * d.iw * d.ic, d.n * d.oh * d.ow * d.oc
        per_copy = (n_src + n_dst * (2 if has_z else 1)) * sdt.itemsize
        copies = max(2, min(16, int(300e6 // per_copy) + 1))
        torch.manual_seed(1)
        t = {"src": [rand((n_src,), sdt) for _ in range(copies)],
             "dst": [torch.empty((n_dst,), dtype=ddt, device="cuda") for _ in range(copies)],
             "z": [rand((n_dst,), ddt, -127, 127) for _ in range(copies)] if has_z else [None] * copies,
             "flt": rand((d.oc * d.ic,), sdt), "bias": rand((d.oc,), TORCH_DT[d.bias_dtype], -50, 50)}
        for algo in range(lib.b200cb_algo_count()):
            nm = lib.b200cb_algo_name(algo).decode()
            if "IGEMM" not in nm:
                continue
            cells, ref = [], None
            variants = [("stream", {})] + [("res", {"B200CB_BRES": "1", "B200CB_PAIR": "0"})] + \
                ([("res_nz4", {"B200CB_BRES": "1", "B200CB_PAIR": "0", "B200CB_NZ": "4"}),
                  ("res_nz8", {"B200CB_BRES": "1", "B200CB_PAIR": "0", "B200CB_NZ": "8"})] if has_z else [])
            for vn, env in variants:
                for k in KNOBS:
                    os.environ.pop(k, None)
                os.environ.update(env)
                plan = C.Plan()
                if lib.b200cb_query_algo(ctypes.byref(d), algo, ctypes.byref(plan)) != 0:
                    continue
                if vn != "stream" and not (plan.reserved[2] & 8):
                    continue
                ms, out = run(plan, d, t)
                if ref is None:
                    ref = out
                same = bool(torch.equal(out.view(torch.uint8), ref.view(torch.uint8)))
                cells.append("%s st%d nz%d p%d %.4f%s" % (vn, plan.stages, plan.reserved[5] >> 16, plan.reserved[7],
                                                          ms, "" if same else " MISMATCH"))
            print("%-5s %-9s %-8s %s" % (kind, name, nm.split("_")[-1], " | ".join(cells)), flush=True)

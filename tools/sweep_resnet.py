#!/usr/bin/env python
"""Tile / CTA-pair sweep over the 24 unique ResNet-50 ConvBias shapes (N = 256): every registered
igemm algorithm (tile width) x {single CTA, cta_group::2 pair}, timed L2-cold (rotating replicas,
CUDA events).  Feeds the heuristic in csrc/capi.cpp and profiles/.

    python tools/sweep_resnet.py --dtype bf16 --out gpurun_out/sweep_bf16.json
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megengine_b200 import _capi as C  # noqa: E402
from tests import _cases as K  # noqa: E402
from tools.bench_layers import TORCH_DT, rand  # noqa: E402


def time_plan(lib, plan, d, reps):
    sdt, ddt = TORCH_DT[d.src_dtype], TORCH_DT[d.dst_dtype]
    n_src, n_dst = d.n * d.ih * d.iw * d.ic, d.n * d.oh * d.ow * d.oc
    n_flt = d.oc * d.ic * d.fh * d.fw
    per_copy = (n_src + n_dst * (2 if d.has_z else 1)) * sdt.itemsize
    copies = max(2, min(16, int(300e6 // per_copy) + 1))
    srcs = [rand((n_src,), sdt) for _ in range(copies)]
    dsts = [torch.empty((n_dst,), dtype=ddt, device="cuda") for _ in range(copies)]
    zs = [rand((n_dst,), ddt, -127, 127) for _ in range(copies)] if d.has_z else [None] * copies
    flt = rand((n_flt,), sdt)
    bias = rand((d.oc,), TORCH_DT[d.bias_dtype], -50, 50)
    ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device="cuda") if plan.workspace_bytes else None
    s = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr() if t is not None else None

    def launch(i):
        rc = lib.b200cb_launch(ctypes.byref(plan), p(srcs[i]), p(flt), p(bias), p(zs[i]), p(dsts[i]), p(ws),
                               int(plan.workspace_bytes), s)
        assert rc == 0, lib.b200cb_strerror(rc)

    for i in range(copies):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        launch(r % copies)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    lib = C.load()
    rows = []
    for (name, ic, h, w, oc, f, s, pd, has_z, cnt) in K.RESNET50_LAYERS:
        if a.only and a.only not in name:
            continue
        if ic < 8:
            ic = 16 if a.dtype == "qs8" else 8
        d = K.make_desc(a.dtype, a.batch, ic, h, w, oc, f, stride=s, pad=pd,
                        nonline=C.NL_IDENTITY if name.endswith(".ds") else C.NL_RELU, has_z=bool(has_z))
        res = {}
        for algo in range(lib.b200cb_algo_count()):
            nm = lib.b200cb_algo_name(algo).decode()
            if "IGEMM" not in nm:
                continue
            for pair in (0, 1):
                os.environ["B200CB_PAIR"] = str(pair)
                plan = C.Plan()
                if lib.b200cb_query_algo(ctypes.byref(d), algo, ctypes.byref(plan)) != 0:
                    continue
                if pair and not plan.reserved[7]:
                    continue
                if a.verbose:
                    print("  %s %s pair %d stages %d smem %d nz %d ..." % (name, nm, pair, plan.stages, plan.smem_bytes,
                                                                       plan.reserved[5] >> 16), flush=True)
                ms = time_plan(lib, plan, d, a.reps)
                res["%s%s" % (nm.split("_")[-1], "p" if pair else "")] = round(ms, 4)
        os.environ.pop("B200CB_PAIR", None)
        plan = C.Plan()
        lib.b200cb_query(ctypes.byref(d), ctypes.byref(plan))
        pick = "%s%s" % (lib.b200cb_algo_name(plan.algo).decode().split("_")[-1], "p" if plan.reserved[7] else "")
        best = min(res, key=res.get)
        tf = plan.flops / res[best] / 1e9
        print("%-10s x%d mode%d pick %-9s %.4f | best %-9s %.4f (%.0f TF) | %s" % (
            name, cnt, plan.a_mode, pick, res.get(pick, -1), best, res[best], tf,
            " ".join("%s=%.4f" % kv for kv in sorted(res.items()))), flush=True)
        rows.append({"name": name, "count": cnt, "pick": pick, "best": best, "ms": res, "flops": int(plan.flops),
                     "bytes": int(plan.bytes)})
    tot_pick = sum(r["ms"].get(r["pick"], 0) * r["count"] for r in rows)
    tot_best = sum(r["ms"][r["best"]] * r["count"] for r in rows)
    print("sum over 53 layers: heuristic pick %.3f ms, best-of-sweep %.3f ms" % (tot_pick, tot_best))
    if a.out:
        json.dump({"dtype": a.dtype, "batch": a.batch, "layers": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

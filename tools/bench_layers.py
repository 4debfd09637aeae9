#!/usr/bin/env python
"""Per-layer ConvBias timing on one B200 for the shape tables of SURVEY.md 8(d): MobileNetV2 /
ShuffleNetV2 depthwise (config 4), MobileNet pointwise, and (optionally) the ResNet-50 table.

Every layer is timed L2-cold: its tensors are replicated until one rotation touches > 256 MB, and
consecutive launches rotate through the replicas.  CUDA events around REPS back-to-back launches.

    python tools/bench_layers.py --set dw --dtype bf16 --out gpurun_out/dw_bf16.json
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

SHUFFLENET_DW = [(24, 56, 56, 2), (58, 56, 56, 2), (58, 28, 28, 1), (116, 28, 28, 2), (116, 14, 14, 1),
                 (232, 14, 14, 2), (232, 7, 7, 1)]
# MobileNet pointwise 1x1 (OC, IC, H, W) -- dnn/test/arm_common/conv_bias.cpp:2210-2219
MOBILENET_PW = [(16, 32, 112, 112), (96, 16, 112, 112), (144, 24, 56, 56), (192, 32, 28, 28), (384, 64, 28, 28),
                (576, 96, 14, 14), (960, 160, 7, 7), (320, 960, 7, 7), (1280, 320, 7, 7)]

TORCH_DT = {C.DT_BF16: torch.bfloat16, C.DT_F16: torch.float16, C.DT_F32: torch.float32, C.DT_QS8: torch.int8,
            C.DT_QS32: torch.int32}


def rand(shape, dt, lo=-8, hi=8):
    if dt in (torch.int8, torch.int32):
        return torch.randint(lo, hi + 1, shape, dtype=dt, device="cuda")
    return torch.randn(shape, dtype=torch.float32, device="cuda").to(dt)


def time_desc(lib, d, reps, peaks, algo=-1):
    plan = C.Plan()
    rc = lib.b200cb_query(ctypes.byref(d), ctypes.byref(plan)) if algo < 0 else \
        lib.b200cb_query_algo(ctypes.byref(d), algo, ctypes.byref(plan))
    if rc != 0:
        return {"desc": K.desc_str(d), "error": lib.b200cb_strerror(rc).decode()}
    sdt, ddt = TORCH_DT[d.src_dtype], TORCH_DT[d.dst_dtype]
    n_src = d.n * d.ih * d.iw * d.ic
    n_dst = d.n * d.oh * d.ow * d.oc
    n_flt = d.oc * (d.ic // d.group) * d.fh * d.fw
    per_copy = (n_src + n_dst * (2 if d.has_z else 1)) * sdt.itemsize
    copies = max(2, min(64, int(300e6 // per_copy) + 1))
    srcs = [rand((n_src,), sdt) for _ in range(copies)]
    dsts = [torch.empty((n_dst,), dtype=ddt, device="cuda") for _ in range(copies)]
    zs = [rand((n_dst,), ddt, -127, 127) for _ in range(copies)] if d.has_z else [None] * copies
    flt = rand((n_flt,), sdt)
    bias = rand((d.oc,), TORCH_DT[d.bias_dtype], -50, 50) if d.bias_mode == C.BIAS_CHANNEL else None
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
    ms = e0.elapsed_time(e1) / reps
    gbs = plan.bytes / ms / 1e6
    tf = plan.flops / ms / 1e9
    return {"desc": K.desc_str(d), "algo": lib.b200cb_algo_name(plan.algo).decode(), "ms": round(ms, 5),
            "gbs": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 3), "tflops": round(tf, 2),
            "bytes": int(plan.bytes), "flops": int(plan.flops), "copies": copies}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="dw", choices=["dw", "pw", "resnet"])
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--out", default="")
    ap.add_argument("--algo", type=int, default=-1, help="force an algorithm index (default: heuristic)")
    args = ap.parse_args()
    lib = C.load()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    peaks = {"hbm_gbs": pk["hbm_gbs"], "tf": pk["bf16_tflops"]}
    descs = []
    if args.set == "dw":
        for c, h, w, s in K.MOBILENETV2_DW + SHUFFLENET_DW:
            if args.dtype == "qs8" and c % 16:
                continue
            if args.dtype == "f32" and c % 4:
                continue
            descs.append(K.dw_desc(args.dtype, args.batch, c, h, w, s))
    elif args.set == "pw":
        for oc, ic, h, w in MOBILENET_PW:
            if args.dtype == "qs8" and (ic % 16 or oc % 16):
                continue
            descs.append(K.make_desc(args.dtype, args.batch, ic, h, w, oc, 1, nonline=C.NL_RELU))
    else:
        for layer in K.RESNET50_LAYERS:
            descs.append(K.resnet50_desc(layer, args.dtype, args.batch, ic_pad=16 if args.dtype == "qs8" else 8))
    rows = []
    for d in descs:
        r = time_desc(lib, d, args.reps, peaks, args.algo)
        rows.append(r)
        print(json.dumps(r), flush=True)
    ok = [r for r in rows if "ms" in r]
    tot_b, tot_ms = sum(r["bytes"] for r in ok), sum(r["ms"] for r in ok)
    summary = {"set": args.set, "dtype": args.dtype, "batch": args.batch, "layers": len(ok), "total_ms": round(tot_ms, 4),
               "total_gbs": round(tot_b / tot_ms / 1e6, 1), "frac_hbm": round(tot_b / tot_ms / 1e6 / peaks["hbm_gbs"], 3),
               "peak_hbm_gbs": peaks["hbm_gbs"]}
    if args.set == "dw":
        # per-network aggregates: MobileNetV2 (channel counts in 16-byte pixels, TMA-staged kernel) and
        # ShuffleNetV2-1x (58 / 116 channels are not 16-byte rows: register-path kernel)
        nm = len([1 for c, _h, _w, _s in K.MOBILENETV2_DW if not (args.dtype == "qs8" and c % 16)])
        for name, part in (("mobilenetv2", rows[:nm]), ("shufflenetv2", rows[nm:])):
            part = [r for r in part if "ms" in r]
            if part:
                b, ms = sum(r["bytes"] for r in part), sum(r["ms"] for r in part)
                summary[name] = {"layers": len(part), "total_ms": round(ms, 4), "total_gbs": round(b / ms / 1e6, 1),
                                 "frac_hbm": round(b / ms / 1e6 / peaks["hbm_gbs"], 3)}
    print(json.dumps(summary), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"summary": summary, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

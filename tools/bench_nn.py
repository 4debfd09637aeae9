#!/usr/bin/env python
"""Timing of the operators either side of ConvBias (include/b200cb_nn.h) at ResNet-50 / MobileNet sizes, N = 256,
L2-cold (rotating replicas), CUDA events.  GB/s = algorithmic bytes (every tensor once) / time, against the measured
HBM copy bandwidth; deconvolutions in TFLOP/s of the gradient's own flops.
    python tools/bench_nn.py [--out gpurun_out/nn.json]"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megengine_b200 import _capi as C  # noqa: E402
from tests import _nn as N  # noqa: E402

TDT = {C.DT_F32: torch.float32, C.DT_F16: torch.float16, C.DT_BF16: torch.bfloat16, C.DT_QS8: torch.int8,
       C.DT_QS32: torch.int32}


def _rand(n, dt):
    if dt in (C.DT_QS8, C.DT_QS32):
        return torch.randint(-100, 100, (n,), dtype=TDT[dt], device="cuda")
    return (torch.rand(n, device="cuda") * 2 - 1).to(TDT[dt])


def _time(launch, copies, reps=20):
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


def _copies(nbytes):
    return max(2, min(16, int(400e6 // max(nbytes, 1)) + 1))


def run(hbm_gbs):
    lib = C.load()
    s = lambda: torch.cuda.current_stream().cuda_stream
    rows = []

    def add(name, ms, nbytes, flops=None):
        r = {"op": name, "ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / hbm_gbs, 3)}
        if flops:
            r["tflops"] = round(flops / ms / 1e9, 1)
        rows.append(r)

    # pooling
    for name, fmt, mode, dt, c, ih, iw, win, st, pad in (
            ("maxpool 3x3/2 112x112x64 bf16 (ResNet stem)", C.FORMAT_NHWC, C.POOL_MAX, C.DT_BF16, 64, 112, 112, 3, 2, 1),
            ("maxpool 3x3/2 112x112x64 qs8", C.FORMAT_NHWC, C.POOL_MAX, C.DT_QS8, 64, 112, 112, 3, 2, 1),
            ("global avgpool 7x7x2048 bf16", C.FORMAT_NHWC, C.POOL_AVERAGE, C.DT_BF16, 2048, 7, 7, 7, 7, 0),
            ("global avgpool 7x7x2048 qs8", C.FORMAT_NHWC, C.POOL_AVERAGE, C.DT_QS8, 2048, 7, 7, 7, 7, 0),
            ("avgpool 2x2/2 56x56x256 f16", C.FORMAT_NHWC, C.POOL_AVERAGE, C.DT_F16, 256, 56, 56, 2, 2, 0),
            ("maxpool 3x3/2 112x112x64 qs8 NCHW4", C.FORMAT_NCHW4, C.POOL_MAX, C.DT_QS8, 64, 112, 112, 3, 2, 1)):
        d = N.pool_desc(fmt, mode, dt, 256, c, ih, iw, win, st, pad)
        es = C.DT_SIZE[dt]
        n_in, n_out = 256 * c * ih * iw, 256 * c * d.oh * d.ow
        k = _copies((n_in + n_out) * es)
        srcs, dsts = [_rand(n_in, dt) for _ in range(k)], [torch.empty(n_out, dtype=TDT[dt], device="cuda") for _ in range(k)]

        def launch(i, d=d, srcs=srcs, dsts=dsts):
            assert lib.b200cb_pooling(ctypes.byref(d), srcs[i].data_ptr(), dsts[i].data_ptr(), s()) == 0
        add(name, _time(launch, k), (n_in + n_out) * es)
        del srcs, dsts

    # elementwise (the residual add of a bottleneck output, 256 x 56 x 56 x 256) and TypeCvt
    n = 256 * 56 * 56 * 256
    for name, dt, mode in (("add+relu 56x56x256 bf16", C.DT_BF16, C.ELEM_FUSE_ADD_RELU), ("add+relu 56x56x256 f16", C.DT_F16, C.ELEM_FUSE_ADD_RELU),
                           ("add 56x56x256 f32", C.DT_F32, C.ELEM_ADD), ("relu 56x56x256 bf16", C.DT_BF16, C.ELEM_RELU)):
        es = C.DT_SIZE[dt]
        binary = mode != C.ELEM_RELU
        k = 2
        a, b, o = [_rand(n, dt) for _ in range(k)], [_rand(n, dt) for _ in range(k)], [torch.empty(n, dtype=TDT[dt], device="cuda") for _ in range(k)]

        def launch(i, mode=mode, dt=dt, a=a, b=b, o=o, binary=binary):
            assert lib.b200cb_elemwise(mode, dt, n, a[i].data_ptr(), b[i].data_ptr() if binary else None, o[i].data_ptr(), s()) == 0
        add(name, _time(launch, k), n * es * (3 if binary else 2))
        del a, b, o
    a, b, o = [_rand(n, C.DT_QS8) for _ in range(2)], [_rand(n, C.DT_QS8) for _ in range(2)], [torch.empty(n, dtype=torch.int8, device="cuda") for _ in range(2)]

    def launch_q(i):
        assert lib.b200cb_elemwise_q(C.ELEM_FUSE_ADD_RELU, n, a[i].data_ptr(), 0.37, b[i].data_ptr(), 0.71, o[i].data_ptr(), 0.9, s()) == 0
    add("QFUSE_ADD_RELU 56x56x256 qs8", _time(launch_q, 2), n * 3)
    del a, b, o
    for name, sdt, ddt in (("TypeCvt QuantizedS32 -> QuantizedS8", C.DT_QS32, C.DT_QS8), ("TypeCvt Float32 -> BFloat16", C.DT_F32, C.DT_BF16),
                           ("TypeCvt QuantizedS8 -> Float32", C.DT_QS8, C.DT_F32)):
        src, dst = [_rand(n, sdt) for _ in range(2)], [torch.empty(n, dtype=TDT[ddt], device="cuda") for _ in range(2)]

        def launch_c(i, sdt=sdt, ddt=ddt, src=src, dst=dst):
            assert lib.b200cb_typecvt(sdt, 0.0013, ddt, 0.9, n, src[i].data_ptr(), dst[i].data_ptr(), s()) == 0
        add(name, _time(launch_c, 2), n * (C.DT_SIZE[sdt] + C.DT_SIZE[ddt]))
        del src, dst

    # ConvolutionBackwardData
    for name, kind, ic, ih, iw, oc, f, st, pad in (("deconv 3x3 s1 256->256 14x14 bf16", "bf16", 256, 14, 14, 256, 3, 1, 1),
                                                    ("deconv 4x4 s2 128->64 56x56 (28x28 diff) bf16", "bf16", 64, 56, 56, 128, 4, 2, 1),
                                                    ("deconv 3x3 s1 256->256 14x14 qs8", "qs8", 256, 14, 14, 256, 3, 1, 1)):
        d = N.deconv_desc(kind, 256, ic, ih, iw, oc, f, stride=st, pad=pad)
        plan = C.Plan()
        rc = lib.b200cb_deconv_query(ctypes.byref(d), ctypes.byref(plan))
        if rc != 0:
            rows.append({"op": name, "unsupported": rc})
            continue
        dt = d.dst_dtype
        diff, grad = _rand(256 * d.oh * d.ow * oc, dt), torch.empty(256 * ih * iw * ic, dtype=TDT[d.src_dtype], device="cuda")
        flt = _rand(oc * f * f * ic, d.filter_dtype)
        ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device="cuda")

        def launch_d(i, plan=plan, flt=flt, diff=diff, grad=grad, ws=ws):
            assert lib.b200cb_deconv_launch(ctypes.byref(plan), flt.data_ptr(), diff.data_ptr(), grad.data_ptr(), ws.data_ptr(),
                                            int(plan.workspace_bytes), s()) == 0
        add(name, _time(launch_d, 1), int(plan.bytes), int(plan.flops))
        del diff, grad, ws
    torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    rows = run(peaks["hbm_gbs"])
    for r in rows:
        print(r)
    if a.out:
        json.dump({"hbm_gbs": peaks["hbm_gbs"], "rows": rows}, open(a.out, "w"), indent=1)

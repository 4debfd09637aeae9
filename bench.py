#!/usr/bin/env python
"""bench.py -- ResNet-50 ConvBias images/s on B200 (BASELINE.json metric), self-validating.

Workload (BASELINE configs 2 + 5): one global batch of 8192 images, batch-sharded 8192/G per GPU
(megengine_b200.sharding.shard_range), every GPU running its shard as micro-batches of 256 images through
the 53 ConvBiasForward layers of ResNet-50 v1.5 (24 unique shapes; shapes =
dnn/test/cuda/conv_test_utils.cpp:23-55): NHWC, bias + ReLU (+ residual z on the last conv of each
bottleneck), enqueued in network order on one stream through the megdnn-style operator
(megengine_b200.ConvBiasForward.exec -> algorithm cache -> C-ABI b200cb_launch -> sm_100a kernels).
Synthetic inputs, random-init weights.  One "step" = one pass over the 8192-image global batch.

  value     images/s, every tensor resident in HBM (each layer owns private tensors: 13.5 GB are touched
            per micro-batch, so every layer starts L2-cold; conv1 reads the rank's slice of the real
            8192-image input, the last layer writes the rank's slice of the real output).
  e2e       the same job from HOST buffers: per micro-batch the un-padded 3-channel input is copied from
            pinned host memory, channel-padded on the device (b200cb_relayout), run through the 53 layers,
            and the final activations are copied back and read on the host -- all inside the timed region.
  roofline  dense-conv kernel (igemm) vs the measured bf16 tensor peak (MEASURED_PEAKS.json).
  parity    after timing: every one of the 53 layers at N=256 against the CPU oracle on >= 4096 sampled
            outputs, the first 2 images against an N=2 run (batch independence), and -- over NCCL -- the
            per-layer / per-micro-batch output checksums of every rank against rank 0's recomputation.
            A mismatch fails the bench (no JSON line, exit code 1).
  extra     BASELINE configs 3 + 4 and fp16 in the same line: QuantizedS8 NHWC / NCHW4 steps, fp16 step,
            MobileNetV2 + ShuffleNetV2 depthwise tables (bf16 / int8 / fp32), MobileNet pointwise tables.
  cpu_baseline / --impl reference
            the reference's own x86 megdnn ConvBias (oracle/_ref, compiled from /root/reference) on the
            host cores, fp32 NCHW, a bounded batch.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU; no
collective in the timed path; NCCL for the barrier, the max-reduce of the times and the validation gather).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BLOCKS = [3, 4, 6, 3]
GLOBAL_BATCH = 8192
MICRO_BATCH = 256


def network_order():
    """names of the 53 convs in execution order (keys of tests/_cases.RESNET50_LAYERS)."""
    order = ["conv1"]
    for s, nb in enumerate(BLOCKS, start=1):
        for b in range(nb):
            if b == 0:
                c1 = "l%d.b0.c1" % s
                c2 = "l1.c2" if s == 1 else "l%d.b0.c2" % s
                order += [c1, c2, "l%d.ds" % s, "l%d.c3" % s]
            else:
                c2 = "l1.c2" if s == 1 else "l%d.bN.c2" % s
                order += ["l%d.bN.c1" % s, c2, "l%d.c3" % s]
    assert len(order) == 53
    return order


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            pass
        self.t = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t is not None:
            self.t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(steps, warmup, budget_s, batch=None):
    """Times the reference's x86 megdnn ConvBias (oracle/_ref) on the 53 ResNet-50 layers, fp32 NCHW,
    bias+ReLU (the x86 fast path has no z input: dnn/src/fallback/conv_bias/opr_impl.cpp:534-549), all
    host threads.  Falls back to the C port of the naive operator when oracle/_ref is absent."""
    import numpy as np
    from megengine_b200 import _capi as C
    from tests import _cases as K, _oracle as O
    layers = {l[0]: l for l in K.RESNET50_LAYERS}
    order = network_order()
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    have_ref = O.have_ref()
    n = batch or (4 if have_ref else 1)
    uniq = {}
    for name in set(order):
        _nm, ic, h, w, oc, f, s, p, _z, _cnt = layers[name]
        d = K.make_desc("f32", n, ic, h, w, oc, f, stride=s, pad=p, fmt=C.FORMAT_NCHW,
                        nonline=C.NL_IDENTITY if name.endswith(".ds") else C.NL_RELU)
        rng = np.random.default_rng(len(uniq))
        src = rng.standard_normal(n * ic * h * w, dtype=np.float32)
        flt = rng.standard_normal(oc * ic * f * f, dtype=np.float32) * np.float32(np.sqrt(2.0 / (oc * f * f)))
        bias = rng.uniform(-0.5, 0.5, oc).astype(np.float32)
        uniq[name] = (d, src, flt, bias)
    flops_step = sum(2.0 * n * layers[nm][4] * uniq[nm][0].oh * uniq[nm][0].ow * layers[nm][1] * layers[nm][5] ** 2
                     for nm in order)

    def one_pass():
        t = 0.0
        for nm in order:
            d, src, flt, bias = uniq[nm]
            if have_ref:
                _dst, secs, _algo = O.ref_convbias(d, src, flt, bias, None, handle_kind=0, threads=cores, repeat=1)
                t += secs[0]
            else:
                t0 = time.perf_counter()
                O.oracle_convbias(d, src, flt, bias, None, threads=cores)
                t += time.perf_counter() - t0
        return t

    t_start = time.perf_counter()
    for _ in range(warmup if have_ref else 0):
        one_pass()
        if time.perf_counter() - t_start > budget_s:
            break
    times = []
    for _ in range(max(steps, 1)):
        times.append(one_pass())
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    t_med = sorted(times)[len(times) // 2]
    return {"value": n / t_med, "unit": "images/s", "cores": cores, "kind": "reference" if have_ref else "port",
            "sample": "%d pass(es) of the 53 ResNet-50 ConvBias layers at batch %d, fp32 NCHW, bias+ReLU, %s, "
                      "%d threads; median pass %.3f s (%.1f GFLOP/s)" % (
                          len(times), n, "megdnn x86 handle (oracle/_ref)" if have_ref else "C port of megdnn naive",
                          cores, t_med, flops_step / t_med / 1e9),
            "_steps": len(times), "_ms": t_med * 1e3, "_batch": n}


def run_reference_arm(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference(args.steps, min(args.warmup, 1), budget_s=150.0)
    line = {"impl": "reference", "metric": "resnet50_convbias_images_per_s", "value": cb["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": cb["_steps"], "warmup": min(args.warmup, 1), "ms_per_step": cb["_ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "resnet50_v1.5_53_convbias_layers", "format": "NCHW", "batch_per_step": cb["_batch"],
                       "note": "reference CPU implementation (megdnn x86/fallback ConvBias) on host cores; a bounded "
                               "sample (one batch-%d pass per step) of the B200 arm's 8192-image workload" % cb["_batch"]},
            "cpu_baseline": {k: v for k, v in cb.items() if not k.startswith("_")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def build_layers(kind, batch, device, conv1_pixel_bytes=8, fmt="nhwc"):
    """private tensors + operator for each of the 53 layer instances."""
    import torch
    import megengine_b200 as mb
    from megengine_b200 import _capi as C
    from tests import _cases as K
    layers = {l[0]: l for l in K.RESNET50_LAYERS}
    handle = mb.Handle(device)
    out = []
    tdt = {"bf16": torch.bfloat16, "f16": torch.float16, "qs8": torch.int8}[kind]
    g = torch.Generator(device=device)
    for idx, name in enumerate(network_order()):
        _nm, ic, h, w, oc, f, s, p, has_z, _cnt = layers[name]
        ic_real = ic
        if ic < 8:
            # conv1: IC 3 zero-padded (gopt PaddingChannelPass) to an 8- or 16-byte pixel -- bf16 IC 4 / 8,
            # int8 IC 8 / 16 (SURVEY 8d config 2 runs both); flops and algorithmic bytes count IC = 3
            ic = conv1_pixel_bytes // (1 if kind == "qs8" else 2)
        oh, ow = K.conv_out(h, f, s, p), K.conv_out(w, f, s, p)
        g.manual_seed(1000 + idx)
        opr = handle.create_operator()
        P = opr.param
        P.format, P.pad_h, P.pad_w, P.stride_h, P.stride_w = mb.Param.Format.NHWC, p, p, s, s
        if fmt == "nchw4":
            P.format = mb.Param.Format.NCHW4
        P.nonlineMode = mb.Param.NonlineMode.IDENTITY if name.endswith(".ds") else mb.Param.NonlineMode.RELU
        if kind == "qs8":
            sdt, fdt = mb.dtype.QuantizedS8(0.05), mb.dtype.QuantizedS8(0.02)
            bdt, ddt = mb.dtype.QuantizedS32(mb.mul_scale(sdt, fdt)), mb.dtype.QuantizedS8(0.11)
            zdt = mb.dtype.QuantizedS8(0.07)
            src = torch.randint(-127, 128, (batch, h, w, ic), dtype=torch.int8, device=device, generator=g)
            flt = torch.randint(-127, 128, (oc, f, f, ic), dtype=torch.int8, device=device, generator=g)
            bias = torch.randint(-5000, 5001, (1, 1, 1, oc), dtype=torch.int32, device=device, generator=g)
            z = torch.randint(-127, 128, (batch, oh, ow, oc), dtype=torch.int8, device=device, generator=g) if has_z else None
        else:
            P.compute_mode = mb.Param.ComputeMode.FLOAT32
            sdt = fdt = bdt = ddt = zdt = mb.dtype.BFloat16() if kind == "bf16" else mb.dtype.Float16()
            src = torch.randn((batch, h, w, ic), dtype=tdt, device=device, generator=g)
            flt = (torch.randn((oc, f, f, ic), dtype=torch.float32, device=device, generator=g)
                   * (2.0 / (oc * f * f)) ** 0.5).to(tdt)
            bias = (torch.rand((1, 1, 1, oc), dtype=torch.float32, device=device, generator=g) - 0.5).to(tdt)
            z = torch.randn((batch, oh, ow, oc), dtype=tdt, device=device, generator=g) if has_z else None
        if ic != ic_real:
            src[..., ic_real:] = 0
            flt[..., ic_real:] = 0
        dst = torch.empty((batch, oh, ow, oc), dtype=tdt, device=device)
        if fmt == "nchw4":
            # QuantizedS8 NCHW4 (BASELINE config 3): the same logical tensors as (N, C/4, H, W, 4)
            to4 = lambda t: t.reshape(t.shape[0], t.shape[1], t.shape[2], t.shape[3] // 4, 4).permute(0, 3, 1, 2, 4).contiguous()
            src, flt, dst = to4(src), to4(flt), to4(dst)
            z = to4(z) if has_z else None
            bias = bias.reshape(1, oc // 4, 1, 1, 4)
        S, F, B, D = mb.make_tensor(src, sdt), mb.make_tensor(flt, fdt), mb.make_tensor(bias, bdt), mb.make_tensor(dst, ddt)
        Z = mb.make_tensor(z, zdt) if has_z else mb.empty_like_dtype(zdt)
        plan = opr.get_plan(S.layout, F.layout, B.layout, Z.layout, D.layout)
        ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=device) if plan.workspace_bytes else None
        flops = 2.0 * batch * oc * oh * ow * ic_real * f * f
        es = src.element_size()
        # algorithmic bytes count the un-padded conv1 input / filter
        abytes = float(plan.bytes) - (ic - ic_real) * (batch * h * w + oc * f * f) * es
        out.append({"name": name, "opr": opr, "args": (S, F, B, Z, D), "plan": plan, "ws": ws, "flops": flops,
                    "bytes": abytes, "algo": C.load().b200cb_algo_name(plan.algo).decode(),
                    "launches": C.load().b200cb_plan_num_launches(ctypes.byref(plan)),
                    "geom": (ic, ic_real, h, w, oc, f, s, p, bool(has_z)), "kind": kind})
    return handle, out


def run_microbatch(layers, first_src=None, last_dst=None):
    """the 53 layers through the operator's public exec (algorithm cache, no pre-computed plan)."""
    n = len(layers)
    for i, L in enumerate(layers):
        S, F, B, Z, D = L["args"]
        if i == 0 and first_src is not None:
            S = first_src
        if i == n - 1 and last_dst is not None:
            D = last_dst
        L["opr"].exec(S, F, B, Z, D, None, L["ws"])


def per_layer_report(layers, peaks, reps=5, peak_mul=1.0):
    """CUDA-event time of every layer inside the running network (warm pipeline, L2-cold data)."""
    import torch
    n = len(layers)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
          for _ in range(reps)]
    for r in range(reps):
        for i, L in enumerate(layers):
            S, F, B, Z, D = L["args"]
            ev[r][i][0].record()
            L["opr"].exec(S, F, B, Z, D, None, L["ws"])
            ev[r][i][1].record()
    torch.cuda.synchronize()
    rows = {}
    for i, L in enumerate(layers):
        ms = sorted(ev[r][i][0].elapsed_time(ev[r][i][1]) for r in range(reps))[reps // 2]
        row = rows.setdefault(L["name"], {"name": L["name"], "algo": L["algo"], "count": 0, "ms": 0.0,
                                          "flops": L["flops"], "bytes": L["bytes"]})
        row["count"] += 1
        row["ms"] += ms
    out = []
    tp = peaks["tf_burst"] * peak_mul
    for row in rows.values():
        ms = row["ms"] / row["count"]
        tf = row["flops"] / ms / 1e9
        gbs = row["bytes"] / ms / 1e6
        t_floor = max(row["flops"] / (tp * 1e12), row["bytes"] / (peaks["hbm_gbs"] * 1e9)) * 1e3
        out.append({"name": row["name"], "algo": row["algo"], "count": row["count"], "ms": round(ms, 4),
                    "tflops": round(tf, 1), "gbs": round(gbs, 1),
                    "frac_tensor": round(tf / tp, 3), "frac_hbm": round(gbs / peaks["hbm_gbs"], 3),
                    "bound": "tensor" if row["flops"] / (tp * 1e12) > row["bytes"] / (peaks["hbm_gbs"] * 1e9) else "hbm",
                    "floor_ms": round(t_floor, 4), "frac_of_floor": round(t_floor / ms, 3)})
    return out


# ------------------------------------------------------------------------------ validation
def _sampled_oracle_check(L, src_t, dst_t, nsamples, seed):
    """>= nsamples random outputs of one layer instance at its full batch against the CPU oracle.  The
    receptive fields of the samples are gathered on the GPU, so that only S x FH x FW x IC inputs travel
    to the host: sample s becomes image s of a 1x1-output convolution, of which the oracle computes the one
    sampled channel (tests/_oracle.oracle_convbias_sample).  int8: bit-exact; bf16 / fp16: fp32 oracle,
    tolerance 1e-2 / 2e-3 on |x-y|/max(|x|,|y|,1) (north_star; metric dnn/test/common/utils.h:189-193)."""
    import numpy as np
    import torch
    from megengine_b200 import _capi as C
    from tests import _cases as K, _oracle as O
    ic, _icr, h, w, oc, f, s, p, has_z = L["geom"]
    kind = L["kind"]
    S_, F_, B_, Z_, _D = L["args"]
    n = src_t.shape[0]
    oh, ow = dst_t.shape[1], dst_t.shape[2]
    g = torch.Generator(device=src_t.device)
    g.manual_seed(seed)
    total = n * oh * ow * oc
    idx = torch.randint(0, total, (nsamples,), device=src_t.device, generator=g, dtype=torch.int64)
    c_i = idx % oc
    pix = idx // oc
    ow_i, oh_i, n_i = pix % ow, (pix // ow) % oh, pix // (ow * oh)
    ar = torch.arange(f, device=src_t.device)
    ih = (oh_i * s - p)[:, None] + ar[None, :]
    iw = (ow_i * s - p)[:, None] + ar[None, :]
    okh, okw = (ih >= 0) & (ih < h), (iw >= 0) & (iw < w)
    win = src_t[n_i[:, None, None], ih.clamp(0, h - 1)[:, :, None], iw.clamp(0, w - 1)[:, None, :], :]
    win = win * (okh[:, :, None] & okw[:, None, :])[..., None].to(win.dtype)
    got = dst_t.reshape(-1)[idx]
    zs = Z_.tensor[n_i, oh_i, ow_i, :] if has_z else None
    nl = C.NL_IDENTITY if L["name"].endswith(".ds") else C.NL_RELU
    sidx = (torch.arange(nsamples, device=src_t.device) * oc + c_i).cpu().numpy().astype(np.uint64)
    if kind == "qs8":
        sd, fd, dd = S_.layout.dtype, F_.layout.dtype, _D.layout.dtype
        d = K.make_desc("qs8", nsamples, ic, f, f, oc, f, nonline=nl, has_z=has_z,
                        scales=(sd.scale, fd.scale, B_.layout.dtype.scale, Z_.layout.dtype.scale if has_z else 0.0, dd.scale))
        want = O.oracle_convbias_sample(
            d, win.contiguous().cpu().numpy().reshape(-1), F_.tensor.cpu().numpy().reshape(-1),
            B_.tensor.cpu().numpy().reshape(-1), zs.contiguous().cpu().numpy().reshape(-1) if has_z else None, sidx)
        bad = int((got.cpu().numpy() != want).sum())
        return {"mismatch": bad, "ok": bad == 0}
    d = K.make_desc("f32", nsamples, ic, f, f, oc, f, nonline=nl, has_z=has_z)
    f32 = lambda t: t.float().contiguous().cpu().numpy().reshape(-1)
    want = O.oracle_convbias_sample(d, f32(win), f32(F_.tensor), f32(B_.tensor), f32(zs) if has_z else None, sidx)
    err = O.megdnn_rel_err(got.float().cpu().numpy(), want)
    tol = 1e-2 if kind == "bf16" else 2e-3
    return {"max_rel": float(err.max()), "tol": tol, "ok": bool(err.max() <= tol)}


def _slice_check(L, handle, src_t, dst_t):
    """SURVEY 8(c) protocol ii: the full-batch output restricted to the first 2 images equals an N=2 run of
    the same operator on the first 2 images (batch independence: tile schedule, CTA pairs, wave tails,
    phantom tiles and the multi-image depthwise tiles all differ between N=2 and N=256)."""
    import torch
    import megengine_b200 as mb
    S_, F_, B_, Z_, D_ = L["args"]
    opr = handle.create_operator()
    opr.param = L["opr"].param
    s2 = mb.make_tensor(src_t[:2].contiguous(), S_.layout.dtype)
    d2 = mb.make_tensor(torch.empty_like(dst_t[:2]), D_.layout.dtype)
    z2 = mb.make_tensor(Z_.tensor[:2].contiguous(), Z_.layout.dtype) if Z_.layout.ndim else Z_
    ws_b = opr.get_workspace_in_bytes(s2.layout, F_.layout, B_.layout, z2.layout, d2.layout)
    ws = torch.empty(ws_b, dtype=torch.uint8, device=src_t.device) if ws_b else None
    opr.exec(s2, F_, B_, z2, d2, None, ws)
    torch.cuda.synchronize()
    view = lambda t: t.view(torch.int16) if t.element_size() == 2 else t
    return bool(torch.equal(view(d2.tensor), view(dst_t[:2].contiguous())))


def validate_layers(handle, layers, first_src, last_dst, nsamples=4096):
    """every layer instance of the step, at the timed batch, against the oracle (i) and the N=2 slice (ii)."""
    import torch
    torch.cuda.synchronize()
    res = {"layers": 0, "samples_per_layer": nsamples, "failed": [], "max_rel": 0.0, "slice_mismatch": []}
    n = len(layers)
    for i, L in enumerate(layers):
        S_, _F, _B, _Z, D_ = L["args"]
        src_t = first_src.tensor if (i == 0 and first_src is not None) else S_.tensor
        dst_t = last_dst.tensor if (i == n - 1 and last_dst is not None) else D_.tensor
        r = _sampled_oracle_check(L, src_t, dst_t, nsamples, seed=7000 + i)
        res["layers"] += 1
        if "max_rel" in r:
            res["max_rel"] = max(res["max_rel"], r["max_rel"])
        if not r["ok"]:
            res["failed"].append({"layer": L["name"], "index": i, **r})
        if not _slice_check(L, handle, src_t, dst_t):
            res["slice_mismatch"].append(L["name"])
    res["ok"] = not res["failed"] and not res["slice_mismatch"]
    return res


# ------------------------------------------------------------------------------ extras (configs 3, 4, fp16)
def time_config(kind, fmt, device, steps=10, warmup=3):
    """one GPU, one 256-image micro-batch of the 53 layers per step, tensors resident."""
    import torch
    _h, layers = build_layers(kind, MICRO_BATCH, device, 8, fmt)
    for _ in range(warmup):
        run_microbatch(layers)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run_microbatch(layers)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = sum(L["flops"] for L in layers)
    byts = sum(L["bytes"] for L in layers)
    out = {"ms_per_256": round(ms, 4), "images_per_s": round(MICRO_BATCH / ms * 1e3, 1),
           ("tops" if kind == "qs8" else "tflops"): round(flops / ms / 1e9, 1), "gbs": round(byts / ms / 1e6, 1),
           "launches_per_256": sum(L["launches"] for L in layers)}
    del layers
    torch.cuda.empty_cache()
    return out


def measure_int8_peak():
    """dense int8 GEMM throughput of this GPU measured like MEASURED_PEAKS.json's bf16 figure (library GEMM
    8192^3, best of 10) -- the denominator of the QuantizedS8 roofline fractions."""
    import torch
    try:
        a = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device="cuda")
        b = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device="cuda")
        torch._int_mm(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192 ** 3 / best / 1e9
    except Exception:  # noqa: BLE001
        return None


def layer_tables(peaks):
    """BASELINE config 4: MobileNetV2 + ShuffleNetV2 depthwise 3x3 and MobileNet pointwise 1x1, N = 256,
    every layer timed L2-cold (rotating replicas > 256 MB, tools/bench_layers.py); GB/s = algorithmic bytes
    (src + filter + bias + z + dst; the reference's dnn/test/cuda/chanwise_convolution.cpp:797-801
    accounting plus the fused bias) / time."""
    from megengine_b200 import _capi as C
    from tests import _cases as K
    from tools import bench_layers as BL
    lib = C.load()
    pk = {"hbm_gbs": peaks["hbm_gbs"], "tf": peaks["tf_burst"]}
    out = {}
    for dt in ("bf16", "qs8", "f32"):
        rows, nets = [], {"mobilenetv2": [], "shufflenetv2": []}
        for net, table in (("mobilenetv2", K.MOBILENETV2_DW), ("shufflenetv2", BL.SHUFFLENET_DW)):
            for c, h, w, s in table:
                if (dt == "qs8" and c % 16) or (dt == "f32" and c % 4):
                    continue
                r = BL.time_desc(lib, K.dw_desc(dt, MICRO_BATCH, c, h, w, s), 30, pk)
                if "ms" in r:
                    r["shape"] = "c%d %dx%d s%d" % (c, h, w, s)
                    rows.append(r)
                    nets[net].append(r)
        agg = lambda rs: {"layers": len(rs), "ms": round(sum(r["ms"] for r in rs), 4),
                          "gbs": round(sum(r["bytes"] for r in rs) / sum(r["ms"] for r in rs) / 1e6, 1),
                          "frac_hbm": round(sum(r["bytes"] for r in rs) / sum(r["ms"] for r in rs) / 1e6 / peaks["hbm_gbs"], 3)}
        out["depthwise_" + dt] = dict(agg(rows), **{k: agg(v) for k, v in nets.items() if v},
                                      rows=[{"shape": r["shape"], "ms": r["ms"], "frac_hbm": r["frac_hbm"]} for r in rows])
    for dt in ("bf16", "qs8"):
        rows = []
        for oc, ic, h, w in BL.MOBILENET_PW:
            if dt == "qs8" and (ic % 16 or oc % 16):
                continue
            r = BL.time_desc(lib, K.make_desc(dt, MICRO_BATCH, ic, h, w, oc, 1, nonline=C.NL_RELU), 30, pk)
            if "ms" in r:
                rows.append(dict(r, shape="%d->%d %dx%d" % (ic, oc, h, w)))
        out["pointwise_" + dt] = {
            "layers": len(rows), "ms": round(sum(r["ms"] for r in rows), 4),
            "gbs": round(sum(r["bytes"] for r in rows) / sum(r["ms"] for r in rows) / 1e6, 1),
            "frac_hbm": round(sum(r["bytes"] for r in rows) / sum(r["ms"] for r in rows) / 1e6 / peaks["hbm_gbs"], 3),
            "rows": [{"shape": r["shape"], "ms": r["ms"], "frac_hbm": r["frac_hbm"]} for r in rows]}
    return out


def small_batch_latency(device):
    """Serving-style latency of the 53-layer stack at batch 1 and 8 (bf16 NHWC): one micro-batch enqueued layer by
    layer through the operator API (launch bound: ~11 us of host time per layer) against the same 54 launches
    replayed as ONE captured CUDA graph."""
    import torch
    out = {}
    for n in (1, 8):
        _h, layers = build_layers("bf16", n, device)
        for _ in range(3):
            run_microbatch(layers)
        torch.cuda.synchronize()

        def timed(fn, reps=30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        eager = timed(lambda: run_microbatch(layers))
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run_microbatch(layers)
            g.capture_begin()
            run_microbatch(layers)
            g.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = timed(g.replay)
        out["batch%d" % n] = {"eager_ms": round(eager, 4), "cuda_graph_ms": round(graph, 4),
                              "images_per_s_graph": round(n / graph * 1e3, 1)}
        del g, layers
        torch.cuda.empty_cache()
    return out


def config1_table(peaks, cpu=True):
    """BASELINE config 1 (`configs[0]`, the reference's own CPU-runnable case): the ResNet-18 conv stack, fp32
    NCHW, batch 8, 224x224, bias + ReLU.  GPU: every layer through the Float32 NCHW cell of the C-ABI (bf16 x 3
    operand split), L2-cold, summed in network order; CPU: the reference's x86 megdnn ConvBias (oracle/_ref) on
    1 thread and on all host threads, 1 warm-up + 2 timed passes per layer."""
    from megengine_b200 import _capi as C
    from tests import _cases as K
    from tools import bench_layers as BL
    lib = C.load()
    pk = {"hbm_gbs": peaks["hbm_gbs"], "tf": peaks["tf_burst"]}
    out = {"config": "ResNet-18 conv stack, fp32 NCHW, batch 8, 224x224, bias+ReLU"}
    for n in (8, MICRO_BATCH):
        ms, fl, bad = 0.0, 0.0, []
        for layer in K.RESNET18_LAYERS:
            r = BL.time_desc(lib, K.resnet18_desc(layer, n), 20, pk)
            if "ms" not in r:
                bad.append(layer[0])
                continue
            ms += r["ms"] * layer[-1]
            fl += r["flops"] * layer[-1]
        out["b200_batch%d" % n] = {"stack_ms": round(ms, 4), "images_per_s": round(n / ms * 1e3, 1),
                                   "tflops": round(fl / ms / 1e9, 1), "unsupported_layers": bad,
                                   "tolerance": "1e-3 max rel vs the fp32 oracle (tests/test_parity_gpu.py, every layer)"}
    if cpu:
        try:
            from tests import _oracle as O
            if O.have_ref():
                import numpy as np
                cores = len(os.sched_getaffinity(0))
                for threads in sorted({1, cores}):
                    tot, fl = 0.0, 0.0
                    for layer in K.RESNET18_LAYERS:
                        d = K.resnet18_desc(layer, 8)
                        src, flt, bias, _ = K.gen_inputs(d, 1)
                        _dst, secs, _algo = O.ref_convbias(d, src, flt, bias, None, handle_kind=0, threads=threads, repeat=3)
                        t = float(np.median(list(secs)[1:]))
                        tot += t * layer[-1]
                        fl += 2.0 * d.n * d.oc * d.oh * d.ow * d.ic * d.fh * d.fw * layer[-1]
                    out["reference_x86_%d_threads" % threads] = {"stack_ms": round(tot * 1e3, 2),
                                                                 "images_per_s": round(8 / tot, 2),
                                                                 "gflops": round(fl / tot / 1e9, 1)}
        except Exception as e:  # noqa: BLE001 -- the CPU leg is a reported baseline, never a reason to fail the bench
            out["reference_error"] = str(e)[:200]
    return out


def pinned_input(shape, dtype, kind):
    """host staging buffer of the end-to-end input.  'wc': page-locked write-combined memory
    (cudaHostAllocWriteCombined: the CPU only ever writes it, device DMA reads skip the cache snoops)."""
    import torch
    if kind == "wc":
        try:
            n = 1
            for d in shape:
                n *= d
            nbytes = n * torch.empty((), dtype=dtype).element_size()
            rt = ctypes.CDLL("libcudart.so.12")
            ptr = ctypes.c_void_p()
            if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01)) == 0:
                buf = (ctypes.c_char * nbytes).from_address(ptr.value)
                t = torch.frombuffer(buf, dtype=torch.uint8).view(dtype).reshape(shape)
                if t.is_pinned():
                    return t, "write-combined pinned (cudaHostAllocWriteCombined)"
        except Exception:  # noqa: BLE001
            pass
    return torch.empty(shape, dtype=dtype).pin_memory(), "pinned"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5, help="passes over the 8192-image global batch")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16", "qs8"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH, help="images per step over all GPUs")
    ap.add_argument("--format", default="nhwc", choices=["nhwc", "nchw4"], help="nchw4: QuantizedS8 only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-input", default="pinned", choices=["wc", "pinned"],
                    help="host staging memory of the e2e input: write-combined pinned (cudaHostAllocWriteCombined) or pinned")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3 / 4 / fp16 sections")
    ap.add_argument("--no-validate", action="store_true", help="skip the parity checks (tools only)")
    ap.add_argument("--conv1-pixel-bytes", type=int, default=8, choices=[8, 16],
                    help="channel padding of the 3-channel network input: 8 bytes (bf16 IC=4 / int8 IC=8) or 16")
    ap.add_argument("--layer-report", default="", help="write the per-layer roofline table (JSON) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # stdout carries exactly one JSON line: library chatter (e.g. NCCL's version banner, printed at
    # communicator creation) is routed to stderr until that line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args, emit)
        return

    import torch
    import torch.distributed as dist
    import megengine_b200 as mb
    from megengine_b200 import _capi as C
    from megengine_b200 import sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    # bind the rank to the CPUs next to its GPU before any pinned allocation (first touch puts the pinned
    # staging buffers on the GPU's NUMA node): the 8-GPU end-to-end pipeline is bound by host memory traffic
    affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        affinity = "%d cpus (nvmlDeviceSetCpuAffinity)" % len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
    peaks = load_peaks()
    lib = C.load()

    if args.format == "nchw4" and args.dtype != "qs8":
        raise SystemExit("bench.py: --format nchw4 is the QuantizedS8 configuration (use --dtype qs8)")
    nhwc = args.format == "nhwc"
    handle, layers = build_layers(args.dtype, MICRO_BATCH, device, args.conv1_pixel_bytes, args.format)
    flops_mb = sum(L["flops"] for L in layers)
    bytes_mb = sum(L["bytes"] for L in layers)
    launches_mb = sum(L["launches"] for L in layers)
    n_layers = len(layers)
    peak_mul = 2.0 if args.dtype == "qs8" else 1.0
    floor_ms = sum(max(L["flops"] / (peaks["tf_burst"] * peak_mul * 1e12), L["bytes"] / (peaks["hbm_gbs"] * 1e9))
                   for L in layers) * 1e3

    # ---------------- the sharded global batch (config 5): rank g owns images [g*N/G, (g+1)*N/G)
    first, count = sharding.shard_range(args.global_batch, world, rank)
    if count % MICRO_BATCH or first % MICRO_BATCH:
        raise SystemExit("bench.py: the per-GPU shard (%d images) must be a multiple of the %d-image micro-batch"
                         % (count, MICRO_BATCH))
    mbs = count // MICRO_BATCH
    L0, LN = layers[0], layers[-1]
    S0, DN = L0["args"][0], LN["args"][4]
    ic_pad, ic_real = L0["geom"][0], L0["geom"][1]

    def gen_input(global_mb):
        """the un-padded 3-channel images of global micro-batch `global_mb` (seeded by its index)."""
        g = torch.Generator(device=device)
        g.manual_seed(5000 + global_mb)
        shape = (MICRO_BATCH, 224, 224, ic_real)
        if args.dtype == "qs8":
            return torch.randint(-127, 128, shape, dtype=torch.int8, device=device, generator=g)
        return torch.randn(shape, dtype=S0.tensor.dtype, device=device, generator=g)

    x_shard = y_shard = None
    src_views, dst_views = [None] * mbs, [None] * mbs
    if nhwc:
        x_shard = torch.zeros((count,) + tuple(S0.tensor.shape[1:]), dtype=S0.tensor.dtype, device=device)
        for j in range(mbs):
            x_shard[j * MICRO_BATCH:(j + 1) * MICRO_BATCH, ..., :ic_real] = gen_input(first // MICRO_BATCH + j)
        y_shard = torch.empty((count,) + tuple(DN.tensor.shape[1:]), dtype=DN.tensor.dtype, device=device)
        src_views = [mb.make_tensor(x_shard[j * MICRO_BATCH:(j + 1) * MICRO_BATCH], S0.layout.dtype) for j in range(mbs)]
        dst_views = [mb.make_tensor(y_shard[j * MICRO_BATCH:(j + 1) * MICRO_BATCH], DN.layout.dtype) for j in range(mbs)]

    def run_step():
        for j in range(mbs):
            run_microbatch(layers, src_views[j], dst_views[j])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        run_step()
    barrier()

    # ---------------- timed region 1: tensors resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    images_s = args.global_batch / (ms_per_step * 1e-3)
    ms_per_mb = ms_per_step / mbs
    # host cost of one exec() (enqueue only): one micro-batch = 53 launches never fills the launch queue
    t_host = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_microbatch(layers, src_views[0], dst_views[0])
        t_host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    host_us_per_exec = sorted(t_host)[1] * 1e6 / n_layers

    # ---------------- timed region 2: end to end from host buffers
    # The serving pipeline a user of the operators runs.  Per micro-batch, inside the timed region: the 3-channel
    # int8 image batch is copied from pinned host memory (H2D, 38.5 MB), converted on the device to the network's
    # dtype (b200cb_typecvt: QuantizedS8(1/64) -> BFloat16 / Float16, exact; the QuantizedS8 network takes it as is),
    # zero-padded to the stem's channel count (b200cb_relayout NHWC_PAD_C -- gopt's channel padding as a first-class
    # op), run through the 53 layers, reduced by the network's next operator -- the global average pool
    # (b200cb_pooling) -- and the pooled features (256 x 2048) are copied back (D2H) and read on the host.  Double
    # buffered on copy streams: the H2D of micro-batch i+1 and the D2H of i-1 overlap the layers of i.  (Round 1
    # shipped bf16 images and read back the whole 7x7x2048 activations, 154 MB per micro-batch: at 8 GPUs that
    # pipeline was bound by the host's aggregate DMA bandwidth, 139 GB/s, not by any GPU.)
    e2e = None
    if nhwc:
        es = S0.tensor.element_size()
        img_scale = 1.0 / 64.0
        h_in, h_in_kind = pinned_input((MICRO_BATCH, 224, 224, ic_real), torch.int8, args.host_input)
        x0 = gen_input(first // MICRO_BATCH)
        img = x0 if args.dtype == "qs8" else torch.clamp(torch.round(x0.float() * 32.0), -128, 127).to(torch.int8)
        h_in.copy_(img.cpu())
        del x0, img
        to_net = args.dtype != "qs8"                           # int8 image -> 16-bit network input on the device
        net_dt = {"bf16": C.DT_BF16, "f16": C.DT_F16, "qs8": C.DT_QS8}[args.dtype]
        d_raw = [torch.empty_like(h_in, device=device) for _ in range(2)]
        d_cvt = [torch.empty(h_in.shape, dtype=S0.tensor.dtype, device=device) for _ in range(2)] if to_net else d_raw
        d_in = [mb.make_tensor(torch.zeros_like(S0.tensor), S0.layout.dtype) for _ in range(2)]
        d_out = [mb.make_tensor(torch.empty_like(DN.tensor), DN.layout.dtype) for _ in range(2)]
        n_img, fh, fw, fc = DN.tensor.shape
        d_feat = [torch.empty((n_img, 1, 1, fc), dtype=DN.tensor.dtype, device=device) for _ in range(2)]
        h_out = [torch.empty((n_img, 1, 1, fc), dtype=DN.tensor.dtype).pin_memory() for _ in range(2)]
        pd = C.PoolDesc()
        pd.format, pd.mode, pd.dtype = C.FORMAT_NHWC, C.POOL_AVERAGE, net_dt
        pd.n, pd.c, pd.ih, pd.iw, pd.oh, pd.ow = n_img, fc, fh, fw, 1, 1
        pd.window_h, pd.window_w, pd.stride_h, pd.stride_w, pd.pad_h, pd.pad_w = fh, fw, fh, fw, 0, 0
        comp = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        pixels = MICRO_BATCH * 224 * 224

        def enqueue(i):
            b = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[b])                 # micro-batch i-2 no longer reads this buffer
                d_raw[b].copy_(h_in, non_blocking=True)
                ev_in[b].record(s_in)
            comp.wait_event(ev_in[b])
            comp.wait_event(ev_out[b])                      # micro-batch i-2's result has left d_feat[b]
            if to_net:
                rc = lib.b200cb_typecvt(C.DT_QS8, img_scale, net_dt, 0.0, pixels * ic_real, d_raw[b].data_ptr(),
                                        d_cvt[b].data_ptr(), comp.cuda_stream)
                assert rc == 0, C.strerror(rc)
            rc = lib.b200cb_relayout(C.RELAYOUT_NHWC_PAD_C, es, 1, ic_real, 1, pixels, ic_pad,
                                     d_cvt[b].data_ptr(), d_in[b].tensor.data_ptr(), comp.cuda_stream)
            assert rc == 0, C.strerror(rc)
            ev_free[b].record(comp)
            run_microbatch(layers, d_in[b], d_out[b])
            rc = lib.b200cb_pooling(ctypes.byref(pd), d_out[b].tensor.data_ptr(), d_feat[b].data_ptr(), comp.cuda_stream)
            assert rc == 0, C.strerror(rc)
            ev_done[b].record(comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                h_out[b].copy_(d_feat[b], non_blocking=True)
                ev_out[b].record(s_out)

        e2e_steps = max(1, min(args.steps, 3))
        n_mb = e2e_steps * mbs
        for b in range(2):
            ev_free[b].record(comp)
            ev_out[b].record(comp)
        for i in range(2):                                  # warm the pipeline
            enqueue(i)
        barrier()
        t0 = time.perf_counter()
        checksum = 0.0
        for i in range(n_mb):
            enqueue(i)
            if i >= 1:
                ev_out[(i - 1) & 1].synchronize()           # the user reads micro-batch i-1's result
                checksum += float(h_out[(i - 1) & 1].view(-1)[0])
        ev_out[(n_mb - 1) & 1].synchronize()
        checksum += float(h_out[(n_mb - 1) & 1].view(-1)[0])
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        h_in_bytes = h_in.numel() * h_in.element_size()
        h_out_bytes = h_out[0].numel() * h_out[0].element_size()
        # the result that reached the host is the pooling of the activations the last micro-batch left on the
        # device: 2 images against the oracle's pooling (test infrastructure, after the timed region), bit for bit
        result_check = None
        if not args.no_validate and rank == 0:
            try:
                from tests import _nn as N
                lb = (n_mb - 1) & 1
                torch.cuda.synchronize()
                k = 2
                act = d_out[lb].tensor[:k].contiguous().cpu()
                raw = act.view(torch.int16).numpy().view("uint16") if es == 2 else act.numpy()
                od = N.pool_desc(C.FORMAT_NHWC, C.POOL_AVERAGE, net_dt, k, fc, fh, fw, (fh, fw), (fh, fw), 0)
                want = N.pooling("oracle", od, raw.reshape(-1))
                got_t = h_out[lb][:k].contiguous()
                got = got_t.view(torch.int16).numpy().view("uint16") if es == 2 else got_t.numpy()
                bad = int((got.reshape(-1) != want).sum())
                result_check = {"images": k, "mismatching_elements": bad}
                if bad:
                    sys.stderr.write("bench.py: e2e result check FAILED: %d pooled features differ from the oracle\n" % bad)
                    sys.exit(1)
            except ImportError:
                result_check = {"skipped": "oracle not importable"}
        e2e = {"value": args.global_batch * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
               "h2d_bytes_per_step": h_in_bytes * mbs * world, "d2h_bytes_per_step": h_out_bytes * mbs * world,
               "steps": e2e_steps, "h2d_bytes_per_gpu_microbatch": h_in_bytes, "d2h_bytes_per_gpu_microbatch": h_out_bytes,
               "host_input_memory": h_in_kind, "host_input": "int8 images (256 x 224 x 224 x 3), scale 1/64",
               "result": "global-average-pooled features (256 x %d) of the last layer's activations" % fc,
               "result_check_vs_oracle": result_check,
               "host_copy_gbs_all_gpus": (h_in_bytes + h_out_bytes) * mbs * world * e2e_steps / (e2e_ms * 1e-3) / 1e9,
               "pipeline": "per 256-image micro-batch: H2D of the 3-channel int8 images, device TypeCvt to the network dtype "
                           "(b200cb_typecvt) and channel padding (b200cb_relayout), 53 ConvBias layers, global average pool "
                           "(b200cb_pooling), D2H of the pooled features; double-buffered on copy streams"}
        del d_raw, d_cvt, d_in, d_out, d_feat, h_out, h_in

    # ---------------- parity (after the timed regions; never inside them)
    parity = None
    if not args.no_validate and nhwc:
        torch.cuda.synchronize()
        run_microbatch(layers, src_views[mbs - 1], dst_views[mbs - 1])
        parity = validate_layers(handle, layers, src_views[mbs - 1], dst_views[mbs - 1])
        # cross-GPU: per-layer output checksums (every rank ran the same 53 layers from the same seeds) and
        # the conv1 output checksum of every micro-batch of the sharded 8192-image batch, gathered over NCCL;
        # rank 0 recomputes the micro-batches of the other ranks from their seeds and compares
        layer_sums = [sharding.checksum((dst_views[mbs - 1] if i == len(layers) - 1 else L["args"][4]).tensor[:8])
                      for i, L in enumerate(layers)]
        mb_sums = []
        for j in range(mbs):
            L0["opr"].exec(src_views[j], *L0["args"][1:], None, L0["ws"])
            mb_sums.append(sharding.checksum(L0["args"][4].tensor[:16]))
        mine = torch.tensor(layer_sums + mb_sums, dtype=torch.int64, device=device)
        if world > 1:
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            tail = sharding.gather_batch(L0["args"][4].tensor[:2].contiguous(), 2 * world)
        else:
            allv, tail = [mine], L0["args"][4].tensor[:2].contiguous()
        shard = {"global_batch": args.global_batch, "images_per_gpu": count, "microbatches_per_gpu": mbs,
                 "checksums_gathered": world * (len(layers) + mbs), "collective": "nccl all_gather" if world > 1 else "none (1 GPU)"}
        if rank == 0:
            nl = len(layers)
            # (layer 0 reads the rank's own images: its outputs are checked per micro-batch below; layers
            # 1..52 run on the same seeded private tensors on every rank and must agree bit for bit)
            shard["layer_checksums_equal_across_ranks"] = all(bool(torch.equal(v[1:nl], allv[0][1:nl])) for v in allv)
            bad = []
            for r in range(world):
                r_first = sharding.shard_range(args.global_batch, world, r)[0] // MICRO_BATCH
                for j in range(mbs):
                    xin = torch.zeros_like(S0.tensor)
                    xin[..., :ic_real] = gen_input(r_first + j)
                    L0["opr"].exec(mb.make_tensor(xin, S0.layout.dtype), *L0["args"][1:], None, L0["ws"])
                    if sharding.checksum(L0["args"][4].tensor[:16]) != int(allv[r][nl + j].item()):
                        bad.append((r, j))
                    if j == mbs - 1 and not torch.equal(L0["args"][4].tensor[:2], tail[2 * r:2 * r + 2]):
                        bad.append((r, "gathered rows"))
            shard["microbatch_checksums_verified_on_rank0"] = world * mbs
            shard["microbatch_checksum_mismatches"] = bad
            parity["sharding"] = shard
            parity["ok"] = parity["ok"] and shard["layer_checksums_equal_across_ranks"] and not bad
        okt = torch.tensor([1 if parity["ok"] else 0], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)  # every rank validated its own 53 layers
        if int(okt.item()) != 1:
            sys.stderr.write("bench.py: PARITY FAILED (rank %d): %s\n" % (rank, json.dumps(parity)[:4000]))
            sys.stderr.flush()
            if world > 1:
                dist.destroy_process_group()
            sys.exit(1)

    report = None
    if rank == 0 and args.layer_report:
        report = per_layer_report(layers, peaks, peak_mul=2.0 if args.dtype == "qs8" else 1.0)
        os.makedirs(os.path.dirname(os.path.abspath(args.layer_report)), exist_ok=True)
        with open(args.layer_report, "w") as f:
            json.dump({"dtype": args.dtype, "batch": MICRO_BATCH, "peaks": peaks, "layers": report}, f, indent=1)

    # ---------------- the other configurations of BASELINE.json, same run, same line (rank 0, 1 GPU)
    extra = None
    if rank == 0 and world == 1 and not args.no_extras and nhwc and args.dtype == "bf16":
        rows = report if report is not None else per_layer_report(layers, peaks, reps=3)
        del layers, x_shard, y_shard, src_views, dst_views
        torch.cuda.empty_cache()
        int8_peak = measure_int8_peak()
        p8 = int8_peak if int8_peak else 2.0 * peaks["tf_burst"]
        extra = {"int8_peak_tops": int8_peak, "int8_peak_source":
                 "measured: torch._int_mm 8192^3, best of 10" if int8_peak else "assumed 2 x measured bf16 burst",
                 "fp16_nhwc": time_config("f16", "nhwc", device),
                 "qs8_nhwc": time_config("qs8", "nhwc", device),
                 "qs8_nchw4": time_config("qs8", "nchw4", device)}
        for k in ("qs8_nhwc", "qs8_nchw4"):
            extra[k]["frac_of_int8_peak"] = round(extra[k]["tops"] / p8, 4)
            extra[k]["bit_exact"] = "0 mismatching bytes vs the oracle in tests/ (-m gpu) and in this run's parity section" \
                if k == "qs8_nhwc" else "0 mismatching bytes vs the oracle in tests/ (-m gpu)"
        extra["fp16_nhwc"]["frac_of_bf16_peak"] = round(extra["fp16_nhwc"]["tflops"] / peaks["tf_burst"], 4)
        # int8 parity at the timed size, same protocol as the bf16 step
        if not args.no_validate:
            h8, l8 = build_layers("qs8", MICRO_BATCH, device, 8, "nhwc")
            run_microbatch(l8)
            p8r = validate_layers(h8, l8, None, None)
            extra["qs8_nhwc"]["parity"] = {k: p8r[k] for k in ("ok", "layers", "samples_per_layer", "failed", "slice_mismatch")}
            del l8
            torch.cuda.empty_cache()
            if not p8r["ok"]:
                sys.stderr.write("bench.py: PARITY FAILED (qs8): %s\n" % json.dumps(p8r)[:4000])
                sys.exit(1)
        extra.update(layer_tables(peaks))
        # the operators either side of ConvBias (include/b200cb_nn.h), L2-cold
        from tools import bench_nn
        extra["adjacent_ops"] = bench_nn.run(peaks["hbm_gbs"])
        extra["config1_resnet18_fp32_nchw"] = config1_table(peaks, cpu=not args.no_cpu_baseline)
        try:
            extra["latency_small_batch_bf16"] = small_batch_latency(device)
        except Exception as e:  # noqa: BLE001 -- an extra; never a reason to fail the bench line
            extra["latency_small_batch_bf16"] = {"error": str(e)[:200]}
        extra["layers_bf16"] = [{k: r[k] for k in ("name", "algo", "count", "ms", "tflops", "gbs", "bound", "frac_of_floor")}
                                for r in rows]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(steps=3, warmup=1, budget_s=25.0)
        cpu = {k: v for k, v in cpu.items() if not k.startswith("_")}

    if rank == 0:
        tf = flops_mb / (ms_per_mb * 1e-3) / 1e12 / 1.0
        tf_all = tf * world
        # DRAM traffic of the 53 launches of one micro-batch from the committed ncu pass (static, profiles/)
        traffic, traffic_src = None, None
        for rnd in ("r02", "r01"):
            tp = os.path.join(ROOT, "profiles", rnd, "dram_traffic_%s.json" % args.dtype)
            if os.path.exists(tp):
                with open(tp) as f:
                    tj = json.load(f)
                traffic, traffic_src = tj["bytes_per_step"], "static, from the committed ncu pass: " + tj["source"]
                break
        line = {
            "metric": "resnet50_convbias_images_per_s", "value": images_s, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": {"bf16": "bf16", "f16": "f16", "qs8": "s8"}[args.dtype],
            "data": "synthetic",
            "config": {"workload": "resnet50_v1.5_53_convbias_layers_%s_%s" % (args.dtype, args.format),
                       "global_batch": args.global_batch, "images_per_gpu": count, "micro_batch": MICRO_BATCH,
                       "microbatches_per_gpu_per_step": mbs, "format": args.format.upper(),
                       "epilogue": "bias+relu, +z on 16 layers",
                       "conv1_input": "IC 3 zero-padded to %d-byte pixels" % args.conv1_pixel_bytes,
                       "sharding": "batch: 8192/G images per GPU, no collective in the timed path",
                       "l2": "inputs larger than L2: every layer owns private tensors, %.1f GB touched per micro-batch"
                             % (bytes_mb / 1e9)},
            "ms_per_microbatch_256": ms_per_mb, "tflops": tf_all, "tflops_per_gpu": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_burst"] * peak_mul, "unit": "TFLOP/s",
                         "frac": tf / (peaks["tf_burst"] * peak_mul), "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": peaks["src"] + (" burst bf16" if peak_mul == 1 else " burst bf16 x2 (int8)"),
                         "frac_vs_sustained": tf / (peaks["tf_sustained"] * peak_mul),
                         "kernel": "igemm_kernel (the 53 launches of one 256-image micro-batch, per GPU)",
                         "flops_per_launch_set": flops_mb, "algorithmic_bytes_per_launch_set": bytes_mb,
                         "hbm_gbs_achieved": bytes_mb / (ms_per_mb * 1e-3) / 1e9,
                         "composite_floor_ms": floor_ms, "frac_of_composite_floor": floor_ms / ms_per_mb},
            "gpu_launches": launches_mb * mbs * args.steps * world,
            "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
            "host": {"us_per_exec": round(host_us_per_exec, 2), "path": "ConvBiasForward.exec without a pre-computed plan: "
                     "algorithm cache lookup + b200cb_launch (kernel-parameter / tensor-map cache)",
                     "algorithm_cache_hits": mb.ConvBiasForward.cache_hits, "algorithm_cache_misses": mb.ConvBiasForward.cache_misses,
                     "cpu_affinity": affinity},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if parity is not None:
            line["parity"] = {"ok": parity["ok"], "parity_checked_layers": parity["layers"],
                              "samples_per_layer": parity["samples_per_layer"], "max_rel_err": parity["max_rel"],
                              "tolerance": {"bf16": 1e-2, "f16": 2e-3, "qs8": 0}[args.dtype],
                              "slice_n2_bitwise_equal_layers": parity["layers"] - len(parity["slice_mismatch"]),
                              "sharding": parity.get("sharding")}
        if extra is not None:
            line["extra"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

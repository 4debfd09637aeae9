#!/usr/bin/env python
"""bench.py -- ResNet-50 ConvBias images/s on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch: the 53 ConvBiasForward layers of ResNet-50
v1.5 (24 unique shapes; shapes = dnn/test/cuda/conv_test_utils.cpp:23-55),
NHWC, batch 256 per GPU, bias + ReLU (+ residual z on the last conv of each bottleneck), enqueued in
network order on one stream through the megdnn-style operator (megengine_b200.ConvBiasForward ->
C-ABI b200cb_launch -> sm_100a kernels).  Synthetic inputs, random-init weights.

  value     images/s with every tensor resident in HBM (each layer owns private buffers: 14 GB are
            touched per step, so every layer starts L2-cold).
  e2e       images/s with the step's input batch copied from pinned host memory and the final
            activations copied back inside the timed region.
  roofline  dense-conv kernel (igemm) vs the measured bf16 tensor peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference
            the reference's own x86 megdnn ConvBias (oracle/_ref, compiled from /root/reference)
            on the host cores, fp32 NCHW, a bounded batch.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU,
batch-sharded weak scaling, no collective in the timed path; NCCL only for barrier / max-reduce).
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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "resnet50_v1.5_53_convbias_layers", "format": "NCHW", "batch_per_step": cb["_batch"],
                       "note": "reference CPU implementation (megdnn x86/fallback ConvBias) on host cores"},
            "cpu_baseline": {k: v for k, v in cb.items() if not k.startswith("_")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def build_layers(kind, batch, device, conv1_pixel_bytes=8, fmt="nhwc"):
    """private tensors + plan for each of the 53 layer instances."""
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
                    "launches": C.load().b200cb_plan_num_launches(ctypes.byref(plan))})
    return handle, out


def run_step(layers):
    for L in layers:
        S, F, B, Z, D = L["args"]
        L["opr"].exec(S, F, B, Z, D, None, L["ws"], plan=L["plan"])


def per_layer_report(layers, peaks, reps=5):
    """CUDA-event time of every layer inside the running network (warm pipeline, L2-cold data)."""
    import torch
    n = len(layers)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
          for _ in range(reps)]
    for r in range(reps):
        for i, L in enumerate(layers):
            S, F, B, Z, D = L["args"]
            ev[r][i][0].record()
            L["opr"].exec(S, F, B, Z, D, None, L["ws"], plan=L["plan"])
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
    for row in rows.values():
        ms = row["ms"] / row["count"]
        tf = row["flops"] / ms / 1e9
        gbs = row["bytes"] / ms / 1e6
        t_floor = max(row["flops"] / (peaks["tf_burst"] * 1e12), row["bytes"] / (peaks["hbm_gbs"] * 1e9)) * 1e3
        out.append({"name": row["name"], "algo": row["algo"], "count": row["count"], "ms": round(ms, 4),
                    "tflops": round(tf, 1), "gbs": round(gbs, 1),
                    "frac_tensor": round(tf / peaks["tf_burst"], 3), "frac_hbm": round(gbs / peaks["hbm_gbs"], 3),
                    "bound": "tensor" if row["flops"] / (peaks["tf_burst"] * 1e12) > row["bytes"] / (peaks["hbm_gbs"] * 1e9) else "hbm",
                    "floor_ms": round(t_floor, 4), "frac_of_floor": round(t_floor / ms, 3)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16", "qs8"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--format", default="nhwc", choices=["nhwc", "nchw4"], help="nchw4: QuantizedS8 only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
    peaks = load_peaks()

    if args.format == "nchw4" and args.dtype != "qs8":
        raise SystemExit("bench.py: --format nchw4 is the QuantizedS8 configuration (use --dtype qs8)")
    handle, layers = build_layers(args.dtype, args.batch, device, args.conv1_pixel_bytes, args.format)
    flops_step = sum(L["flops"] for L in layers)
    bytes_step = sum(L["bytes"] for L in layers)
    launches_step = sum(L["launches"] for L in layers)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        run_step(layers)
    barrier()

    # ---------------- timed region 1: tensors resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        run_step(layers)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    ms_per_step = dev_ms_max / args.steps
    images_s = world * args.batch / (ms_per_step * 1e-3)

    # ---------------- timed region 2: end to end (pinned host input -> 53 layers -> host output)
    # The serving pipeline a user of the operator runs: every step's input batch is copied from pinned
    # host memory and the step's final activations are copied back and read on the host -- both inside
    # the timed region, every step -- but double buffered on copy streams, so that the H2D of step i+1
    # and the D2H of step i-1 overlap the 53 layers of step i.
    import megengine_b200 as mb
    L0, LN = layers[0], layers[-1]
    S0 = L0["args"][0]
    Dl = LN["args"][4].tensor
    h_in = torch.empty(S0.tensor.shape, dtype=S0.tensor.dtype).pin_memory()
    h_in.copy_(S0.tensor)
    d_in = [S0, mb.make_tensor(torch.empty_like(S0.tensor), S0.layout.dtype)]
    d_out = [torch.empty_like(Dl) for _ in range(2)]
    h_out = [torch.empty(Dl.shape, dtype=Dl.dtype).pin_memory() for _ in range(2)]
    comp = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def enqueue(i):
        b = i & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[b])                 # step i-2 no longer reads this input buffer
            d_in[b].tensor.copy_(h_in, non_blocking=True)
            ev_in[b].record(s_in)
        comp.wait_event(ev_in[b])
        comp.wait_event(ev_out[b])                      # step i-2's result has left d_out[b]
        S, F, B, Z, D = L0["args"]
        L0["opr"].exec(d_in[b], F, B, Z, D, None, L0["ws"], plan=L0["plan"])
        ev_free[b].record(comp)
        for L in layers[1:]:
            S, F, B, Z, D = L["args"]
            L["opr"].exec(S, F, B, Z, D, None, L["ws"], plan=L["plan"])
        d_out[b].copy_(Dl, non_blocking=True)
        ev_done[b].record(comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[b])
            h_out[b].copy_(d_out[b], non_blocking=True)
            ev_out[b].record(s_out)

    e2e_steps = max(4, min(args.steps, 20))
    for b in range(2):
        ev_free[b].record(comp)
        ev_out[b].record(comp)
    for i in range(2):                                  # warm the pipeline
        enqueue(i)
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(e2e_steps):
        enqueue(i)
        if i >= 1:
            ev_out[(i - 1) & 1].synchronize()           # the user reads step i-1's result every step
            checksum += float(h_out[(i - 1) & 1].view(-1)[0])
    ev_out[(e2e_steps - 1) & 1].synchronize()
    checksum += float(h_out[(e2e_steps - 1) & 1].view(-1)[0])
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_images_s = world * args.batch * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    h_in_bytes = h_in.numel() * h_in.element_size()
    h_out_bytes = h_out[0].numel() * h_out[0].element_size()

    report = None
    if rank == 0 and args.layer_report:
        report = per_layer_report(layers, peaks)
        os.makedirs(os.path.dirname(os.path.abspath(args.layer_report)), exist_ok=True)
        with open(args.layer_report, "w") as f:
            json.dump({"dtype": args.dtype, "batch": args.batch, "peaks": peaks, "layers": report}, f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(steps=3, warmup=1, budget_s=25.0)
        cpu = {k: v for k, v in cpu.items() if not k.startswith("_")}

    if rank == 0:
        tf = flops_step / (ms_per_step * 1e-3) / 1e12
        peak_mul = 2.0 if args.dtype == "qs8" else 1.0
        floor_ms = sum(max(L["flops"] / (peaks["tf_burst"] * peak_mul * 1e12), L["bytes"] / (peaks["hbm_gbs"] * 1e9))
                       for L in layers) * 1e3
        # DRAM traffic of the step's 53 launches from the committed ncu pass (profiles/), per step
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r01", "dram_traffic_%s.json" % args.dtype)
        if os.path.exists(tp) and args.batch == 256:
            with open(tp) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["bytes_per_step"], tj["source"]
        line = {
            "metric": "resnet50_convbias_images_per_s", "value": images_s, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "f16": "f16", "qs8": "s8"}[args.dtype],
            "data": "synthetic",
            "config": {"workload": "resnet50_v1.5_53_convbias_layers_%s_%s" % (args.dtype, args.format),
                       "batch_per_gpu": args.batch, "global_batch": args.batch * world, "format": args.format.upper(),
                       "epilogue": "bias+relu, +z on 16 layers",
                       "conv1_input": "IC 3 zero-padded to %d-byte pixels" % args.conv1_pixel_bytes, "sharding": "batch, no collective in timed path",
                       "l2": "inputs larger than L2: every layer owns private tensors, %.1f GB touched per step"
                             % (bytes_step / 1e9)},
            "tflops": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_burst"] * peak_mul, "unit": "TFLOP/s",
                         "frac": tf / (peaks["tf_burst"] * peak_mul), "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": peaks["src"] + (" burst bf16" if peak_mul == 1 else " burst bf16 x2 (int8)"),
                         "frac_vs_sustained": tf / (peaks["tf_sustained"] * peak_mul),
                         "kernel": "igemm_kernel (all 53 launches of the step)",
                         "flops_per_step": flops_step, "algorithmic_bytes_per_step": bytes_step,
                         "hbm_gbs_achieved": bytes_step / (ms_per_step * 1e-3) / 1e9,
                         "composite_floor_ms": floor_ms, "frac_of_composite_floor": floor_ms / ms_per_step},
            "e2e": {"value": e2e_images_s, "unit": "images/s", "h2d_bytes_per_step": h_in_bytes,
                    "d2h_bytes_per_step": h_out_bytes, "steps": e2e_steps,
                    "pipeline": "double-buffered: H2D of step i+1 and D2H of step i-1 overlap step i"},
            "gpu_launches": launches_step * args.steps,
            "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

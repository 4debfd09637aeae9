"""Seeded ConvBias test cases shared by the oracle tests, the golden-vector generator and the
GPU parity tests.  Shapes and value ranges follow the reference's own fixtures:

  dnn/test/common/conv_bias.cpp:19-107,306-402   arg generators (kernel / stride / pad sweeps)
  dnn/test/common/conv_bias.cpp:760-902          check_conv_bias RNG ranges (int8 +-3..8, bias +-50)
  dnn/test/cuda/conv_bias_int8.cpp:798-835       NHWC int8: rng +-8, bias +-50, scales 1.2/1.3/19.99
  dnn/test/cuda/conv_test_utils.cpp:23-55        ResNet-50 layer shapes
  dnn/test/cuda/chanwise_convolution.cpp:846-860 MobileNetV2 depthwise shapes
"""
import ctypes

import numpy as np

from megengine_b200 import _capi as C
from . import _oracle as O

KIND_DT = {"f32": C.DT_F32, "f16": C.DT_F16, "bf16": C.DT_BF16, "qs8": C.DT_QS8}


def f32(x):
    return ctypes.c_float(x).value


def conv_out(i, f, s, p, d=1):
    return (i + 2 * p - ((f - 1) * d + 1)) // s + 1


def make_desc(kind, n, ic, ih, iw, oc, fh, fw=None, stride=1, pad=0, dilate=1, group=1, fmt=C.FORMAT_NHWC,
              nonline=C.NL_IDENTITY, has_z=False, bias_mode=C.BIAS_CHANNEL, mode=C.MODE_CROSS_CORRELATION,
              dst_kind=None, scales=(0.1, 0.2, None, 0.3, 0.3), pad_w=None, stride_w=None, compute_mode=None):
    """scales = (src, filter, bias (None -> src*filter), z, dst) for quantized kinds."""
    fw = fh if fw is None else fw
    d = C.Desc()
    d.format, d.mode, d.nonline_mode = fmt, mode, nonline
    d.sparse = C.SPARSE_GROUP if group > 1 else C.SPARSE_DENSE
    d.group = group
    d.bias_mode, d.has_z = bias_mode, 1 if has_z else 0
    dt = KIND_DT[kind]
    d.src_dtype = d.filter_dtype = dt
    if kind == "qs8":
        ddt = {None: C.DT_QS8, "qs8": C.DT_QS8, "qs32": C.DT_QS32, "f32": C.DT_F32}[dst_kind]
        d.dst_dtype = ddt
        d.bias_dtype = C.DT_F32 if ddt == C.DT_F32 else C.DT_QS32   # naive DISPATCH(QuantizedS8, Float32): fp32 bias / z
        d.z_dtype = ddt if has_z else C.DT_NONE
        d.src_scale, d.filter_scale = f32(scales[0]), f32(scales[1])
        d.bias_scale = f32(d.src_scale * d.filter_scale) if scales[2] is None else f32(scales[2])
        d.z_scale = f32(scales[3]) if has_z else 0.0
        d.dst_scale = f32(scales[4]) if ddt == C.DT_QS8 else d.bias_scale
        d.compute_mode = C.COMPUTE_DEFAULT
    else:
        d.dst_dtype = dt
        d.bias_dtype = dt
        d.z_dtype = dt if has_z else C.DT_NONE
        d.compute_mode = C.COMPUTE_DEFAULT if kind == "f32" else C.COMPUTE_FLOAT32
    if compute_mode is not None:
        d.compute_mode = compute_mode
    d.n, d.ic, d.ih, d.iw, d.oc, d.fh, d.fw = n, ic, ih, iw, oc, fh, fw
    d.pad_h, d.pad_w = pad, pad if pad_w is None else pad_w
    d.stride_h, d.stride_w = stride, stride if stride_w is None else stride_w
    d.dilate_h = d.dilate_w = dilate
    d.oh = conv_out(ih, fh, d.stride_h, d.pad_h, dilate)
    d.ow = conv_out(iw, fw, d.stride_w, d.pad_w, dilate)
    return d


def desc_str(d):
    return ("%s %s n%d ic%d %dx%d oc%d f%dx%d s%d,%d p%d,%d d%d g%d nl%d z%d b%d m%d->%s" % (
        {0: "NCHW", 1: "NHWC", 3: "NCHW4"}[d.format], C.DT_NAME[d.src_dtype], d.n, d.ic, d.ih, d.iw, d.oc,
        d.fh, d.fw, d.stride_h, d.stride_w, d.pad_h, d.pad_w, d.dilate_h, d.group, d.nonline_mode, d.has_z,
        d.bias_mode, d.mode, C.DT_NAME[d.dst_dtype]))


def gen_inputs(d, seed, int_range=8, bias_range=50, float_like="resnet"):
    """Raw-storage numpy arrays (src, filter, bias, z) in the tensor's own memory order."""
    rng = np.random.default_rng(seed)
    icpg = d.ic // d.group
    n_src = int(d.n) * d.ic * d.ih * d.iw
    n_flt = int(d.oc) * icpg * d.fh * d.fw
    n_dst = int(d.n) * d.oc * d.oh * d.ow
    n_bias = {C.BIAS_NONE: 0, C.BIAS_CHANNEL: d.oc, C.BIAS_FULL: n_dst}[d.bias_mode]
    if d.src_dtype == C.DT_QS8:
        src = rng.integers(-int_range, int_range + 1, n_src, dtype=np.int8)
        flt = rng.integers(-int_range, int_range + 1, n_flt, dtype=np.int8)
        bias = rng.integers(-bias_range, bias_range + 1, n_bias).astype(np.int32) if n_bias else None
        if n_bias and d.bias_dtype == C.DT_F32:
            bias = rng.uniform(-0.5, 0.5, n_bias).astype(np.float32)
        z = None
        if d.has_z:
            if d.z_dtype == C.DT_F32:
                z = rng.standard_normal(n_dst, dtype=np.float32)
            elif d.z_dtype == C.DT_QS8:
                z = rng.integers(-127, 128, n_dst, dtype=np.int8)
            else:
                z = rng.integers(-5000, 5001, n_dst).astype(np.int32)
        return src, flt, bias, z
    dt = d.src_dtype
    # src,z ~ N(0,1); filter ~ N(0, sqrt(2/(OC*F^2))) (msra fan_out); bias ~ U(-.5,.5)
    src = O.from_f32(rng.standard_normal(n_src, dtype=np.float32), dt)
    std = np.sqrt(2.0 / (max(d.oc // d.group, 1) * d.fh * d.fw))
    flt = O.from_f32(rng.standard_normal(n_flt, dtype=np.float32) * np.float32(std), dt)
    bias = O.from_f32(rng.uniform(-0.5, 0.5, n_bias).astype(np.float32), dt) if n_bias else None
    z = O.from_f32(rng.standard_normal(n_dst, dtype=np.float32), dt) if d.has_z else None
    return src, flt, bias, z


def fp32_reference(d, src, flt, bias, z, threads=8):
    """The fp32 oracle for a bf16/fp16 case (north_star): inputs rounded to T, upcast to fp32, naive
    Float32 ConvBias."""
    d32 = d.copy()
    dt = d.src_dtype
    d32.src_dtype = d32.filter_dtype = d32.dst_dtype = C.DT_F32
    d32.bias_dtype = C.DT_F32
    if d.has_z:
        d32.z_dtype = C.DT_F32
    d32.compute_mode = C.COMPUTE_DEFAULT
    up = lambda a: O.to_f32(a, dt) if a is not None else None
    return O.oracle_convbias(d32, up(src), up(flt), up(bias), up(z), threads)


# ------------------------------------------------------------------------------ shape tables
# ResNet-50 v1.5 unique conv shapes: (name, IC, H, W, OC, F, stride, pad, has_z)
RESNET50_LAYERS = [
    ("conv1", 3, 224, 224, 64, 7, 2, 3, False, 1),
    ("l1.b0.c1", 64, 56, 56, 64, 1, 1, 0, False, 1),
    ("l1.c2", 64, 56, 56, 64, 3, 1, 1, False, 3),
    ("l1.c3", 64, 56, 56, 256, 1, 1, 0, True, 3),
    ("l1.ds", 64, 56, 56, 256, 1, 1, 0, False, 1),
    ("l1.bN.c1", 256, 56, 56, 64, 1, 1, 0, False, 2),
    ("l2.b0.c1", 256, 56, 56, 128, 1, 1, 0, False, 1),
    ("l2.b0.c2", 128, 56, 56, 128, 3, 2, 1, False, 1),
    ("l2.c3", 128, 28, 28, 512, 1, 1, 0, True, 4),
    ("l2.ds", 256, 56, 56, 512, 1, 2, 0, False, 1),
    ("l2.bN.c1", 512, 28, 28, 128, 1, 1, 0, False, 3),
    ("l2.bN.c2", 128, 28, 28, 128, 3, 1, 1, False, 3),
    ("l3.b0.c1", 512, 28, 28, 256, 1, 1, 0, False, 1),
    ("l3.b0.c2", 256, 28, 28, 256, 3, 2, 1, False, 1),
    ("l3.c3", 256, 14, 14, 1024, 1, 1, 0, True, 6),
    ("l3.ds", 512, 28, 28, 1024, 1, 2, 0, False, 1),
    ("l3.bN.c1", 1024, 14, 14, 256, 1, 1, 0, False, 5),
    ("l3.bN.c2", 256, 14, 14, 256, 3, 1, 1, False, 5),
    ("l4.b0.c1", 1024, 14, 14, 512, 1, 1, 0, False, 1),
    ("l4.b0.c2", 512, 14, 14, 512, 3, 2, 1, False, 1),
    ("l4.c3", 512, 7, 7, 2048, 1, 1, 0, True, 3),
    ("l4.ds", 1024, 14, 14, 2048, 1, 2, 0, False, 1),
    ("l4.bN.c1", 2048, 7, 7, 512, 1, 1, 0, False, 2),
    ("l4.bN.c2", 512, 7, 7, 512, 3, 1, 1, False, 2),
]


def resnet50_desc(layer, kind, n, ic_pad=8):
    name, ic, h, w, oc, f, s, p, has_z, _cnt = layer
    if ic < ic_pad:
        ic = ic_pad  # conv1: IC=3 zero-padded as gopt's PaddingChannelPass would
    is_ds = name.endswith(".ds")
    return make_desc(kind, n, ic, h, w, oc, f, stride=s, pad=p, has_z=has_z,
                     nonline=C.NL_IDENTITY if is_ds else C.NL_RELU)


# MobileNetV2 depthwise 3x3 (C, H, W, stride) -- dnn/test/cuda/chanwise_convolution.cpp:846-860
MOBILENETV2_DW = [
    (32, 112, 112, 1), (96, 112, 112, 2), (144, 56, 56, 1), (144, 56, 56, 2), (192, 28, 28, 1),
    (192, 28, 28, 2), (384, 14, 14, 1), (576, 14, 14, 1), (576, 14, 14, 2), (960, 7, 7, 1),
]


def dw_desc(kind, n, c, h, w, stride, f=3, nonline=C.NL_RELU, has_z=False):
    return make_desc(kind, n, c, h, w, c, f, stride=stride, pad=f // 2, group=c, nonline=nonline, has_z=has_z)


# BASELINE.json configs[0]: the ResNet-18 conv stack (fp32 NCHW, 224x224) the reference runs on its CPU
# handles -- conv1 7x7/2, four stages of two BasicBlocks (3x3) with 1x1/2 projections
# (lite/load_and_run/examples/script/resnet50_mge.py:12-53, src/gopt/test/network.cpp:188-224).
# (name, ic, h, w, oc, f, stride, pad, count)
RESNET18_LAYERS = [
    ("conv1", 3, 224, 224, 64, 7, 2, 3, 1),
    ("l1.c3x3", 64, 56, 56, 64, 3, 1, 1, 4),
    ("l2.b0.c1", 64, 56, 56, 128, 3, 2, 1, 1), ("l2.ds", 64, 56, 56, 128, 1, 2, 0, 1), ("l2.c3x3", 128, 28, 28, 128, 3, 1, 1, 3),
    ("l3.b0.c1", 128, 28, 28, 256, 3, 2, 1, 1), ("l3.ds", 128, 28, 28, 256, 1, 2, 0, 1), ("l3.c3x3", 256, 14, 14, 256, 3, 1, 1, 3),
    ("l4.b0.c1", 256, 14, 14, 512, 3, 2, 1, 1), ("l4.ds", 256, 14, 14, 512, 1, 2, 0, 1), ("l4.c3x3", 512, 7, 7, 512, 3, 1, 1, 3),
]


def resnet18_desc(layer, n):
    name, ic, h, w, oc, f, s, p, _cnt = layer
    return make_desc("f32", n, ic, h, w, oc, f, stride=s, pad=p, fmt=C.FORMAT_NCHW,
                     nonline=C.NL_IDENTITY if name.endswith(".ds") else C.NL_RELU)

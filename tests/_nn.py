"""Test-side access to the oracle / reference of the operators either side of ConvBias (TEST INFRASTRUCTURE):
oracle/nn_oracle.c (C restatement, in liboracle.so) and, when built, the reference's own naive operators through
oracle/ref_driver.cpp.  Case generators follow dnn/test/common/pooling.h, dnn/test/cuda/elemwise_multi_type.cpp,
dnn/test/cuda/type_cvt.cpp and dnn/test/common/convolution.cpp (get_args of the backward-data tests)."""
import ctypes

import numpy as np

from megengine_b200 import _capi as C
from . import _cases as K
from . import _oracle as O

_P = ctypes.c_void_p
_U32, _U64, _F = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float
_SIG = {
    "pooling": [ctypes.POINTER(C.PoolDesc), _P, _P],
    "elemwise": [_U32, _U32, _U64, _P, _P, _P],
    "elemwise_q": [_U32, _U64, _P, _F, _P, _F, _P, _F],
    "typecvt": [_U32, _F, _U32, _F, _U64, _P, _P],
    "deconv": [ctypes.POINTER(C.Desc), _P, _P, _P],
}
_typed = set()


def _fn(kind, name):
    lib = O.oracle_lib() if kind == "oracle" else O.ref_lib()
    f = getattr(lib, ("oracle_" if kind == "oracle" else "ref_") + name)
    if (kind, name) not in _typed:
        f.restype = ctypes.c_int
        if name in _SIG:
            f.argtypes = _SIG[name]
        elif name == "convolution":
            f.argtypes = [ctypes.POINTER(C.Desc), _P, _P, _P]
        elif name == "batch_convbias":
            f.argtypes = [ctypes.POINTER(C.Desc), _P, _P, _P, _P, _P]
        _typed.add((kind, name))
    return f


def _call(kind, name, *args):
    rc = _fn(kind, name)(*args)
    if rc != 0:
        msg = O.ref_lib().ref_last_error().decode() if kind == "ref" else ""
        raise RuntimeError("%s %s failed rc=%d %s" % (kind, name, rc, msg))


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


# ------------------------------------------------------------------ data
def rand_storage(dt, n, rng, lo=-3.0, hi=3.0):
    """raw storage of n values of dtype dt (uint16 bit patterns for half types)."""
    if dt == C.DT_QS8:
        return rng.integers(-128, 128, size=n, dtype=np.int64).astype(np.int8)
    if dt == C.DT_QS32:
        return rng.integers(-200000, 200000, size=n, dtype=np.int64).astype(np.int32)
    x = rng.uniform(lo, hi, size=n).astype(np.float32)
    return O.from_f32(x, dt)


def pool_desc(fmt, mode, dt, n, c, ih, iw, window, stride, pad):
    d = C.PoolDesc()
    d.format, d.mode, d.dtype = fmt, mode, dt
    d.n, d.c, d.ih, d.iw = n, c, ih, iw
    wh, ww = window if isinstance(window, tuple) else (window, window)
    sh, sw = stride if isinstance(stride, tuple) else (stride, stride)
    ph, pw = pad if isinstance(pad, tuple) else (pad, pad)
    d.window_h, d.window_w, d.stride_h, d.stride_w, d.pad_h, d.pad_w = wh, ww, sh, sw, ph, pw
    d.oh = (ih + 2 * ph - wh) // sh + 1
    d.ow = (iw + 2 * pw - ww) // sw + 1
    return d


def pooling(kind, d, src):
    dst = np.zeros(d.n * d.c * d.oh * d.ow, dtype=O.NP_STORAGE[d.dtype])
    _call(kind, "pooling", ctypes.byref(d), _ptr(src), _ptr(dst))
    return dst


def elemwise(kind, mode, dt, a, b):
    dst = np.zeros_like(a)
    _call(kind, "elemwise", mode, dt, len(a), _ptr(a), _ptr(b), _ptr(dst))
    return dst


def elemwise_q(kind, mode, a, sa, b, sb, sd):
    dst = np.zeros_like(a)
    _call(kind, "elemwise_q", mode, len(a), _ptr(a), sa, _ptr(b), sb, _ptr(dst), sd)
    return dst


def typecvt(kind, sdt, ss, ddt, sd, src):
    dst = np.zeros(len(src), dtype=O.NP_STORAGE[ddt])
    _call(kind, "typecvt", sdt, ss, ddt, sd, len(src), _ptr(src), _ptr(dst))
    return dst


def deconv_desc(kind, n, ic, ih, iw, oc, f, stride=1, pad=0, dilate=1, mode=C.MODE_CROSS_CORRELATION,
                scales=(0.11, 0.07, 0.9), grad_kind=None):
    """forward-convolution descriptor of a data-gradient problem: src == grad, dst == diff."""
    d = K.make_desc(kind, n, ic, ih, iw, oc, f, stride=stride, pad=pad, dilate=dilate, mode=mode,
                    bias_mode=C.BIAS_NONE)
    d.bias_dtype = C.DT_NONE
    if kind == "qs8":
        d.src_dtype = d.dst_dtype = d.filter_dtype = C.DT_QS8
        d.dst_scale, d.filter_scale, d.src_scale = K.f32(scales[0]), K.f32(scales[1]), K.f32(scales[2])
        if grad_kind == "qs32":
            d.src_dtype = C.DT_QS32
    return d


def deconv(kind, d, flt, diff):
    grad = np.zeros(d.n * d.ih * d.iw * d.ic, dtype=O.NP_STORAGE[d.src_dtype])
    _call(kind, "deconv", ctypes.byref(d), _ptr(flt), _ptr(diff), _ptr(grad))
    return grad


def ref_convolution(d, src, flt):
    dst = np.zeros(d.n * d.oc * d.oh * d.ow, dtype=O.NP_STORAGE[d.dst_dtype])
    _call("ref", "convolution", ctypes.byref(d), _ptr(src), _ptr(flt), _ptr(dst))
    return dst


def ref_batch_convbias(d, src, flt, bias, z):
    dst = np.zeros(d.n * d.oc * d.oh * d.ow, dtype=O.NP_STORAGE[d.dst_dtype])
    _call("ref", "batch_convbias", ctypes.byref(d), _ptr(src), _ptr(flt), _ptr(bias), _ptr(z), _ptr(dst))
    return dst


# ------------------------------------------------------------------ case tables (shared by the CPU and GPU tests)
F, H, B, Q = C.DT_F32, C.DT_F16, C.DT_BF16, C.DT_QS8
NHWC, NCHW, NCHW4 = C.FORMAT_NHWC, C.FORMAT_NCHW, C.FORMAT_NCHW4
MAXP, AVG, AVGX = C.POOL_MAX, C.POOL_AVERAGE, C.POOL_AVERAGE_COUNT_EXCLUDE_PADDING

POOL_CASES = [
    # name, fmt, mode, dtype, n, c, ih, iw, window, stride, pad
    ("resnet_stem_max3s2_bf16", NHWC, MAXP, B, 2, 64, 28, 28, 3, 2, 1),
    ("resnet_stem_max3s2_qs8", NHWC, MAXP, Q, 2, 64, 28, 28, 3, 2, 1),
    ("global_avg7_bf16", NHWC, AVG, B, 3, 128, 7, 7, 7, 7, 0),
    ("global_avg7_f16", NHWC, AVG, H, 3, 128, 7, 7, 7, 7, 0),
    ("global_avg7_f32", NHWC, AVG, F, 3, 96, 7, 7, 7, 7, 0),
    ("global_avg7_qs8", NHWC, AVG, Q, 3, 128, 7, 7, 7, 7, 0),
    ("avg2s2_f16", NHWC, AVG, H, 2, 24, 12, 10, 2, 2, 0),
    ("avg3s1_pad_include_bf16", NHWC, AVG, B, 2, 32, 9, 11, 3, 1, 1),
    ("avg3s2_pad_exclude_bf16", NHWC, AVGX, B, 2, 40, 9, 11, 3, 2, 1),
    ("avg3s2_pad_exclude_f32", NHWC, AVGX, F, 2, 20, 9, 11, 3, 2, 1),
    ("avg3s2_pad_exclude_qs8", NHWC, AVGX, Q, 2, 48, 9, 11, 3, 2, 1),
    ("avg3s1_pad_include_qs8", NHWC, AVG, Q, 2, 16, 8, 8, 3, 1, 1),
    ("max2s2_f32_c3", NHWC, MAXP, F, 2, 3, 10, 10, 2, 2, 0),
    ("max_rect_f16_c10", NHWC, MAXP, H, 1, 10, 13, 9, (3, 2), (2, 1), (1, 0)),
    ("avg17_bf16_count289", NHWC, AVG, B, 1, 8, 17, 17, 17, 17, 0),
    ("max3s2_nchw_f32", NCHW, MAXP, F, 2, 5, 15, 15, 3, 2, 1),
    ("avg2s2_nchw_bf16", NCHW, AVG, B, 2, 6, 8, 8, 2, 2, 0),
    ("max3s2_nchw4_qs8", NCHW4, MAXP, Q, 2, 16, 15, 15, 3, 2, 1),
    ("avg7_nchw4_qs8", NCHW4, AVG, Q, 2, 32, 7, 7, 7, 7, 0),
]

ELEM_CASES = [(m, dt) for m in (C.ELEM_RELU, C.ELEM_ADD, C.ELEM_MAX, C.ELEM_MUL, C.ELEM_FUSE_ADD_RELU, C.ELEM_H_SWISH)
              for dt in (F, H, B)]
ELEM_Q_CASES = [(C.ELEM_ADD, 0.37, 0.71, 0.9), (C.ELEM_FUSE_ADD_RELU, 0.11, 0.23, 0.19), (C.ELEM_RELU, 0.5, 0.0, 0.3),
                (C.ELEM_H_SWISH, 0.05, 0.0, 0.02), (C.ELEM_ADD, 1.9, 23.0, 0.7)]
TYPECVT_CASES = [(F, 0, H, 0), (F, 0, B, 0), (H, 0, F, 0), (B, 0, F, 0), (F, 0, Q, 0.05), (Q, 0.3, F, 0), (Q, 0.3, H, 0),
                 (Q, 0.3, B, 0), (Q, 0.37, Q, 0.71), (C.DT_QS32, 0.0013, Q, 0.9), (C.DT_QS32, 0.002, F, 0),
                 (H, 0, Q, 0.04), (B, 0, Q, 0.04), (Q, 0.5, C.DT_QS32, 0.01), (F, 0, C.DT_QS32, 0.001)]

DECONV_CASES = [
    # name, kind, n, ic, ih, iw, oc, f, stride, pad, dilate, mode
    ("bf16_3x3_s1_p1", "bf16", 2, 16, 9, 11, 24, 3, 1, 1, 1, 0),
    ("bf16_3x3_s2_p1", "bf16", 2, 16, 12, 12, 32, 3, 2, 1, 1, 0),
    ("bf16_4x4_s2_p1_upsample", "bf16", 2, 8, 16, 16, 16, 4, 2, 1, 1, 0),
    ("f16_2x2_s2_p0", "f16", 2, 24, 10, 10, 8, 2, 2, 0, 1, 0),
    ("f16_3x3_s2_odd_extent", "f16", 1, 8, 13, 11, 16, 3, 2, 1, 1, 0),
    ("bf16_1x1_s1", "bf16", 3, 32, 7, 7, 64, 1, 1, 0, 1, 0),
    ("bf16_3x3_dilate2", "bf16", 1, 8, 12, 12, 16, 3, 1, 2, 2, 0),
    ("bf16_3x3_mode_convolution", "bf16", 2, 8, 8, 8, 16, 3, 1, 1, 1, 1),
    ("f32_3x3_s1_p1", "f32", 2, 8, 8, 9, 12, 3, 1, 1, 1, 0),
    ("f32_4x4_s2_p1", "f32", 1, 4, 12, 12, 8, 4, 2, 1, 1, 0),
    ("qs8_3x3_s1_p1", "qs8", 2, 16, 9, 11, 32, 3, 1, 1, 1, 0),
    ("qs8_4x4_s2_p1", "qs8", 2, 16, 12, 12, 16, 4, 2, 1, 1, 0),
    ("qs8_3x3_s2_p1_sat", "qs8", 1, 32, 10, 10, 64, 3, 2, 1, 1, 0),
    # stride > filter (phases that receive nothing are zero), stride 3, flipped filter with stride 2
    ("bf16_1x1_s2_p0", "bf16", 2, 16, 9, 9, 8, 1, 2, 0, 1, 0),
    ("qs8_5x5_s3_p2", "qs8", 1, 16, 14, 14, 16, 5, 3, 2, 1, 0),
    ("f16_3x3_s2_mode_convolution", "f16", 1, 8, 9, 9, 16, 3, 2, 1, 1, 1),
    ("bf16_3x3_s2_dilate2_upsampled", "bf16", 1, 8, 13, 13, 16, 3, 2, 2, 2, 0),
]

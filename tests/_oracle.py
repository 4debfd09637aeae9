"""Test-side access to the oracle (TEST INFRASTRUCTURE -- never imported by megengine_b200).

  liboracle.so      oracle/convbias_oracle.c  -- C restatement of the naive megdnn ConvBias
  libref_driver.so  oracle/ref_driver.cpp     -- the REAL reference operator (naive + x86
                                                 handles), compiled from /root/reference by
                                                 oracle/build_ref.sh; optional at test time
"""
import ctypes
import os
import subprocess

import numpy as np

from megengine_b200 import _capi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_driver.so")

_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        lib = ctypes.CDLL(ORACLE_SO)
        lib.oracle_convbias.restype = ctypes.c_int
        lib.oracle_convbias.argtypes = [ctypes.POINTER(C.Desc)] + [ctypes.c_void_p] * 5 + [ctypes.c_int]
        lib.oracle_convbias_sample.restype = ctypes.c_int
        lib.oracle_convbias_sample.argtypes = (
            [ctypes.POINTER(C.Desc)] + [ctypes.c_void_p] * 5 + [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int])
        lib.oracle_f2h.restype = ctypes.c_uint16
        lib.oracle_f2h.argtypes = [ctypes.c_float]
        lib.oracle_f2bf.restype = ctypes.c_uint16
        lib.oracle_f2bf.argtypes = [ctypes.c_float]
        lib.oracle_h2f.restype = ctypes.c_float
        lib.oracle_h2f.argtypes = [ctypes.c_uint16]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(REF_SO)
        lib.ref_convbias.restype = ctypes.c_int
        lib.ref_convbias.argtypes = (
            [ctypes.POINTER(C.Desc), ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5 +
            [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_int])
        lib.ref_last_error.restype = ctypes.c_char_p
        lib.ref_hardware_threads.restype = ctypes.c_int
        _ref = lib
    return _ref


NP_STORAGE = {C.DT_F32: np.float32, C.DT_F16: np.uint16, C.DT_BF16: np.uint16,
              C.DT_QS8: np.int8, C.DT_QS32: np.int32}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def dst_elems(d):
    return int(d.n) * d.oc * d.oh * d.ow


def oracle_convbias(d, src, flt, bias, z, threads=8):
    """Run the C restatement. Arrays are raw storage (uint16 bit patterns for fp16/bf16)."""
    dst = np.zeros(dst_elems(d), dtype=NP_STORAGE[d.dst_dtype])
    rc = oracle_lib().oracle_convbias(ctypes.byref(d), _ptr(src), _ptr(flt), _ptr(bias), _ptr(z), _ptr(dst), threads)
    if rc != 0:
        raise RuntimeError("oracle_convbias: configuration not restated (rc=%d)" % rc)
    return dst


def oracle_convbias_sample(d, src, flt, bias, z, idx, threads=8):
    """Compute only outputs idx (linear N,OH,OW,OC order); returns the values at idx (NHWC only)."""
    assert d.format == C.FORMAT_NHWC
    dst = np.zeros(dst_elems(d), dtype=NP_STORAGE[d.dst_dtype])
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    rc = oracle_lib().oracle_convbias_sample(
        ctypes.byref(d), _ptr(src), _ptr(flt), _ptr(bias), _ptr(z), _ptr(dst), _ptr(idx), len(idx), threads)
    if rc != 0:
        raise RuntimeError("oracle_convbias_sample failed rc=%d" % rc)
    return dst[idx.astype(np.int64)]


def ref_convbias(d, src, flt, bias, z, handle_kind=2, threads=1, repeat=1):
    """Run the reference's own operator (naive handle_kind=2, x86 handle_kind=0).
    Returns (dst, seconds[], algo_name)."""
    lib = ref_lib()
    dst = np.zeros(dst_elems(d), dtype=NP_STORAGE[d.dst_dtype])
    secs = (ctypes.c_double * max(repeat, 1))()
    name = ctypes.create_string_buffer(256)
    rc = lib.ref_convbias(ctypes.byref(d), handle_kind, threads, _ptr(src), _ptr(flt), _ptr(bias), _ptr(z),
                          _ptr(dst), repeat, secs, name, 256)
    if rc != 0:
        raise RuntimeError("reference ConvBias failed: " + lib.ref_last_error().decode())
    return dst, list(secs), name.value.decode()


# ------------------------------------------------------------------ bit-pattern helpers
def f32_to_bf16_bits(x):
    """round-to-nearest-even, same arithmetic as bfloat16.hpp:352-373 (finite inputs)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (u + (0x7FFF + ((u >> 16) & 1))) >> 16
    return r.astype(np.uint16)


def bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def f32_to_f16_bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)


def f16_bits_to_f32(b):
    return b.view(np.float16).astype(np.float32)


def to_f32(a, dt):
    if dt == C.DT_BF16:
        return bf16_bits_to_f32(a)
    if dt == C.DT_F16:
        return f16_bits_to_f32(a)
    return a.astype(np.float32)


def from_f32(x, dt):
    if dt == C.DT_BF16:
        return f32_to_bf16_bits(x)
    if dt == C.DT_F16:
        return f32_to_f16_bits(x)
    return np.ascontiguousarray(x, dtype=np.float32)


def megdnn_rel_err(x, y):
    """diff = |x-y| / max(|x|,|y|,1)  (dnn/test/common/utils.h:189-193)."""
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    return np.abs(x - y) / np.maximum(np.maximum(np.abs(x), np.abs(y)), 1.0)

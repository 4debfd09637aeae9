"""GPU-side helpers for the parity tests: run one b200cb_desc through the C-ABI on cuda:0."""
import ctypes
import os

import numpy as np
import torch

from megengine_b200 import _capi as C
from . import _oracle as O

_TORCH_STORAGE = {np.dtype(np.uint16): torch.int16, np.dtype(np.int8): torch.int8,
                  np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32}


def to_dev(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    return torch.from_numpy(a).cuda()


TOOLS_LIB = os.path.join(os.path.dirname(C.LIB_PATH), "libb200cb_tools.so")
_tools = None


def tools_lib():
    """the tools build of the same sources (make tools): identical kernels plus the experiment knobs read
    from the environment (B200CB_PAIR, B200CB_DIRECT, ...), used to force kernel variants in parity runs."""
    global _tools
    if _tools is None:
        lib = ctypes.CDLL(TOOLS_LIB)
        for name, res, args in C.SYMBOLS:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _tools = lib
    return _tools


def plan_for(d, algo=None, lib=None):
    lib = lib or C.load()
    plan = C.Plan()
    if algo is None:
        rc = lib.b200cb_query(ctypes.byref(d), ctypes.byref(plan))
    else:
        rc = lib.b200cb_query_algo(ctypes.byref(d), algo, ctypes.byref(plan))
    return rc, plan


def available_algos(d):
    lib = C.load()
    out = []
    for a in range(lib.b200cb_algo_count()):
        rc, _ = plan_for(d, a)
        if rc == 0:
            out.append((a, lib.b200cb_algo_name(a).decode()))
    return out


def launch(plan, tsrc, tflt, tbias, tz, tdst, tws=None, stream=None, lib=None):
    lib = lib or C.load()
    s = stream if stream is not None else torch.cuda.current_stream()
    p = lambda t: t.data_ptr() if t is not None else None
    rc = lib.b200cb_launch(ctypes.byref(plan), p(tsrc), p(tflt), p(tbias), p(tz), p(tdst), p(tws),
                           tws.numel() * tws.element_size() if tws is not None else 0, s.cuda_stream)
    if rc != 0:
        raise C.B200CBError(rc, "b200cb_launch")


def run_gpu(d, src, flt, bias, z, algo=None, force_im2col=False, lib=None):
    """Returns the dst tensor as a numpy array of raw storage (uint16 bits for fp16/bf16)."""
    rc, plan = plan_for(d, algo, lib)
    if rc != 0:
        raise C.B200CBError(rc, "b200cb_query")
    if force_im2col:
        plan.reserved[2] |= 1
    tsrc, tflt, tbias, tz = to_dev(src), to_dev(flt), to_dev(bias), to_dev(z)
    np_dt = O.NP_STORAGE[d.dst_dtype]
    n_dst = O.dst_elems(d)
    tdst = torch.full((n_dst,), 0x55, dtype=_TORCH_STORAGE[np.dtype(np_dt)], device="cuda") \
        if np_dt != np.float32 else torch.full((n_dst,), float("nan"), device="cuda")
    tws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device="cuda") if plan.workspace_bytes else None
    launch(plan, tsrc, tflt, tbias, tz, tdst, tws, lib=lib)
    torch.cuda.synchronize()
    out = tdst.cpu().numpy()
    if np_dt == np.uint16:
        out = out.view(np.uint16)
    return out, plan


def compare(d, got, want, want_f32=None):
    """Returns (ok, info). Integer dtypes: bit-exact. Float: megdnn relative error metric
    |x-y|/max(|x|,|y|,1) (dnn/test/common/utils.h:189-193) against the fp32 oracle when given."""
    info = {}
    if d.dst_dtype in (C.DT_QS8, C.DT_QS32):
        bad = np.nonzero(got != want)[0]
        info["mismatch"] = int(bad.size)
        info["total"] = int(got.size)
        if bad.size:
            info["first_bad"] = [(int(i), int(got[i]), int(want[i])) for i in bad[:8]]
            info["max_abs"] = int(np.abs(got.astype(np.int64) - want.astype(np.int64)).max())
        return bad.size == 0, info
    g = O.to_f32(got, d.dst_dtype)
    tol = {C.DT_BF16: 1e-2, C.DT_F16: 2e-3, C.DT_F32: 1e-3}[d.dst_dtype]
    ref = want_f32 if want_f32 is not None else O.to_f32(want, d.dst_dtype)
    err = O.megdnn_rel_err(g, ref)
    err = np.where(np.isfinite(g), err, np.inf)
    info["max_rel"] = float(err.max())
    info["tol"] = tol
    if want is not None:
        info["bits_differ_vs_oracle"] = int((got != want).sum())
    bad = np.nonzero(err > tol)[0]
    if bad.size:
        info["nbad"] = int(bad.size)
        info["first_bad"] = [(int(i), float(g[i]), float(ref[i])) for i in bad[:8]]
    return bad.size == 0, info

"""CPU tests: the C restatement of the operators either side of ConvBias (oracle/nn_oracle.c) against (a) the
committed outputs of the reference's own naive operators (tests/golden/nn_golden.npz, generator
tests/golden/make_nn_golden.py) and (b) the live reference when oracle/_ref is built.  Bit-exact for every dtype
except the fp32-accumulating deconvolutions in mode CONVOLUTION (accumulation order differs: 1e-6 relative)."""
import os
import zlib

import numpy as np
import pytest

from megengine_b200 import _capi as C
from . import _cases as K
from . import _nn as N
from . import _oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nn_golden.npz")


def _rng(key):
    return np.random.default_rng(zlib.crc32(key.encode()))


def pool_inputs(case):
    name, fmt, mode, dt, n, c, ih, iw, win, st, pad = case
    d = N.pool_desc(fmt, mode, dt, n, c, ih, iw, win, st, pad)
    return d, N.rand_storage(dt, n * c * ih * iw, _rng("pool" + name))


def elem_inputs(mode, dt, n=1000):
    r = _rng("elem%d_%d" % (mode, dt))
    a, b = N.rand_storage(dt, n, r, -8, 8), N.rand_storage(dt, n, r, -8, 8)
    return a, (b if mode in (C.ELEM_ADD, C.ELEM_MAX, C.ELEM_MUL, C.ELEM_FUSE_ADD_RELU) else None)


def elemq_inputs(i, mode, n=1000):
    r = _rng("elemq%d" % i)
    a, b = N.rand_storage(C.DT_QS8, n, r), N.rand_storage(C.DT_QS8, n, r)
    return a, (b if mode in (C.ELEM_ADD, C.ELEM_FUSE_ADD_RELU) else None)


def cvt_inputs(i, sdt, n=1000):
    return N.rand_storage(sdt, n, _rng("cvt%d" % i), -20, 20)


def deconv_inputs(case):
    name, kind, n, ic, ih, iw, oc, f, s, p, dl, mode = case
    sc = (1.9, 1.3, 3.0) if "sat" in name else (0.11, 0.07, 0.9)
    d = N.deconv_desc(kind, n, ic, ih, iw, oc, f, stride=s, pad=p, dilate=dl, mode=mode, scales=sc)
    r = _rng("deconv" + name)
    lo, hi = (-1.0, 1.0)
    diff = N.rand_storage(d.dst_dtype, n * d.oh * d.ow * oc, r, lo, hi)
    flt = N.rand_storage(d.filter_dtype, oc * f * f * ic, r, lo, hi)
    return d, flt, diff


def all_cases():
    """(key, run(kind) -> output array) for every case; kind = 'oracle' or 'ref'."""
    out = []
    for case in N.POOL_CASES:
        d, src = pool_inputs(case)
        out.append(("pool_" + case[0], lambda kind, d=d, src=src: N.pooling(kind, d, src)))
    for mode, dt in N.ELEM_CASES:
        a, b = elem_inputs(mode, dt)
        out.append(("elem_%d_%d" % (mode, dt), lambda kind, m=mode, dt=dt, a=a, b=b: N.elemwise(kind, m, dt, a, b)))
    for i, (mode, sa, sb, sd) in enumerate(N.ELEM_Q_CASES):
        a, b = elemq_inputs(i, mode)
        out.append(("elemq_%d" % i, lambda kind, m=mode, a=a, b=b, sa=sa, sb=sb, sd=sd: N.elemwise_q(kind, m, a, sa, b, sb, sd)))
    for i, (sdt, ss, ddt, sd) in enumerate(N.TYPECVT_CASES):
        src = cvt_inputs(i, sdt)
        out.append(("cvt_%d" % i, lambda kind, a=(sdt, ss, ddt, sd), src=src: N.typecvt(kind, *a, src)))
    for case in N.DECONV_CASES:
        d, flt, diff = deconv_inputs(case)
        out.append(("deconv_" + case[0], lambda kind, d=d, flt=flt, diff=diff: N.deconv(kind, d, flt, diff)))
    return out


def _same(key, got, want):
    if key.startswith("deconv_") and ("convolution" in key or key.startswith("deconv_f32")):
        # fp32 accumulation in a different tap order (mode CONVOLUTION walks the filter backwards)
        dt = C.DT_F32 if got.dtype == np.float32 else C.DT_BF16
        err = O.megdnn_rel_err(O.to_f32(got, dt), O.to_f32(want, dt)).max()
        assert err <= (1e-5 if dt == C.DT_F32 else 8e-3), (key, err)
        return
    bad = int((got.view(np.uint8) != want.view(np.uint8)).sum())
    assert bad == 0, (key, bad, got[:8], want[:8])


CASES = all_cases()


@pytest.mark.parametrize("key,run", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden_outputs_of_the_reference(key, run):
    g = np.load(GOLDEN)
    assert key in g.files, "regenerate tests/golden/nn_golden.npz"
    _same(key, run("oracle"), g[key])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference():
    for key, run in CASES:
        _same(key, run("oracle"), run("ref"))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_convolution_forward_is_convbias_without_epilogue():
    """the reference's ConvolutionForward == its ConvBias with no bias / z / activation: the front-end may delegate
    (what the product's b200cb_convolution_* does); QuantizedS8 x QuantizedS8 -> QuantizedS32."""
    for kind, dst_kind in (("bf16", None), ("f16", None), ("qs8", "qs32")):
        d = K.make_desc(kind, 2, 16, 9, 9, 32, 3, pad=1, bias_mode=C.BIAS_NONE, dst_kind=dst_kind)
        d.bias_dtype = C.DT_NONE
        src, flt, _b, _z = K.gen_inputs(d, 5)
        want = N.ref_convolution(d, src, flt)
        got = O.oracle_convbias(d, src, flt, None, None)
        assert (got.view(np.uint8) != want.view(np.uint8)).sum() == 0, kind


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_batch_convbias_is_convbias_per_image():
    """the reference's BatchConvBiasForward == ConvBias image by image with that image's filter."""
    d = K.make_desc("qs8", 3, 16, 7, 7, 32, 1, fmt=C.FORMAT_NCHW4, nonline=C.NL_RELU, scales=(0.37, 0.71, None, 1.9, 23.0))
    rng = np.random.default_rng(3)
    src = rng.integers(-128, 128, d.n * d.ic * d.ih * d.iw).astype(np.int8)
    flt = rng.integers(-128, 128, d.n * d.oc * d.ic).astype(np.int8)
    bias = rng.integers(-5000, 5000, d.oc).astype(np.int32)
    want = N.ref_batch_convbias(d, src, flt, bias, None)
    one = d.copy()
    one.n = 1
    per_src, per_dst, per_flt = d.ic * d.ih * d.iw, d.oc * d.oh * d.ow, d.oc * d.ic
    got = np.concatenate([O.oracle_convbias(one, src[i * per_src:(i + 1) * per_src].copy(),
                                            flt[i * per_flt:(i + 1) * per_flt].copy(), bias, None) for i in range(d.n)])
    assert (got != want).sum() == 0

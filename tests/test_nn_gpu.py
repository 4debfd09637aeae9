"""GPU parity of the operators either side of ConvBias (include/b200cb_nn.h) against the oracle
(oracle/nn_oracle.c, pinned to the reference by tests/test_nn_oracle.py), through the Python mirrors of the
reference's operator classes (megengine_b200/nn_ops.py) and the C-ABI.  Pooling / elementwise / TypeCvt: bit-exact
for every dtype.  ConvolutionBackwardData: QuantizedS8 bit-exact, bf16 <= 1e-2, fp16 <= 2e-3, fp32 <= 1e-3
(the ConvBias tolerances, dnn/test/common/utils.h:189-193)."""
import ctypes

import numpy as np
import pytest
import torch

import megengine_b200 as M
from megengine_b200 import _capi as C
from . import _cases as K
from . import _gpu as G
from . import _nn as N
from . import _oracle as O
from .test_nn_oracle import cvt_inputs, deconv_inputs, elem_inputs, elemq_inputs, pool_inputs

pytestmark = pytest.mark.gpu

_DT = {C.DT_F32: M.dtype.Float32(), C.DT_F16: M.dtype.Float16(), C.DT_BF16: M.dtype.BFloat16()}


def _dtype(dt, scale=0.5):
    if dt == C.DT_QS8:
        return M.dtype.QuantizedS8(scale)
    if dt == C.DT_QS32:
        return M.dtype.QuantizedS32(scale)
    return _DT[dt]


def _tensor(arr, shape, dt):
    t = G.to_dev(arr).view(dt.torch_dtype).reshape(shape)
    return M.TensorND(t, M.TensorLayout(tuple(shape), dt))


def _out(shape, dt):
    t = torch.empty(shape, dtype=dt.torch_dtype, device="cuda")
    t.view(torch.uint8).fill_(0x55)
    return M.TensorND(t, M.TensorLayout(tuple(shape), dt))


def _raw(t, dt):
    a = t.tensor.reshape(-1).view(torch.int16 if dt.size == 2 else t.tensor.dtype).cpu().numpy()
    return a.view(np.uint16) if dt.size == 2 else a


def _bits_equal(got, want, what):
    bad = int((got.view(np.uint8) != want.view(np.uint8)).sum())
    assert bad == 0, (what, bad, got[:8], want[:8])


@pytest.fixture(scope="module")
def handle():
    return M.Handle("cuda:0")


@pytest.mark.parametrize("case", N.POOL_CASES, ids=[c[0] for c in N.POOL_CASES])
def test_pooling_bit_exact(case, handle):
    d, src = pool_inputs(case)
    want = N.pooling("oracle", d, src)
    opr = M.PoolingForward(handle)
    p = opr.param
    p.mode, p.format = d.mode, d.format
    p.window_h, p.window_w, p.stride_h, p.stride_w, p.pad_h, p.pad_w = (
        d.window_h, d.window_w, d.stride_h, d.stride_w, d.pad_h, d.pad_w)
    dt = _dtype(d.dtype)
    shape = {C.FORMAT_NHWC: (d.n, d.ih, d.iw, d.c), C.FORMAT_NCHW: (d.n, d.c, d.ih, d.iw),
             C.FORMAT_NCHW4: (d.n, d.c // 4, d.ih, d.iw, 4)}[d.format]
    tsrc = _tensor(src, shape, dt)
    lay = opr.deduce_layout(tsrc.layout)
    tdst = _out(lay.shape, dt)
    opr.exec(tsrc, tdst)
    torch.cuda.synchronize()
    _bits_equal(_raw(tdst, dt), want, case[0])


def test_pooling_rejects_what_the_reference_throws_on(handle):
    opr = M.PoolingForward(handle)
    opr.param.format, opr.param.window_h, opr.param.window_w, opr.param.pad_h, opr.param.pad_w = C.FORMAT_NHWC, 2, 2, 2, 2
    dt = M.dtype.Float32()
    src = _out((1, 4, 4, 8), dt)
    dst = _out(opr.deduce_layout(src.layout).shape, dt)
    with pytest.raises(M.MegDNNError):
        opr.exec(src, dst)   # windows that lie entirely in the padding


@pytest.mark.parametrize("mode,dt", N.ELEM_CASES, ids=["m%d_dt%d" % c for c in N.ELEM_CASES])
def test_elemwise_bit_exact(mode, dt, handle):
    a, b = elem_inputs(mode, dt, n=100003)   # odd length: exercises the vector tail
    want = N.elemwise("oracle", mode, dt, a, b)
    opr = M.ElemwiseForward(handle)
    opr.mode = mode
    mdt = _dtype(dt)
    srcs = [_tensor(a, (len(a),), mdt)] + ([_tensor(b, (len(a),), mdt)] if b is not None else [])
    dst = _out((len(a),), mdt)
    opr.exec(srcs, dst)
    torch.cuda.synchronize()
    _bits_equal(_raw(dst, mdt), want, (mode, dt))


@pytest.mark.parametrize("i", range(len(N.ELEM_Q_CASES)))
def test_quantized_elemwise_bit_exact(i, handle):
    mode, sa, sb, sd = N.ELEM_Q_CASES[i]
    a, b = elemq_inputs(i, mode, n=100003)
    want = N.elemwise_q("oracle", mode, a, sa, b, sb, sd)
    opr = M.ElemwiseMultiType(handle)
    opr.mode = {C.ELEM_ADD: 6, C.ELEM_FUSE_ADD_RELU: 7, C.ELEM_RELU: 15, C.ELEM_H_SWISH: 51}[mode]
    srcs = [_tensor(a, (len(a),), M.dtype.QuantizedS8(sa))]
    if b is not None:
        srcs.append(_tensor(b, (len(a),), M.dtype.QuantizedS8(sb)))
    dst = _out((len(a),), M.dtype.QuantizedS8(sd))
    opr.exec(srcs, dst)
    torch.cuda.synchronize()
    _bits_equal(_raw(dst, dst.layout.dtype), want, N.ELEM_Q_CASES[i])


@pytest.mark.parametrize("mode", [C.ELEM_ADD, C.ELEM_FUSE_ADD_RELU])
@pytest.mark.parametrize("scales", [(0.37, 0.71, 0.9), (0.11, 0.23, 0.19), (1.9, 23.0, 0.7), (0.013, 0.0071, 0.0005),
                                    (3.0, 0.02, 19.99), (1.0, 1.0, 1.0)])
def test_quantized_add_all_byte_pairs(mode, scales):
    """The packed ADD / FUSE_ADD_RELU kernel (two outputs per instruction, the sum as an FMA by a run-time 1.0) over
    EVERY pair of input bytes (65536 + a ragged tail) and scale triples that saturate, that round at .5 ties
    (all ones) and that do neither: bit for bit the naive ElemwiseMultiType (oracle/nn_oracle.c)."""
    sa, sb, sd = scales
    v = np.arange(-128, 128, dtype=np.int8)
    a = np.concatenate([np.repeat(v, 256), v[:7]]).astype(np.int8)
    b = np.concatenate([np.tile(v, 256), v[5:12]]).astype(np.int8)
    want = N.elemwise_q("oracle", mode, a, sa, b, sb, sd)
    ta, tb = G.to_dev(a), G.to_dev(b)
    td = torch.zeros(len(a), dtype=torch.int8, device="cuda")
    rc = C.load().b200cb_elemwise_q(mode, len(a), ta.data_ptr(), np.float32(sa), tb.data_ptr(), np.float32(sb),
                                    td.data_ptr(), np.float32(sd), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    got = td.cpu().numpy()
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (scales, bad[:5], a[bad[:5]], b[bad[:5]], got[bad[:5]], want[bad[:5]])


@pytest.mark.parametrize("i", range(len(N.TYPECVT_CASES)))
def test_typecvt_bit_exact(i, handle):
    sdt, ss, ddt, sd = N.TYPECVT_CASES[i]
    src = cvt_inputs(i, sdt, n=100003)
    want = N.typecvt("oracle", sdt, ss, ddt, sd, src)
    s, d = _dtype(sdt, ss), _dtype(ddt, sd)
    tsrc, tdst = _tensor(src, (len(src),), s), _out((len(src),), d)
    M.TypeCvt(handle).exec(tsrc, tdst)
    torch.cuda.synchronize()
    _bits_equal(_raw(tdst, d), want, N.TYPECVT_CASES[i])


@pytest.mark.parametrize("case", N.DECONV_CASES, ids=[c[0] for c in N.DECONV_CASES])
def test_convolution_backward_data(case, handle):
    d, flt, diff = deconv_inputs(case)
    want = N.deconv("oracle", d, flt, diff)
    opr = M.ConvolutionBackwardData(handle)
    p = opr.param
    p.format, p.mode, p.compute_mode = C.FORMAT_NHWC, d.mode, d.compute_mode
    p.pad_h = p.pad_w = d.pad_h
    p.stride_h = p.stride_w = d.stride_h
    p.dilate_h = p.dilate_w = d.dilate_h
    fdt, ddt, gdt = (_dtype(d.filter_dtype, d.filter_scale), _dtype(d.dst_dtype, d.dst_scale),
                     _dtype(d.src_dtype, d.src_scale))
    tflt = _tensor(flt, (d.oc, d.fh, d.fw, d.ic), fdt)
    tdiff = _tensor(diff, (d.n, d.oh, d.ow, d.oc), ddt)
    tgrad = _out((d.n, d.ih, d.iw, d.ic), gdt)
    if d.stride_h == 1 and d.src_dtype != C.DT_QS8:
        assert opr.deduce_layout(tflt.layout, tdiff.layout).shape == (d.n, d.ih, d.iw, d.ic)
    ws = torch.empty(opr.get_workspace_in_bytes(tflt.layout, tdiff.layout, tgrad.layout), dtype=torch.uint8, device="cuda")
    opr.exec(tflt, tdiff, tgrad, ws)
    torch.cuda.synchronize()
    got = _raw(tgrad, gdt)
    if d.src_dtype == C.DT_QS8:
        assert int((got != want).sum()) == 0, case[0]
        if "sat" in case[0]:
            assert (np.abs(want.astype(np.int32)) >= 127).mean() > 0.005   # the clamp is exercised
    else:
        tol = {C.DT_BF16: 1e-2, C.DT_F16: 2e-3, C.DT_F32: 1e-3}[d.src_dtype]
        err = O.megdnn_rel_err(O.to_f32(got, d.src_dtype), O.to_f32(want, d.src_dtype)).max()
        assert err <= tol, (case[0], err)


@pytest.mark.parametrize("kind,dst_kind", [("bf16", None), ("f16", None), ("qs8", "qs32")])
def test_convolution_forward_front_end(kind, dst_kind, handle):
    """ConvolutionForward mirror: dtype / layout deduction of the reference, compute == ConvBias without epilogue."""
    d = K.make_desc(kind, 2, 64, 14, 14, 128, 3, pad=1, fmt=C.FORMAT_NHWC, bias_mode=C.BIAS_NONE, dst_kind=dst_kind)
    d.bias_dtype = C.DT_NONE
    src, flt, _b, _z = K.gen_inputs(d, 11)
    want = O.oracle_convbias(d, src, flt, None, None)
    opr = M.ConvolutionForward(handle)
    opr.param.format, opr.param.pad_h, opr.param.pad_w = C.FORMAT_NHWC, 1, 1
    if kind != "qs8":
        opr.param.compute_mode = C.COMPUTE_FLOAT32
    sdt, fdt = _dtype(d.src_dtype, d.src_scale), _dtype(d.filter_dtype, d.filter_scale)
    tsrc, tflt = _tensor(src, (d.n, d.ih, d.iw, d.ic), sdt), _tensor(flt, (d.oc, d.fh, d.fw, d.ic), fdt)
    lay = opr.deduce_layout(tsrc.layout, tflt.layout)
    assert lay.shape == (d.n, d.oh, d.ow, d.oc)
    if kind == "qs8":
        assert lay.dtype.enumv == C.DT_QS32 and abs(lay.dtype.scale - d.src_scale * d.filter_scale) < 1e-9
    tdst = _out(lay.shape, lay.dtype)
    nbytes = opr.get_workspace_in_bytes(tsrc.layout, tflt.layout, tdst.layout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda") if nbytes else None
    opr.exec(tsrc, tflt, tdst, None, ws)
    torch.cuda.synchronize()
    got = _raw(tdst, lay.dtype)
    ok, info = G.compare(d, got, want, K.fp32_reference(d, src, flt, None, None) if kind != "qs8" else None)
    assert ok, info


def test_batch_convbias_per_image_filters(handle):
    d = K.make_desc("qs8", 3, 16, 7, 7, 32, 1, fmt=C.FORMAT_NCHW4, nonline=C.NL_RELU, scales=(0.37, 0.71, None, 1.9, 23.0))
    rng = np.random.default_rng(3)
    src = rng.integers(-128, 128, d.n * d.ic * d.ih * d.iw).astype(np.int8)
    flt = rng.integers(-128, 128, d.n * d.oc * d.ic).astype(np.int8)
    bias = rng.integers(-5000, 5000, d.oc).astype(np.int32)
    one = d.copy()
    one.n = 1
    per_src, per_flt = d.ic * d.ih * d.iw, d.oc * d.ic
    want = np.concatenate([O.oracle_convbias(one, src[i * per_src:(i + 1) * per_src].copy(),
                                             flt[i * per_flt:(i + 1) * per_flt].copy(), bias, None) for i in range(d.n)])
    opr = M.BatchConvBiasForward(handle)
    p = opr.param
    p.format, p.nonlineMode = C.FORMAT_NCHW4, C.NL_RELU
    q8 = M.dtype.QuantizedS8
    tsrc = _tensor(src, (d.n, d.ic // 4, d.ih, d.iw, 4), q8(d.src_scale))
    tflt = _tensor(flt, (d.n, d.oc, d.ic // 4, 1, 1, 4), q8(d.filter_scale))
    tbias = _tensor(bias, (1, d.oc // 4, 1, 1, 4), M.dtype.QuantizedS32(d.bias_scale))
    tdst = _out((d.n, d.oc // 4, d.oh, d.ow, 4), q8(d.dst_scale))
    z = M.empty_like_dtype(tdst.layout.dtype)
    nbytes = opr.get_workspace_in_bytes(tsrc.layout, tflt.layout, tbias.layout, z.layout, tdst.layout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda") if nbytes else None
    opr.exec(tsrc, tflt, tbias, z, tdst, ws)
    torch.cuda.synchronize()
    assert int((_raw(tdst, tdst.layout.dtype) != want).sum()) == 0


def test_deconvolution_is_the_adjoint_of_convolution_full_size(handle):
    """Size-independent property at a size the oracle cannot finish: <conv(x, w), y> == <x, deconv(w, y)> exactly in
    integers (QuantizedS8 operands, QuantizedS32 results) -- stride 2, so the phase decomposition, its sub-filters,
    temporaries and the interleave pass are all on the path, against the forward kernel that the rest of the suite pins."""
    n, ic, ih, oc, f, s, pd = 64, 64, 56, 128, 3, 2, 1
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    q8, q32 = M.dtype.QuantizedS8, M.dtype.QuantizedS32
    x = torch.randint(-20, 21, (n, ih, ih, ic), dtype=torch.int8, device="cuda", generator=g)
    w = torch.randint(-20, 21, (oc, f, f, ic), dtype=torch.int8, device="cuda", generator=g)
    fwd = M.ConvolutionForward(handle)
    fwd.param.format, fwd.param.pad_h, fwd.param.pad_w, fwd.param.stride_h, fwd.param.stride_w = C.FORMAT_NHWC, pd, pd, s, s
    X = M.TensorND(x, M.TensorLayout(tuple(x.shape), q8(0.5)))
    W = M.TensorND(w, M.TensorLayout(tuple(w.shape), q8(0.25)))
    lay = fwd.deduce_layout(X.layout, W.layout)
    oh = lay.shape[1]
    Y = _out(lay.shape, lay.dtype)
    nb = fwd.get_workspace_in_bytes(X.layout, W.layout, Y.layout)
    fwd.exec(X, W, Y, None, torch.empty(nb, dtype=torch.uint8, device="cuda") if nb else None)
    y = torch.randint(-20, 21, (n, oh, oh, oc), dtype=torch.int8, device="cuda", generator=g)
    bwd = M.ConvolutionBackwardData(handle)
    bwd.param.format, bwd.param.pad_h, bwd.param.pad_w, bwd.param.stride_h, bwd.param.stride_w = C.FORMAT_NHWC, pd, pd, s, s
    Yd = M.TensorND(y, M.TensorLayout(tuple(y.shape), q8(0.5)))
    Gx = _out((n, ih, ih, ic), q32(0.125))
    ws = torch.empty(bwd.get_workspace_in_bytes(W.layout, Yd.layout, Gx.layout), dtype=torch.uint8, device="cuda")
    plan = bwd.exec(W, Yd, Gx, ws)
    assert plan.reserved[0] == 1          # the phase decomposition ran
    torch.cuda.synchronize()
    lhs = int((Y.tensor.to(torch.int64) * y.to(torch.int64)).sum().item())
    rhs = int((Gx.tensor.to(torch.int64) * x.to(torch.int64)).sum().item())
    assert lhs == rhs and lhs != 0, (lhs, rhs)


@pytest.mark.parametrize("dt", ["bf16", "qs8"])
def test_max_pool_full_size_against_an_independent_reference(dt, handle):
    """the ResNet stem pool at N = 256 (too large for the CPU oracle): max is exact in any arithmetic, so torch's
    max_pool2d on the same tensor is an independent bit-exact reference for the packed-word maximum paths."""
    n, c, hw = 256, 64, 112
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    if dt == "bf16":
        x = torch.randn((n, hw, hw, c), device="cuda", generator=g).to(torch.bfloat16)
        mdt = M.dtype.BFloat16()
    else:
        x = torch.randint(-128, 128, (n, hw, hw, c), dtype=torch.int8, device="cuda", generator=g)
        mdt = M.dtype.QuantizedS8(0.1)
    opr = M.PoolingForward(handle)
    p = opr.param
    p.mode, p.format, p.window_h, p.window_w, p.stride_h, p.stride_w, p.pad_h, p.pad_w = C.POOL_MAX, C.FORMAT_NHWC, 3, 3, 2, 2, 1, 1
    X = M.TensorND(x, M.TensorLayout(tuple(x.shape), mdt))
    lay = opr.deduce_layout(X.layout)
    Y = _out(lay.shape, mdt)
    opr.exec(X, Y)
    ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2).float(), 3, 2, 1).permute(0, 2, 3, 1).to(x.dtype)
    torch.cuda.synchronize()
    assert torch.equal(Y.tensor, ref.contiguous())

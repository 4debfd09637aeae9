"""GPU parity tests: the CUDA path (through the C-ABI, include/b200cb.h) against the oracle on the
same seeded inputs.  Bit-exact for QuantizedS8/S32; bf16 <= 1e-2 and fp16 <= 2e-3 max relative
error |x-y|/max(|x|,|y|,1) against the fp32 oracle (BASELINE.json north_star; metric from
dnn/test/common/utils.h:189-193).  Case generators mirror dnn/test/cuda/conv_bias.cpp:83-124,
234-265,650-714,1390-1439 and dnn/test/cuda/conv_bias_int8.cpp:798-835.
"""
import ctypes
import os
import zlib

import numpy as np
import pytest

from megengine_b200 import _capi as C
from . import _cases as K
from . import _oracle as O
from .gpu_diag import stage_cases

pytestmark = pytest.mark.gpu


def _check(d, opt=None, seed=1234, lib=None):
    from . import _gpu as G
    opt = opt or {}
    src, flt, bias, z = K.gen_inputs(d, seed, int_range=opt.get("int_range", 8), bias_range=opt.get("bias_range", 50))
    want = O.oracle_convbias(d, src, flt, bias, z)
    want_f32 = K.fp32_reference(d, src, flt, bias, z) if d.dst_dtype in (C.DT_BF16, C.DT_F16) else None
    got, plan = G.run_gpu(d, src, flt, bias, z, algo=opt.get("algo"), force_im2col=opt.get("im2col", False), lib=lib)
    ok, info = G.compare(d, got, want, want_f32)
    assert ok, (K.desc_str(d), info)
    return plan, info


def _ids(cases):
    return [c[0] for c in cases]


@pytest.mark.parametrize("name,d,opt", stage_cases("gemm_tiled") + stage_cases("gemm_im2col"),
                         ids=_ids(stage_cases("gemm_tiled") + stage_cases("gemm_im2col")))
def test_pointwise(name, d, opt):
    _check(d, opt)


@pytest.mark.parametrize("name,d,opt", stage_cases("conv"), ids=_ids(stage_cases("conv")))
def test_dense_conv(name, d, opt):
    _check(d, opt)


@pytest.mark.parametrize("name,d,opt", stage_cases("int8"), ids=_ids(stage_cases("int8")))
def test_int8_bit_exact(name, d, opt):
    _check(d, opt)


@pytest.mark.parametrize("name,d,opt", stage_cases("dw"), ids=_ids(stage_cases("dw")))
def test_depthwise(name, d, opt):
    _check(d, opt)


@pytest.mark.parametrize("kind", ["bf16", "f16", "qs8"])
@pytest.mark.parametrize("layer", K.RESNET50_LAYERS, ids=[l[0] for l in K.RESNET50_LAYERS])
def test_resnet50_layers_n1(layer, kind):
    """every ResNet-50 ConvBias shape at N=1 against the full oracle; conv1 at a
    reduced spatial size for the oracle's sake (same filter / stride / pad, IC padded 3 -> 8/16)."""
    d = K.resnet50_desc(layer, kind, 1, ic_pad=16 if kind == "qs8" else 8)
    if layer[0] == "conv1":
        d = K.make_desc(kind, 1, d.ic, 64, 64, 64, 7, stride=2, pad=3, nonline=C.NL_RELU)
    _check(d, {"int_range": 127, "bias_range": 5000} if kind == "qs8" else None, seed=zlib.crc32(layer[0].encode()) & 0xffff)


@pytest.mark.parametrize("kind", ["bf16", "qs8"])
def test_every_algo_variant_agrees(kind):
    """fastrun executes every available algo: each tile width gets its own parity run."""
    from . import _gpu as G
    d = K.make_desc(kind, 2, 128, 14, 14, 512, 3, pad=1, nonline=C.NL_RELU, has_z=True)
    algos = G.available_algos(d)
    assert len(algos) >= 3
    for a, _name in algos:
        _check(d, {"algo": a})


@pytest.mark.parametrize("nl", [C.NL_IDENTITY, C.NL_RELU, C.NL_H_SWISH])
@pytest.mark.parametrize("has_z", [False, True])
def test_int8_epilogue_matrix(nl, has_z):
    """IDENTITY/RELU/H_SWISH x {no z, z} with ~1% saturation: 0 mismatching bytes."""
    d = K.make_desc("qs8", 3, 64, 9, 9, 64, 3, pad=1, nonline=nl, has_z=has_z, scales=(0.37, 0.71, None, 1.9, 23.0))
    _check(d, {"int_range": 127, "bias_range": 5000}, seed=77 + nl)


NCHW4_CASES = [
    # name, n, ic, h, w, oc, f, stride, pad, nonline, has_z     (dnn/test/cuda/conv_bias_int8.cpp:26-121)
    ("1x1", 2, 64, 14, 14, 128, 1, 1, 0, C.NL_RELU, False),
    ("3x3_z", 3, 32, 9, 11, 64, 3, 1, 1, C.NL_RELU, True),
    ("5x5_s2", 2, 16, 17, 17, 32, 5, 2, 2, C.NL_IDENTITY, False),
    ("7x7_ic4_stem", 2, 4, 30, 30, 64, 7, 2, 3, C.NL_RELU, False),
    ("7x7_ic8_stem_wide_z", 1, 8, 12, 300, 64, 7, 2, 3, C.NL_RELU, True),
    ("3x3_ic4_stem_oc32_identity", 3, 4, 18, 22, 32, 3, 2, 1, C.NL_IDENTITY, False),
    ("7x7_ic8_s2_oc128_folded", 2, 8, 30, 30, 128, 7, 2, 3, C.NL_H_SWISH, False),
    ("oc_tail_36_z_hswish", 2, 24, 8, 8, 36, 3, 1, 1, C.NL_H_SWISH, True),
    ("identity_z", 1, 128, 7, 7, 256, 1, 1, 0, C.NL_IDENTITY, True),
    ("resnet_l3_c2_pairable", 2, 256, 14, 14, 256, 3, 1, 1, C.NL_RELU, False),
]


@pytest.mark.parametrize("case", NCHW4_CASES, ids=[c[0] for c in NCHW4_CASES])
def test_nchw4_int8_bit_exact(case):
    """QuantizedS8 NCHW4 (BASELINE config 3): src / filter re-tiled by the pre-pass, dst and z accessed
    in place by the epilogue -- 0 mismatching bytes against the oracle, saturation included."""
    _name, n, ic, h, w, oc, f, s, pd, nl, has_z = case
    d = K.make_desc("qs8", n, ic, h, w, oc, f, stride=s, pad=pd, fmt=C.FORMAT_NCHW4, nonline=nl, has_z=has_z,
                    scales=(0.37, 0.71, None, 1.9, 23.0))
    plan, info = _check(d, {"int_range": 127, "bias_range": 5000}, seed=31 + ic)
    assert info["mismatch"] == 0


NCHW_CASES = [
    # name, kind, n, ic, h, w, oc, f, stride, pad, group, nonline, has_z
    ("bf16_1x1", "bf16", 2, 64, 14, 14, 128, 1, 1, 0, 1, C.NL_RELU, False),
    ("bf16_3x3_z", "bf16", 3, 32, 9, 11, 64, 3, 1, 1, 1, C.NL_RELU, True),
    ("f16_5x5_s2_ic20", "f16", 2, 20, 17, 17, 32, 5, 2, 2, 1, C.NL_IDENTITY, False),
    ("bf16_7x7_ic3_stem", "bf16", 2, 3, 30, 30, 64, 7, 2, 3, 1, C.NL_RELU, False),
    ("f16_3x3_hswish_z", "f16", 2, 24, 8, 8, 40, 3, 1, 1, 1, C.NL_H_SWISH, True),
    ("bf16_depthwise3x3", "bf16", 2, 32, 14, 14, 32, 3, 1, 1, 32, C.NL_RELU, False),
    ("f16_depthwise3x3_s2_z", "f16", 2, 48, 15, 15, 48, 3, 2, 1, 48, C.NL_RELU, True),
]


@pytest.mark.parametrize("case", NCHW_CASES, ids=[c[0] for c in NCHW_CASES])
def test_nchw_16bit_through_relayout_adapter(case):
    """NCHW fp16 / bf16 (the reference's default layout): re-tiled to NHWC in the workspace, the NHWC plan
    runs, dst is re-tiled back -- same tolerances against the fp32 oracle as the NHWC path."""
    _name, kind, n, ic, h, w, oc, f, s, pd, grp, nl, has_z = case
    d = K.make_desc(kind, n, ic, h, w, oc, f, stride=s, pad=pd, group=grp, fmt=C.FORMAT_NCHW, nonline=nl, has_z=has_z)
    _check(d, seed=41 + ic)


def test_nchw4_equals_nhwc_full_size():
    """N = 64 ResNet-50 l2 3x3 (pair tiles): the NCHW4 run equals the NHWC run of the same logical
    tensors, element for element (layout independence at a size the oracle does not reach)."""
    import torch
    from . import _gpu as G
    n, ic, hw, oc = 64, 128, 28, 128
    dn = K.make_desc("qs8", n, ic, hw, hw, oc, 3, pad=1, nonline=C.NL_RELU, has_z=True,
                     scales=(0.05, 0.02, None, 0.07, 0.11))
    d4 = K.make_desc("qs8", n, ic, hw, hw, oc, 3, pad=1, fmt=C.FORMAT_NCHW4, nonline=C.NL_RELU, has_z=True,
                     scales=(0.05, 0.02, None, 0.07, 0.11))
    rng = np.random.default_rng(9)
    src = rng.integers(-127, 128, (n, hw, hw, ic), dtype=np.int8)
    flt = rng.integers(-127, 128, (oc, 3, 3, ic), dtype=np.int8)
    z = rng.integers(-127, 128, (n, hw, hw, oc), dtype=np.int8)
    bias = rng.integers(-5000, 5001, oc, dtype=np.int32)
    got_n, _ = G.run_gpu(dn, src.reshape(-1), flt.reshape(-1), bias, z.reshape(-1))
    to4 = lambda a: a.reshape(a.shape[0], a.shape[1], a.shape[2], a.shape[3] // 4, 4).transpose(0, 3, 1, 2, 4)
    got_4, _ = G.run_gpu(d4, to4(src).reshape(-1), to4(flt).reshape(-1), bias, to4(z).reshape(-1))
    back = got_4.reshape(n, oc // 4, hw, hw, 4).transpose(0, 2, 3, 1, 4).reshape(-1)
    assert np.array_equal(back, got_n)
    assert len(np.unique(got_n)) > 50


def test_batch_stride_noncontiguous_src():
    """AlgoBase must accept a non-contiguous batch stride (dnn/src/cuda/conv_bias/algo.h:32-37)."""
    import torch
    from . import _gpu as G
    d = K.make_desc("bf16", 3, 64, 10, 10, 64, 3, pad=1, nonline=C.NL_RELU)
    src, flt, bias, z = K.gen_inputs(d, 5)
    want_f32 = K.fp32_reference(d, src, flt, bias, z)
    img = d.ih * d.iw * d.ic
    bstride = img + 256
    padded = np.zeros(d.n * bstride, dtype=np.uint16)
    for n in range(d.n):
        padded[n * bstride:n * bstride + img] = src[n * img:(n + 1) * img]
    d2 = d.copy()
    d2.src_batch_stride = bstride
    got, _ = G.run_gpu(d2, padded, flt, bias, z)
    ok, info = G.compare(d, got, None, want_f32)
    assert ok, info


def test_full_size_batch_independence_and_sampled_oracle():
    """BASELINE full size (N=256): (i) the N=256 output restricted to the first 2 images equals the
    N=2 run bit for bit (batch independence, SURVEY 8c protocol ii); (ii) a random sample of outputs
    matches the oracle."""
    import torch
    from . import _gpu as G
    layer = [l for l in K.RESNET50_LAYERS if l[0] == "l3.bN.c2"][0]
    d = K.resnet50_desc(layer, "bf16", 256)
    src, flt, bias, z = K.gen_inputs(d, 11)
    got, _ = G.run_gpu(d, src, flt, bias, z)
    d2 = K.resnet50_desc(layer, "bf16", 2)
    n_src2, n_dst2 = 2 * d.ih * d.iw * d.ic, 2 * d.oh * d.ow * d.oc
    got2, _ = G.run_gpu(d2, src[:n_src2], flt, bias, None if z is None else z[:n_dst2])
    assert np.array_equal(got[:n_dst2], got2)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, got.size, 4096).astype(np.uint64)
    d32 = d.copy()
    d32.src_dtype = d32.filter_dtype = d32.dst_dtype = d32.bias_dtype = C.DT_F32
    d32.compute_mode = C.COMPUTE_DEFAULT
    up = lambda a: O.to_f32(a, d.src_dtype) if a is not None else None
    want = O.oracle_convbias_sample(d32, up(src), up(flt), up(bias), up(z), idx)
    err = O.megdnn_rel_err(O.to_f32(got, d.dst_dtype)[idx.astype(np.int64)], want)
    assert err.max() <= 1e-2, err.max()


def test_full_size_int8_linearity_property():
    """Size-independent property at N=256: with IDENTITY, no z, QuantizedS32 dst, conv is linear:
    conv(a) + conv(b) - conv(a+b) == bias  (exact in int32)."""
    from . import _gpu as G
    d = K.make_desc("qs8", 256, 64, 28, 28, 64, 3, pad=1, dst_kind="qs32")
    src_a, flt, bias, _ = K.gen_inputs(d, 21, int_range=40)
    src_b = np.random.default_rng(22).integers(-40, 41, src_a.size, dtype=np.int8)
    ya, _ = G.run_gpu(d, src_a, flt, bias, None)
    yb, _ = G.run_gpu(d, src_b, flt, bias, None)
    yab, _ = G.run_gpu(d, (src_a + src_b).astype(np.int8), flt, bias, None)
    resid = ya.astype(np.int64) + yb - yab
    assert np.array_equal(resid.reshape(-1, d.oc), np.broadcast_to(bias.astype(np.int64), (resid.size // d.oc, d.oc)))


def test_host_mirror_exec_matches_capi():
    """The megdnn-style Python operator (Handle / ConvBiasForward.exec) drives the same kernels."""
    import torch
    import megengine_b200 as mb
    from . import _gpu as G
    d = K.make_desc("bf16", 2, 64, 14, 14, 128, 3, pad=1, nonline=C.NL_RELU, has_z=True)
    src, flt, bias, z = K.gen_inputs(d, 9)
    got_capi, _ = G.run_gpu(d, src, flt, bias, z)
    h = mb.Handle("cuda:0")
    opr = h.create_operator()
    opr.param.format = mb.Param.Format.NHWC
    opr.param.nonlineMode = mb.Param.NonlineMode.RELU
    opr.param.pad_h = opr.param.pad_w = 1
    opr.param.compute_mode = mb.Param.ComputeMode.FLOAT32
    bf = mb.dtype.BFloat16()
    t = lambda a, shape: torch.from_numpy(a.view(np.int16)).cuda().view(torch.bfloat16).reshape(shape)
    tsrc, tflt = t(src, (2, 14, 14, 64)), t(flt, (128, 3, 3, 64))
    tbias, tz = t(bias, (1, 1, 1, 128)), t(z, (2, 14, 14, 128))
    tdst = torch.empty((2, 14, 14, 128), dtype=torch.bfloat16, device="cuda")
    S, F, B, Z, D = (mb.make_tensor(x, bf) for x in (tsrc, tflt, tbias, tz, tdst))
    assert tuple(opr.deduce_layout(S.layout, F.layout, B.layout, Z.layout).shape) == (2, 14, 14, 128)
    assert opr.get_workspace_in_bytes(S.layout, F.layout, B.layout, Z.layout, D.layout) == 0
    opr.exec(S, F, B, Z, D)
    torch.cuda.synchronize()
    got = tdst.view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1)
    assert np.array_equal(got, got_capi)


def test_requant_division_selftest():
    """the epilogue's reciprocal-based division == IEEE division for all 2^32 dividends (8 scales)."""
    import ctypes
    import torch
    lib = C.load()
    bad = ctypes.c_uint64(123)
    rc = lib.b200cb_selftest(0, torch.cuda.current_stream().cuda_stream, ctypes.byref(bad))
    assert rc == 0 and bad.value == 0, (rc, bad.value)


def test_requant_fast_quantize_selftest():
    """the ALU-only quantize of the fast int8 epilogue (unguarded Markstein division, magic-number
    round-half-away, clamp) == roundf(IEEE division) + clamp for all dividends with |quotient| < 2^60."""
    import ctypes
    import torch
    lib = C.load()
    bad = ctypes.c_uint64(123)
    rc = lib.b200cb_selftest(1, torch.cuda.current_stream().cuda_stream, ctypes.byref(bad))
    assert rc == 0 and bad.value == 0, (rc, bad.value)


def test_requant_packed_selftest():
    """the packed fp32x2 requantize (FMUL2 / FFMA2 / FADD2.RZ / .RM, RELU and signed forms, with z) equals
    the scalar fast sequence for all 2^32 accumulator values x 8 scales, both lanes."""
    import ctypes
    import torch
    lib = C.load()
    bad = ctypes.c_uint64(123)
    rc = lib.b200cb_selftest(2, torch.cuda.current_stream().cuda_stream, ctypes.byref(bad))
    assert rc == 0 and bad.value == 0, (rc, bad.value)


@pytest.mark.parametrize("kind", ["bf16", "qs8"])
def test_stem_weight_preprocess_path(kind):
    """exec_preprocess + exec(preprocessed_filter) (nn.h:400-447) == the workspace path, bit for bit."""
    import ctypes
    import torch
    from . import _gpu as G
    lib = C.load()
    d = K.make_desc(kind, 2, 16 if kind == "qs8" else 8, 30, 30, 64, 7, stride=2, pad=3, nonline=C.NL_RELU,
                    scales=(1.2, 1.3, None, 0.7, 19.99))
    src, flt, bias, z = K.gen_inputs(d, 17, int_range=127)
    want, plan = G.run_gpu(d, src, flt, bias, z)
    assert plan.workspace_bytes > 0 and lib.b200cb_plan_num_launches(ctypes.byref(plan)) == 2
    nbytes = lib.b200cb_preprocessed_filter_bytes(ctypes.byref(plan))
    assert nbytes == d.oc * d.fh * 4 * (2 * d.ic) * (1 if kind == "qs8" else 2)
    tflt = G.to_dev(flt)
    packed = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert lib.b200cb_preprocess_filter(ctypes.byref(plan), tflt.data_ptr(), packed.data_ptr(), s) == 0
    assert lib.b200cb_plan_use_preprocessed_filter(ctypes.byref(plan)) == 0
    assert plan.workspace_bytes == 0 and lib.b200cb_plan_num_launches(ctypes.byref(plan)) == 1
    tsrc, tbias = G.to_dev(src), G.to_dev(bias)
    tdst = torch.zeros(O.dst_elems(d), dtype=torch.int8 if kind == "qs8" else torch.int16, device="cuda")
    G.launch(plan, tsrc, packed, tbias, None, tdst)
    torch.cuda.synchronize()
    got = tdst.cpu().numpy()
    assert np.array_equal(got.view(want.dtype), want)


@pytest.mark.parametrize("meta,want", __import__("tests.golden.make_golden", fromlist=["load"]).load(),
                         ids=[m["name"] for m, _ in __import__("tests.golden.make_golden", fromlist=["load"]).load()])
def test_gpu_matches_golden_reference_outputs(meta, want):
    """CUDA path vs the committed outputs of the reference's own naive operator (tests/golden)."""
    from . import _gpu as G
    from .golden import make_golden as GG
    d = GG.dict_to_desc(meta["desc"])
    rc, _ = G.plan_for(d)
    if rc != 0:
        pytest.skip("cell outside the B200 algo family's is_available matrix: " + C.strerror(rc))
    opt = meta["opt"]
    src, flt, bias, z = K.gen_inputs(d, meta["seed"], int_range=opt.get("int_range", 8),
                                     bias_range=opt.get("bias_range", 50))
    got, _ = G.run_gpu(d, src, flt, bias, z)
    if d.dst_dtype in (C.DT_QS8, C.DT_QS32):
        assert np.array_equal(got, want)
    else:
        want_f32 = K.fp32_reference(d, src, flt, bias, z) if d.dst_dtype != C.DT_F32 else None
        ok, info = G.compare(d, got, want, want_f32)
        assert ok, info


# ------------------------------------------------------------------ forced kernel variants (tools build)
_VARIANT_CASES = (stage_cases("gemm_tiled") + stage_cases("conv") + stage_cases("int8"))


@pytest.fixture
def knob(monkeypatch):
    def set_(name, value):
        monkeypatch.setenv(name, str(value))
    return set_


@pytest.mark.parametrize("name,d,opt", _VARIANT_CASES, ids=_ids(_VARIANT_CASES))
def test_cta_pair_forced_on_every_case(name, d, opt, knob):
    """cta_group::2 tiles forced wherever the kernel can pair (>= 4 M tiles), whatever the heuristic says:
    same parity bar.  Runs the tools build of the same sources (B200CB_PAIR is its knob)."""
    from . import _gpu as G
    if not os.path.exists(G.TOOLS_LIB):
        pytest.skip("tools build (make -C megengine_b200/csrc tools) not present")
    knob("B200CB_PAIR", 1)
    _check(d, opt, lib=G.tools_lib())


@pytest.mark.parametrize("name,d,opt", _VARIANT_CASES, ids=_ids(_VARIANT_CASES))
def test_direct_store_epilogue_forced_on_every_case(name, d, opt, knob):
    """the st.global epilogue (instead of the TMA store) on every case: same parity bar."""
    from . import _gpu as G
    if not os.path.exists(G.TOOLS_LIB):
        pytest.skip("tools build not present")
    knob("B200CB_DIRECT", 1)
    _check(d, opt, lib=G.tools_lib())


@pytest.mark.parametrize("kind", ["bf16", "f16", "qs8"])
@pytest.mark.parametrize("layer", [l for l in K.RESNET50_LAYERS if l[0] in ("l2.bN.c2", "l3.c3", "l4.bN.c1")],
                         ids=["l2.bN.c2", "l3.c3", "l4.bN.c1"])
def test_pair_tiles_at_n8_against_full_oracle(layer, kind, knob):
    """N = 8 is the smallest batch at which the ResNet-50 layers have enough M tiles to pair: forced pairs
    against the full oracle (phantom tiles: 8 * 7 * 7 = 392 pixels = 3.06 M tiles on 2 CTAs)."""
    from . import _gpu as G
    if not os.path.exists(G.TOOLS_LIB):
        pytest.skip("tools build not present")
    knob("B200CB_PAIR", 1)
    d = K.resnet50_desc(layer, kind, 8, ic_pad=16 if kind == "qs8" else 8)
    plan, _info = _check(d, {"int_range": 127, "bias_range": 5000} if kind == "qs8" else None, lib=G.tools_lib())
    assert plan.reserved[7] == 1


@pytest.mark.parametrize("name,d,opt", stage_cases("gemm_tiled") + stage_cases("int8"),
                         ids=_ids(stage_cases("gemm_tiled") + stage_cases("int8")))
def test_resident_filter_tile_forced_on_every_pointwise_case(name, d, opt, knob):
    """IgemmParams::b_res (the CTA's filter tile loaded once, activation-only pipeline stages) forced on every
    1x1 stride-1 case of every dtype through the tools build; cases of another shape ignore the knob."""
    from . import _gpu as G
    if not os.path.exists(G.TOOLS_LIB):
        pytest.skip("tools build not present")
    knob("B200CB_BRES", 1)
    knob("B200CB_PAIR", 0)
    _check(d, opt, lib=G.tools_lib())


@pytest.mark.parametrize("kind,spec,n", [("qs8", "l1.c3", 256), ("qs8", "l4.c3", 256), ("qs8", "l2.bN.c1", 64),
                                         ("bf16", "l3.c3", 64), ("f16", "l2.c3", 32),
                                         ("bf16", (16, 112, 96), 64), ("f16", (24, 56, 144), 128),
                                         ("bf16", (32, 28, 192), 64), ("qs8", (16, 112, 96), 64)],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else str(v))
def test_resident_filter_tile_full_size_equals_streamed(kind, spec, n, knob):
    """Many tiles per CTA (grid cut to a multiple of n_tiles, every CTA keeps ONE N tile of the filter): the
    resident-filter kernel must reproduce the streamed-filter kernel bit for bit (same MMA order), and the
    oracle on 4096 samples.  QuantizedS8 and low-K 16-bit layers (MobileNetV2 pointwise: ic, hw, oc) are resident
    by the product heuristic; the other 16-bit layers force it through the tools build."""
    import torch
    from . import _gpu as G
    if isinstance(spec, tuple):
        ic, hw, oc = spec
        d = K.make_desc(kind, n, ic, hw, hw, oc, 1, nonline=C.NL_RELU)
    else:
        layer = [l for l in K.RESNET50_LAYERS if l[0] == spec][0]
        d = K.resnet50_desc(layer, kind, n)
    by_default = kind == "qs8" or d.ic * 2 <= 64
    src, flt, bias, z = K.gen_inputs(d, 31, int_range=127, bias_range=5000)
    outs = {}
    for mode in ("resident", "streamed"):
        if by_default and mode == "resident":
            lib = None                                  # product library, product heuristic
        else:
            if not os.path.exists(G.TOOLS_LIB):
                pytest.skip("tools build not present")
            knob("B200CB_BRES", 1 if mode == "resident" else 0)
            knob("B200CB_PAIR", 0)
            lib = G.tools_lib()
        algo = None
        if not by_default:      # a 128-wide tile: the 256-wide filter tile of a K = 512-byte layer does not fit
            names = [(lib or C.load()).b200cb_algo_name(a).decode() for a in range((lib or C.load()).b200cb_algo_count())]
            algo = [a for a, nm in enumerate(names) if nm.endswith("128x128")][0]
        rc, plan = G.plan_for(d, algo, lib=lib)
        assert rc == 0
        assert bool(plan.reserved[2] & 8) == (mode == "resident"), (mode, plan.block_n)
        t = [G.to_dev(a) for a in (src, flt, bias, z)]
        out = torch.zeros(O.dst_elems(d), dtype=torch.int8 if kind == "qs8" else torch.int16, device="cuda")
        G.launch(plan, t[0], t[1], t[2], t[3], out, lib=lib)
        torch.cuda.synchronize()
        outs[mode] = out.cpu().numpy()
    assert np.array_equal(outs["resident"], outs["streamed"])
    got = outs["resident"]
    idx = np.random.default_rng(7).integers(0, got.size, 4096).astype(np.uint64)
    if kind == "qs8":
        want = O.oracle_convbias_sample(d, src, flt, bias, z, idx)
        assert np.array_equal(got[idx.astype(np.int64)], want)
    else:
        d32 = d.copy()
        d32.src_dtype = d32.filter_dtype = d32.dst_dtype = d32.bias_dtype = d32.z_dtype = C.DT_F32
        d32.compute_mode = C.COMPUTE_DEFAULT
        up = lambda a: None if a is None else O.to_f32(a, d.src_dtype)
        want = O.oracle_convbias_sample(d32, up(src), up(flt), up(bias), up(z), idx)
        err = O.megdnn_rel_err(O.to_f32(got.view(np.uint16), d.dst_dtype)[idx.astype(np.int64)], want)
        assert err.max() <= (1e-2 if kind == "bf16" else 2e-3), err.max()


PW_FOLD_CASES = [
    # name, kind, n, ic, h, w, oc, has_z, nonline, bias_mode
    ("bf16_32to16_112", "bf16", 16, 32, 112, 112, 16, False, C.NL_RELU, C.BIAS_CHANNEL),
    ("f16_16to24_odd_h_z", "f16", 3, 16, 7, 10, 24, True, C.NL_IDENTITY, C.BIAS_CHANNEL),
    ("bf16_24to32_nobias", "bf16", 2, 24, 9, 6, 32, False, C.NL_H_SWISH, C.BIAS_NONE),
    ("qs8_32to16_z_relu", "qs8", 5, 32, 14, 14, 16, True, C.NL_RELU, C.BIAS_CHANNEL),
    ("qs8_16to32_w2", "qs8", 2, 16, 5, 2, 32, False, C.NL_IDENTITY, C.BIAS_CHANNEL),
    ("qs8_64to16_many_tiles", "qs8", 64, 64, 28, 28, 16, False, C.NL_RELU, C.BIAS_CHANNEL),
]


@pytest.mark.parametrize("case", PW_FOLD_CASES, ids=[c[0] for c in PW_FOLD_CASES])
def test_pointwise_pixel_pair_fold(case, knob):
    """Narrow 1x1 layers run as the pixel-pair folded layer (2 IC -> 2 OC over W / 2, block-diagonal filter built in
    the workspace): the oracle's bar, and bit for bit the unfolded kernel (tools build, B200CB_PWFOLD=0) -- the
    extra products are exact zeros."""
    import torch
    from . import _gpu as G
    _name, kind, n, ic, h, w, oc, has_z, nl, bias_mode = case
    d = K.make_desc(kind, n, ic, h, w, oc, 1, nonline=nl, has_z=has_z, bias_mode=bias_mode)
    opt = {"int_range": 127, "bias_range": 3000} if kind == "qs8" else None
    plan, _info = _check(d, opt)
    assert plan.workspace_bytes >= 4 * oc * ic and C.load().b200cb_plan_num_launches(ctypes.byref(plan)) == 2
    if not os.path.exists(G.TOOLS_LIB):
        pytest.skip("tools build not present")
    src, flt, bias, z = K.gen_inputs(d, 11, **(opt or {}))
    got, _ = G.run_gpu(d, src, flt, bias, z)
    knob("B200CB_PWFOLD", 0)
    ref, plan0 = G.run_gpu(d, src, flt, bias, z, lib=G.tools_lib())
    assert plan0.workspace_bytes == 0
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind", ["bf16", "f16", "qs8"])
def test_full_size_residual_layer_repeated_and_sampled(kind):
    """ResNet-50 l1.c3 at N = 256 (802816 x 256 outputs, residual z, one k-block per tile): the layer whose
    z ring hung / faulted in round 2 when a ring slot was shared by two epilogue groups.  40 back-to-back
    launches must complete, agree with each other bit for bit, and match the oracle on 4096 samples."""
    import torch
    from . import _gpu as G
    layer = [l for l in K.RESNET50_LAYERS if l[0] == "l1.c3"][0]
    d = K.resnet50_desc(layer, kind, 256)
    src, flt, bias, z = K.gen_inputs(d, 23, int_range=127, bias_range=5000)
    rc, plan = G.plan_for(d)
    assert rc == 0
    tsrc, tflt, tbias, tz = G.to_dev(src), G.to_dev(flt), G.to_dev(bias), G.to_dev(z)
    outs = [torch.zeros(O.dst_elems(d), dtype=torch.int8 if kind == "qs8" else torch.int16, device="cuda") for _ in range(2)]
    for i in range(40):
        G.launch(plan, tsrc, tflt, tbias, tz, outs[i & 1])
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    got = outs[0].cpu().numpy()
    idx = np.random.default_rng(5).integers(0, got.size, 4096).astype(np.uint64)
    if kind == "qs8":
        want = O.oracle_convbias_sample(d, src, flt, bias, z, idx)
        assert np.array_equal(got[idx.astype(np.int64)], want)
    else:
        d32 = d.copy()
        d32.src_dtype = d32.filter_dtype = d32.dst_dtype = d32.bias_dtype = d32.z_dtype = C.DT_F32
        d32.compute_mode = C.COMPUTE_DEFAULT
        up = lambda a: O.to_f32(a, d.src_dtype)
        want = O.oracle_convbias_sample(d32, up(src), up(flt), up(bias), up(z), idx)
        err = O.megdnn_rel_err(O.to_f32(got.view(np.uint16), d.dst_dtype)[idx.astype(np.int64)], want)
        assert err.max() <= (1e-2 if kind == "bf16" else 2e-3), err.max()


# ------------------------------------------------------------------ host-buffer entry point
@pytest.mark.parametrize("kind,has_z", [("bf16", True), ("qs8", True), ("f16", False)])
def test_exec_host_matches_device_path(kind, has_z):
    """b200cb_exec_host (include/b200cb.h): host src / filter / bias / z in, host dst out, device scratch
    owned by the caller -- the same bytes as b200cb_launch on device tensors."""
    import ctypes
    import torch
    from . import _gpu as G
    lib = C.load()
    d = K.make_desc(kind, 3, 64, 14, 14, 128, 3, pad=1, nonline=C.NL_RELU, has_z=has_z, scales=(0.37, 0.71, None, 1.9, 23.0))
    src, flt, bias, z = K.gen_inputs(d, 3, int_range=127, bias_range=5000)
    want, plan = G.run_gpu(d, src, flt, bias, z)
    need = lib.b200cb_host_scratch_bytes(ctypes.byref(plan))
    assert need >= src.nbytes + flt.nbytes + want.nbytes
    scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    got = np.zeros_like(want)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    s = torch.cuda.current_stream().cuda_stream
    assert lib.b200cb_exec_host(ctypes.byref(plan), p(src), p(flt), p(bias), p(z), p(got), scratch.data_ptr(), need, s) == 0
    assert np.array_equal(got, want)
    # too little scratch / missing buffers are reported, not executed
    assert lib.b200cb_exec_host(ctypes.byref(plan), p(src), p(flt), p(bias), p(z), p(got), scratch.data_ptr(), need - 1, s) == -3
    assert lib.b200cb_exec_host(ctypes.byref(plan), None, p(flt), p(bias), p(z), p(got), scratch.data_ptr(), need, s) == -5


# ------------------------------------------------------------------ first-class relayout operators (8f rank 1)
def _np_relayout(mode, x, n, c, h, w, c_out):
    """numpy statement of megdnn RelayoutFormat for the modes of include/b200cb.h (zero channel padding)."""
    def padc(a, axis, to):
        pad = [(0, 0)] * a.ndim
        pad[axis] = (0, to - a.shape[axis])
        return np.pad(a, pad)
    if mode == C.RELAYOUT_NCHW_NHWC:
        return padc(x.reshape(n, c, h, w).transpose(0, 2, 3, 1), 3, c_out)
    if mode == C.RELAYOUT_NHWC_NCHW:
        return x.reshape(n, h, w, c).transpose(0, 3, 1, 2)
    if mode == C.RELAYOUT_NCHW_NCHW4:
        return padc(x.reshape(n, c, h, w), 1, c_out).reshape(n, c_out // 4, 4, h, w).transpose(0, 1, 3, 4, 2)
    if mode == C.RELAYOUT_NCHW4_NCHW:
        return x.reshape(n, c // 4, h, w, 4).transpose(0, 1, 4, 2, 3).reshape(n, c, h, w)[:, :c_out]
    if mode == C.RELAYOUT_NCHW4_NHWC:
        return padc(x.reshape(n, c // 4, h, w, 4).transpose(0, 2, 3, 1, 4).reshape(n, h, w, c), 3, c_out)
    if mode == C.RELAYOUT_NHWC_NCHW4:
        return padc(x.reshape(n, h, w, c), 3, c_out).reshape(n, h, w, c_out // 4, 4).transpose(0, 3, 1, 2, 4)
    return padc(x.reshape(n, h, w, c), 3, c_out)


RELAYOUT_CASES = [
    # mode, elem bytes, n, c, h, w, c_out
    (C.RELAYOUT_NCHW_NHWC, 2, 3, 20, 9, 11, 24), (C.RELAYOUT_NCHW_NHWC, 4, 2, 3, 17, 5, 3), (C.RELAYOUT_NCHW_NHWC, 1, 2, 64, 7, 7, 64),
    (C.RELAYOUT_NHWC_NCHW, 2, 3, 40, 9, 11, 40), (C.RELAYOUT_NHWC_NCHW, 4, 2, 7, 5, 33, 7), (C.RELAYOUT_NHWC_NCHW, 1, 2, 96, 14, 14, 96),
    (C.RELAYOUT_NCHW_NCHW4, 1, 2, 3, 30, 30, 4), (C.RELAYOUT_NCHW_NCHW4, 1, 3, 22, 9, 11, 24), (C.RELAYOUT_NCHW_NCHW4, 2, 2, 8, 6, 7, 8),
    (C.RELAYOUT_NCHW4_NCHW, 1, 2, 24, 9, 11, 22), (C.RELAYOUT_NCHW4_NCHW, 4, 2, 8, 5, 5, 8),
    (C.RELAYOUT_NCHW4_NHWC, 1, 3, 32, 9, 11, 32), (C.RELAYOUT_NCHW4_NHWC, 1, 2, 4, 30, 30, 16), (C.RELAYOUT_NCHW4_NHWC, 2, 2, 36, 7, 7, 40),
    (C.RELAYOUT_NHWC_NCHW4, 1, 2, 3, 30, 30, 4), (C.RELAYOUT_NHWC_NCHW4, 1, 3, 64, 9, 11, 64), (C.RELAYOUT_NHWC_NCHW4, 2, 2, 10, 5, 6, 12),
    (C.RELAYOUT_NHWC_PAD_C, 2, 2, 3, 224, 224, 4), (C.RELAYOUT_NHWC_PAD_C, 1, 3, 3, 31, 17, 8), (C.RELAYOUT_NHWC_PAD_C, 4, 1, 5, 9, 9, 8),
    (C.RELAYOUT_NHWC_PAD_C, 2, 1, 24, 13, 3, 32),
]


@pytest.mark.parametrize("case", RELAYOUT_CASES, ids=["m%d_e%d_n%d_c%d_%dx%d_co%d" % c for c in RELAYOUT_CASES])
def test_relayout_modes_match_numpy(case):
    """b200cb_relayout: NCHW <-> NHWC <-> NCHW4 and channel padding, byte-exact against the numpy statement
    of megdnn's RelayoutFormat (dnn/src/naive/relayout_format/opr_impl.cpp semantics: pure permutation with
    zero-filled channel padding)."""
    import torch
    mode, es, n, c, h, w, c_out = case
    lib = C.load()
    npdt = {1: np.int8, 2: np.int16, 4: np.int32}[es]
    x = np.random.default_rng(mode * 7 + es).integers(-100, 100, n * c * h * w).astype(npdt)
    want = np.ascontiguousarray(_np_relayout(mode, x, n, c, h, w, c_out))
    tx = torch.from_numpy(x).cuda()
    ty = torch.full((want.size,), 0x55, dtype=tx.dtype, device="cuda")
    rc = lib.b200cb_relayout(mode, es, n, c, h, w, c_out, tx.data_ptr(), ty.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, C.strerror(rc)
    torch.cuda.synchronize()
    assert np.array_equal(ty.cpu().numpy(), want.reshape(-1))


# ------------------------------------------------------------------ stem row ring (8-byte pixels)
STEM_RING_CASES = [
    # name, kind, n, ic, h, w, oc, f, pad          (stride 2; bf16 / fp16 IC = 4, int8 IC = 8: 8-byte pixels)
    ("bf16_7x7_224_n4", "bf16", 4, 4, 224, 224, 64, 7, 3),    # 448 output rows: 3 per CTA, image changes mid-range
    ("qs8_7x7_224_n2", "qs8", 2, 8, 224, 224, 64, 7, 3),
    ("f16_7x7_odd_rows", "f16", 3, 4, 74, 62, 64, 7, 3),
    ("bf16_3x3_p1", "bf16", 5, 4, 66, 40, 32, 3, 1),          # 2 row pairs per window
    ("qs8_5x5_p2", "qs8", 3, 8, 50, 100, 48, 5, 2),           # even padding: 3 row pairs, last row unused
    ("bf16_7x7_one_row_per_cta", "bf16", 1, 4, 30, 30, 64, 7, 3),
    ("bf16_7x7_ow125", "bf16", 2, 4, 20, 250, 64, 7, 3),      # widest row that still fits one box
    ("bf16_7x7_ow129_classic", "bf16", 2, 4, 20, 258, 64, 7, 3),  # one pixel too wide: classic stem path
]


@pytest.mark.parametrize("case", STEM_RING_CASES, ids=[c[0] for c in STEM_RING_CASES])
def test_stem_row_ring(case):
    """network stems with 8-byte pixels: whole input rows staged into the row ring, contiguous output-row
    ranges per CTA (first / last rows of an image, image changes inside a range, padding rows)."""
    _name, kind, n, ic, h, w, oc, f, pd = case
    d = K.make_desc(kind, n, ic, h, w, oc, f, stride=2, pad=pd, nonline=C.NL_RELU, scales=(0.37, 0.71, None, 1.9, 23.0))
    plan, _info = _check(d, {"int_range": 127, "bias_range": 5000} if kind == "qs8" else None, seed=61 + h)
    assert plan.a_mode == 2
    assert ((plan.reserved[4] >> 1) & 1) == (0 if ("classic" in _name or kind == "qs8") else 1)


# ------------------------------------------------------------------ a13 cells closed in round 2
GROUP_CASES = [
    # name, kind, n, ic, h, w, oc, f, stride, pad, group, nonline, has_z   (dnn/test/cuda/conv_bias.cpp:234-265 group args)
    ("bf16_g2_3x3_z", "bf16", 2, 64, 14, 14, 128, 3, 1, 1, 2, C.NL_RELU, True),
    ("f16_g4_1x1", "f16", 3, 64, 9, 9, 64, 1, 1, 0, 4, C.NL_IDENTITY, False),
    ("bf16_g32_resnext", "bf16", 2, 256, 14, 14, 256, 3, 1, 1, 32, C.NL_RELU, False),
    ("qs8_g2_3x3_s2_z", "qs8", 2, 64, 15, 15, 64, 3, 2, 1, 2, C.NL_RELU, True),
    ("qs8_g4_5x5_hswish", "qs8", 1, 128, 12, 12, 64, 5, 1, 2, 4, C.NL_H_SWISH, False),
]


@pytest.mark.parametrize("case", GROUP_CASES, ids=[c[0] for c in GROUP_CASES])
def test_general_group_conv(case):
    """param::Sparse::GROUP with icpg, ocpg > 1: one dense launch per group on channel-sliced views of src / z /
    dst (replaces dnn/src/cuda/conv_bias/group_conv.cpp)."""
    _n, kind, n, ic, h, w, oc, f, s, pd, grp, nl, has_z = case
    d = K.make_desc(kind, n, ic, h, w, oc, f, stride=s, pad=pd, group=grp, nonline=nl, has_z=has_z,
                    scales=(0.37, 0.71, None, 1.9, 23.0))
    plan, _info = _check(d, {"int_range": 127, "bias_range": 5000} if kind == "qs8" else None, seed=71 + grp)
    assert C.load().b200cb_algo_name(plan.algo).decode().startswith("B200_IGEMM")


@pytest.mark.parametrize("kind,f,nl", [("bf16", 3, C.NL_RELU), ("f16", 1, C.NL_IDENTITY), ("bf16", 1, C.NL_H_SWISH)])
def test_full_shape_bias(kind, f, nl):
    """BiasMode::BIAS (dst-shaped bias, dnn/src/common/conv_bias.cpp:38-49) through the z ring."""
    d = K.make_desc(kind, 3, 64, 10, 12, 128, f, pad=f // 2, nonline=nl, bias_mode=C.BIAS_FULL)
    _check(d, seed=81 + f)


QS8_F32_CASES = [
    ("3x3_z_relu", 2, 64, 14, 14, 128, 3, 1, 1, C.NL_RELU, True),
    ("1x1_identity", 3, 32, 9, 9, 64, 1, 1, 0, C.NL_IDENTITY, False),
    ("5x5_s2_hswish_z", 2, 16, 17, 17, 36, 5, 2, 2, C.NL_H_SWISH, True),
    ("7x7_s2_ic16", 1, 16, 30, 30, 64, 7, 2, 3, C.NL_RELU, False),
]


@pytest.mark.parametrize("case", QS8_F32_CASES, ids=[c[0] for c in QS8_F32_CASES])
def test_qs8_to_float32(case):
    """QuantizedS8 x QuantizedS8 -> Float32 (naive DISPATCH(QuantizedS8, Float32): fp32 bias / z).  The oracle
    accumulates dequantised products in fp32, the GPU scales the exact int32 sum once: Float32 tolerance 1e-3."""
    _n, n, ic, h, w, oc, f, s, pd, nl, has_z = case
    d = K.make_desc("qs8", n, ic, h, w, oc, f, stride=s, pad=pd, nonline=nl, has_z=has_z, dst_kind="f32",
                    scales=(0.037, 0.071, None, 1.0, 1.0))
    _check(d, {"int_range": 127}, seed=91 + ic)


F32_CASES = [
    # name, fmt, n, ic, h, w, oc, f, stride, pad, nonline, has_z      (dnn/test/cuda/conv_bias.cpp:83-100 eps 1e-3)
    ("nhwc_3x3_z_relu", C.FORMAT_NHWC, 2, 64, 14, 14, 128, 3, 1, 1, C.NL_RELU, True),
    ("nhwc_1x1_ic20", C.FORMAT_NHWC, 3, 20, 9, 11, 64, 1, 1, 0, C.NL_IDENTITY, False),
    ("nchw_3x3_z_relu", C.FORMAT_NCHW, 2, 32, 14, 14, 64, 3, 1, 1, C.NL_RELU, True),
    ("nchw_7x7_s2_ic3_stem", C.FORMAT_NCHW, 2, 3, 62, 62, 64, 7, 2, 3, C.NL_RELU, False),
    ("nchw_5x5_s2_oc36_hswish", C.FORMAT_NCHW, 2, 12, 17, 17, 36, 5, 2, 2, C.NL_H_SWISH, False),
    ("nchw_1x1_s2_sigmoid_z", C.FORMAT_NCHW, 2, 24, 12, 12, 40, 1, 2, 0, C.NL_SIGMOID, True),
]


@pytest.mark.parametrize("case", F32_CASES, ids=[c[0] for c in F32_CASES])
def test_float32_through_bf16_split(case):
    """Float32 NHWC / NCHW ConvBias on the tensor-core path: operands split into bf16 pieces (x = x0 + x1 +
    O(2^-17)), three products accumulated in fp32, Float32 epilogue -- the reference test's own eps 1e-3."""
    _n, fmt, n, ic, h, w, oc, f, s, pd, nl, has_z = case
    d = K.make_desc("f32", n, ic, h, w, oc, f, stride=s, pad=pd, fmt=fmt, nonline=nl, has_z=has_z)
    plan, info = _check(d, seed=101 + ic)
    assert info["max_rel"] < 2e-4, info     # the split leaves ~2^-16 per product: far inside 1e-3


@pytest.mark.parametrize("layer", K.RESNET18_LAYERS, ids=[l[0] for l in K.RESNET18_LAYERS])
def test_resnet18_fp32_nchw_on_the_gpu_path(layer):
    """BASELINE.json configs[0] (ResNet-18, fp32 NCHW, bias + ReLU) on the GPU path against the naive oracle;
    conv1 on a reduced image for the oracle's sake."""
    d = K.resnet18_desc(layer, 2)
    if layer[0] == "conv1":
        d = K.make_desc("f32", 2, 3, 96, 96, 64, 7, stride=2, pad=3, fmt=C.FORMAT_NCHW, nonline=C.NL_RELU)
    _check(d, seed=111)


F32_DW_CASES = [
    # name, n, c, h, w, stride, pad, nonline, has_z       (3x3: the TMA / bulk-row / global bandwidth kernels)
    ("c32_s1_tma", 3, 32, 14, 14, 1, 1, C.NL_RELU, False),
    ("c96_s2_tma_z", 2, 96, 28, 28, 2, 1, C.NL_RELU, True),
    ("c144_s1_56", 2, 144, 56, 56, 1, 1, C.NL_IDENTITY, False),
    ("c58_s1_bulk", 3, 58, 9, 8, 1, 1, C.NL_RELU, False),
    ("c36_s2_nopad_z", 2, 36, 11, 16, 2, 0, C.NL_H_SWISH, True),
    ("c960_7x7_multi_image", 5, 960, 7, 7, 1, 1, C.NL_RELU, False),
    ("c6_tiny", 2, 6, 5, 5, 1, 1, C.NL_RELU, True),
]


@pytest.mark.parametrize("case", F32_DW_CASES, ids=[c[0] for c in F32_DW_CASES])
def test_float32_depthwise_3x3_bandwidth_kernels(case):
    """Float32 depthwise 3x3 on the bandwidth kernels of the 16-bit kinds (one fp32 channel = one 32-bit word):
    fp32 arithmetic in a different summation order than the oracle, eps 1e-3 (dnn/test/cuda/conv_bias.cpp:650-714)."""
    _n, n, c, h, w, s, pd, nl, has_z = case
    d = K.make_desc("f32", n, c, h, w, c, 3, stride=s, pad=pd, group=c, nonline=nl, has_z=has_z)
    _plan, info = _check(d, seed=121 + c)
    assert info["max_rel"] < 1e-5, info


S8_DW_CASES = [
    # name, n, c, h, w, stride, pad, nonline, has_z, mode      (3x3 QuantizedS8: the DP4A window kernel)
    ("c32_s1_relu", 3, 32, 14, 14, 1, 1, C.NL_RELU, False, 0),
    ("c96_s2_relu_z", 2, 96, 28, 28, 2, 1, C.NL_RELU, True, 0),
    ("c144_s1_identity_z", 2, 144, 56, 56, 1, 1, C.NL_IDENTITY, True, 0),
    ("c24_s2_word_channels", 3, 24, 17, 19, 2, 1, C.NL_RELU, False, 0),
    ("c116_s1_hswish", 2, 116, 14, 14, 1, 1, C.NL_H_SWISH, True, 0),
    ("c36_s2_nopad", 2, 36, 11, 16, 2, 0, C.NL_RELU, True, 0),
    ("c960_7x7", 5, 960, 7, 7, 1, 1, C.NL_RELU, False, 0),
    ("c64_s1_flipped", 2, 64, 9, 13, 1, 1, C.NL_IDENTITY, False, 1),
    ("c8_w3_narrow", 2, 8, 6, 3, 1, 1, C.NL_RELU, True, 0),
]


@pytest.mark.parametrize("case", S8_DW_CASES, ids=[c[0] for c in S8_DW_CASES])
def test_int8_depthwise_3x3_window_kernel(case):
    """QuantizedS8 depthwise 3x3 (C % 4 == 0) on the DP4A window kernel: 0 mismatching bytes, saturation,
    z, every activation, mode CONVOLUTION (flipped filter)."""
    _n, n, c, h, w, s, pd, nl, has_z, mode = case
    d = K.make_desc("qs8", n, c, h, w, c, 3, stride=s, pad=pd, group=c, nonline=nl, has_z=has_z, mode=mode,
                    scales=(0.37, 0.71, None, 1.9, 23.0))
    _plan, info = _check(d, {"int_range": 127, "bias_range": 5000}, seed=131 + c)
    assert info["mismatch"] == 0


FP16_DEFAULT_CASES = [
    ("1x1_relu", 2, 64, 14, 14, 256, 1, 1, 0, C.NL_RELU, False),
    ("3x3_z_relu", 2, 64, 14, 14, 128, 3, 1, 1, C.NL_RELU, True),
    ("3x3_s2_hswish", 3, 32, 15, 15, 64, 3, 2, 1, C.NL_H_SWISH, False),
    ("5x5_ic16_identity", 2, 16, 12, 12, 32, 5, 1, 2, C.NL_IDENTITY, False),
]


@pytest.mark.parametrize("case", FP16_DEFAULT_CASES, ids=[c[0] for c in FP16_DEFAULT_CASES])
def test_fp16_default_compute_mode(case):
    """Float16 with param.compute_mode = DEFAULT (what imperative F.conv2d sends): the naive oracle accumulates in
    fp16, the B200 kernel -- like the reference's CUDA algorithms -- in fp32; the reference's CUDA tests compare the
    two at 2e-2 (dnn/test/cuda/conv_bias.cpp:242-247,663-664) on shapes with K = IC * FH * FW up to a few hundred,
    and so does this one (at K = 2304 the oracle's own fp16 accumulation error exceeds that bound).  Against the fp32
    reference the kernel keeps its usual 2e-3."""
    from . import _gpu as G
    _name, n, ic, h, w, oc, f, s, pd, nl, has_z = case
    d = K.make_desc("f16", n, ic, h, w, oc, f, stride=s, pad=pd, nonline=nl, has_z=has_z, compute_mode=C.COMPUTE_DEFAULT)
    src, flt, bias, z = K.gen_inputs(d, 99)
    want = O.oracle_convbias(d, src, flt, bias, z)            # fp16 accumulation
    got, _plan = G.run_gpu(d, src, flt, bias, z)
    err = O.megdnn_rel_err(O.to_f32(got, C.DT_F16), O.to_f32(want, C.DT_F16)).max()
    assert err <= 2e-2, (K.desc_str(d), err)
    d32 = d.copy()
    d32.compute_mode = C.COMPUTE_FLOAT32
    ok, info = G.compare(d32, got, None, K.fp32_reference(d32, src, flt, bias, z))
    assert ok, info

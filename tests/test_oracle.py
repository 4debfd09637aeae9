"""CPU tests that PIN the oracle (oracle/convbias_oracle.c, the C restatement of megdnn's naive
ConvBiasForward) to the reference:

  (a) the reference's known-answer test dnn/test/naive/conv_bias.cpp:26-64;
  (b) the committed golden fixtures tests/golden/convbias_golden.npz, which are outputs of the
      reference's own compiled naive operator (tests/golden/make_golden.py);
  (c) when oracle/_ref is present (build container), a live differential sweep against the
      reference operator over the arg generators of dnn/test/common/conv_bias.cpp:19-107,306-402.
"""
import numpy as np
import pytest

from megengine_b200 import _capi as C
from . import _cases as K
from . import _oracle as O
from .golden import make_golden as G


def test_reference_kat_quantized8x8x32():
    """dnn/test/naive/conv_bias.cpp:26-64 CONV_BIAS_QUANTIZED8x8x32 (NCHW, QS8 -> QS32 with QS32 z)."""
    d = K.make_desc("qs8", 1, 1, 4, 4, 3, 3, fmt=C.FORMAT_NCHW, dst_kind="qs32", has_z=True,
                    scales=(0.1, 0.2, None, 0.3, None))
    src = (np.array([90, 136, 85, 204, 48, 9, 226, 25, 118, 109, 87, 132, 104, 163, 25, 90]) - 128).astype(np.int8)
    flt = (np.array([153, 170, 102, 103, 23, 213, 116, 195, 191, 44, 50, 247, 172, 42, 32, 233, 163, 247, 120,
                     241, 209, 83, 201, 115, 32, 140, 147]) - 124).astype(np.int8)
    bias = np.zeros(3, np.int32)
    z = np.array([1234, 0, 0, 0, 0, 0, 0, 0, 0, -234, 0, 0], np.int32)
    want = np.array([37127, -22475, -15694, -1920, -12813, 4440, 18190, -13195, -9659, 12423, -5558, -4969], np.int32)
    got = O.oracle_convbias(d, src, flt, bias, z)
    assert np.array_equal(got, want)


def test_survey_probe_all_ones_bf16_f16():
    """SURVEY Appendix C probe: all-ones src(1,4,4,8) flt(8,3,3,8) + bias 1 + z 1, pad 1, RELU ->
    corner 34.0 (bf16 0x4208 / f16 0x5040), centre 74.0 (0x4294 / 0x54a0)."""
    for kind, corner, centre in (("bf16", 0x4208, 0x4294), ("f16", 0x5040, 0x54a0)):
        d = K.make_desc(kind, 1, 8, 4, 4, 8, 3, pad=1, nonline=C.NL_RELU, has_z=True)
        one = O.from_f32(np.ones(1, np.float32), d.src_dtype)[0]
        mk = lambda n: np.full(n, one, np.uint16)
        got = O.oracle_convbias(d, mk(128), mk(8 * 72), mk(8), mk(128)).reshape(4, 4, 8)
        assert got[0, 0, 0] == corner and got[1, 1, 0] == centre


@pytest.mark.parametrize("meta,want", G.load(), ids=[m["name"] for m, _ in G.load()])
def test_oracle_matches_golden(meta, want):
    """restatement == reference output, bit for bit, on every golden case (all dtypes)."""
    d = G.dict_to_desc(meta["desc"])
    opt = meta["opt"]
    src, flt, bias, z = K.gen_inputs(d, meta["seed"], int_range=opt.get("int_range", 8),
                                     bias_range=opt.get("bias_range", 50))
    import zlib
    assert zlib.crc32(src.tobytes()) == meta["src_crc"], "input generator drifted from the fixture"
    got = O.oracle_convbias(d, src, flt, bias, z)
    assert got.dtype == want.dtype
    assert np.array_equal(got, want), "first diff at %d" % int(np.nonzero(got != want)[0][0])


def test_oracle_thread_count_independent():
    d = K.make_desc("bf16", 2, 32, 9, 9, 32, 3, pad=1, nonline=C.NL_H_SWISH, has_z=True)
    src, flt, bias, z = K.gen_inputs(d, 3)
    a = O.oracle_convbias(d, src, flt, bias, z, threads=1)
    b = O.oracle_convbias(d, src, flt, bias, z, threads=7)
    assert np.array_equal(a, b)


def test_oracle_sample_equals_full():
    d = K.make_desc("qs8", 2, 32, 9, 9, 32, 3, pad=1, nonline=C.NL_RELU, has_z=True)
    src, flt, bias, z = K.gen_inputs(d, 4)
    full = O.oracle_convbias(d, src, flt, bias, z)
    idx = np.random.default_rng(0).integers(0, full.size, 200)
    assert np.array_equal(O.oracle_convbias_sample(d, src, flt, bias, z, idx), full[idx])


def test_empty_and_degenerate_shapes():
    """1x1 image, filter == image (single output pixel), batch 1, channel 1."""
    for d in (K.make_desc("f32", 1, 1, 1, 1, 1, 1, fmt=C.FORMAT_NCHW),
              K.make_desc("qs8", 1, 16, 3, 3, 16, 3),
              K.make_desc("bf16", 1, 8, 5, 7, 8, 5, fw=7)):
        src, flt, bias, z = K.gen_inputs(d, 1)
        got = O.oracle_convbias(d, src, flt, bias, z)
        assert got.size == d.n * d.oc * d.oh * d.ow and d.oh == 1 and d.ow == 1


def _sweep_cases():
    """small differential sweep in the spirit of get_conv_bias_args / get_chanwise_args
    (dnn/test/common/conv_bias.cpp:19-107,306-402): kernel x stride x pad x nonlinearity x z."""
    out = []
    rng = np.random.default_rng(99)
    for kind in ("f32", "f16", "bf16", "qs8"):
        for fmt in (C.FORMAT_NHWC, C.FORMAT_NCHW) + ((C.FORMAT_NCHW4,) if kind == "qs8" else ()):
            for f in (1, 2, 3, 5):
                for stride in (1, 2):
                    pad = int(rng.integers(0, f // 2 + 1))
                    nl = int(rng.choice([0, 1, 3]))
                    has_z = bool(rng.integers(0, 2))
                    ic, oc = (8, 12) if fmt == C.FORMAT_NCHW4 else (int(rng.integers(1, 9)), int(rng.integers(1, 9)))
                    bias_mode = int(rng.choice([C.BIAS_NONE, C.BIAS_CHANNEL, C.BIAS_CHANNEL]))
                    out.append(K.make_desc(kind, 2, ic, 7, 8, oc, f, stride=stride, pad=pad, nonline=nl, has_z=has_z,
                                           fmt=fmt, bias_mode=bias_mode, scales=(0.31, 0.17, None, 0.9, 1.7)))
            # grouped / depthwise, flipped, dilated
            if fmt != C.FORMAT_NCHW4:
                out.append(K.make_desc(kind, 2, 8, 7, 7, 8, 3, pad=1, group=8, fmt=fmt, nonline=1))
                out.append(K.make_desc(kind, 1, 8, 7, 7, 12, 3, pad=1, group=4, fmt=fmt))
                out.append(K.make_desc(kind, 1, 4, 7, 7, 4, 3, pad=1, mode=1, fmt=fmt))
                out.append(K.make_desc(kind, 1, 4, 9, 9, 4, 3, pad=2, dilate=2, fmt=fmt, has_z=True))
            if kind == "qs8" and fmt != C.FORMAT_NCHW4:
                # QuantizedS8 -> Float32 (fp32 bias / z, dequantised products accumulated in fp32)
                out.append(K.make_desc("qs8", 2, 16, 7, 7, 16, 3, pad=1, fmt=fmt, nonline=1, has_z=True, dst_kind="f32"))
                out.append(K.make_desc("qs8", 1, 8, 5, 6, 12, 1, fmt=fmt, nonline=3, dst_kind="f32"))
            # depthwise 3x3 as the GPU parity cases run it (get_chanwise_args, dnn/test/cuda/conv_bias.cpp:650-714):
            # channel counts that are not vector multiples, stride 2 on odd extents, no padding, z
            if fmt == C.FORMAT_NHWC and kind in ("f16", "bf16"):
                out.append(K.make_desc(kind, 3, 58, 9, 8, 58, 3, stride=1, pad=1, group=58, fmt=fmt, nonline=1))
                out.append(K.make_desc(kind, 2, 36, 11, 16, 36, 3, stride=2, pad=0, group=36, fmt=fmt, nonline=1, has_z=True))
                out.append(K.make_desc(kind, 2, 12, 9, 8, 12, 3, stride=1, pad=0, group=12, fmt=fmt, nonline=3))
    return out


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (compiled reference) not present on this box")
def test_oracle_matches_live_reference_sweep():
    bad = []
    cases = _sweep_cases()
    for i, d in enumerate(cases):
        src, flt, bias, z = K.gen_inputs(d, 500 + i, int_range=127 if i % 3 == 0 else 8)
        want, _s, _a = O.ref_convbias(d, src, flt, bias, z, handle_kind=2)
        got = O.oracle_convbias(d, src, flt, bias, z)
        if not np.array_equal(got, want):
            bad.append(K.desc_str(d))
    assert not bad, "%d / %d cases differ from the reference: %s" % (len(bad), len(cases), bad[:5])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (compiled reference) not present on this box")
def test_x86_baseline_agrees_with_naive():
    """the CPU baseline handle (x86/fallback algos) computes the same convolution (fp32 1e-3,
    dnn/test/cuda/conv_bias.cpp:96)."""
    d = K.make_desc("f32", 2, 16, 14, 14, 32, 3, pad=1, nonline=C.NL_RELU, fmt=C.FORMAT_NCHW)
    src, flt, bias, _ = K.gen_inputs(d, 8)
    naive, _s, _a = O.ref_convbias(d, src, flt, bias, None, handle_kind=2)
    x86, _s, algo = O.ref_convbias(d, src, flt, bias, None, handle_kind=0, threads=2)
    assert algo != "DEFAULT"
    assert O.megdnn_rel_err(naive, x86).max() < 1e-3


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (compiled reference) not present on this box")
@pytest.mark.parametrize("layer", K.RESNET18_LAYERS, ids=[l[0] for l in K.RESNET18_LAYERS])
def test_resnet18_fp32_nchw_cpu_config(layer):
    """BASELINE.json configs[0] (the reference's own CPU-runnable case): every ConvBias shape of ResNet-18,
    fp32 NCHW, bias + ReLU, at N = 1 -- the C restatement equals the reference's naive operator bit for bit,
    and the reference's x86 / fallback fast path (the CPU baseline arm) agrees with it to the reference
    tests' fp32 eps (dnn/test/cuda/conv_bias.cpp:96)."""
    d = K.resnet18_desc(layer, 1)
    if layer[0] == "conv1":   # naive at 224x224 takes a while: same filter / stride / pad on 96x96
        d = K.make_desc("f32", 1, 3, 96, 96, 64, 7, stride=2, pad=3, fmt=C.FORMAT_NCHW, nonline=C.NL_RELU)
    src, flt, bias, _ = K.gen_inputs(d, 18)
    naive, _s, _a = O.ref_convbias(d, src, flt, bias, None, handle_kind=2)
    port = O.oracle_convbias(d, src, flt, bias, None)
    assert np.array_equal(port, naive)
    x86, _s, algo = O.ref_convbias(d, src, flt, bias, None, handle_kind=0, threads=4)
    assert O.megdnn_rel_err(naive, x86).max() < 1e-3, algo

"""CPU tests of the C-ABI boundary (include/b200cb.h): the library loads without a GPU, exports
every declared symbol, and the pure-host entry points (algo table, query/is_available matrix,
workspace + accounting, error text) behave as the megdnn AlgoBase contract requires
(dnn/src/cuda/conv_bias/algo.h:83-165).  No compute call is made here."""
import ctypes
import os
import re

import pytest

from megengine_b200 import _capi as C
from . import _cases as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="b200cb.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200cb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "libb200cb.so does not export " + s
    assert sorted(n for n, _r, _a in C.SYMBOLS) == syms, "ctypes binding and header disagree"
    nn = declared_symbols("b200cb_nn.h")
    assert len(nn) >= 10
    for s in nn:
        assert hasattr(lib, s), "libb200cb.so does not export " + s
    assert sorted(n for n, _r, _a in C.NN_SYMBOLS) == nn, "ctypes binding and b200cb_nn.h disagree"


def test_struct_sizes_match_header(lib):
    # 33 x 4-byte fields + 3 x u64 (8-aligned) ; plan = 16 x u32 + 3 x u64 + desc
    assert ctypes.sizeof(C.Desc) == 160
    assert ctypes.sizeof(C.Plan) == 64 + 24 + 160


def test_algo_table_is_stable(lib):
    names = [lib.b200cb_algo_name(i).decode() for i in range(lib.b200cb_algo_count())]
    # append-only: megdnn serialises Algorithm::Info::Desc into the fastrun cache (oprs/base.h:135-158)
    assert names[:4] == ["B200_IGEMM_TCGEN05_128x256", "B200_IGEMM_TCGEN05_128x128",
                         "B200_IGEMM_TCGEN05_128x64", "B200_DWCONV_DIRECT"]
    assert len(set(names)) == len(names) and all(":" not in n for n in names)
    assert lib.b200cb_algo_name(-1) is None and lib.b200cb_algo_name(len(names)) is None


def _query(lib, d, algo=None):
    p = C.Plan()
    rc = lib.b200cb_query(ctypes.byref(d), ctypes.byref(p)) if algo is None else \
        lib.b200cb_query_algo(ctypes.byref(d), algo, ctypes.byref(p))
    return rc, p


def test_query_resnet50_all_layers_supported(lib):
    for kind in ("bf16", "f16", "qs8"):
        for layer in K.RESNET50_LAYERS:
            d = K.resnet50_desc(layer, kind, 256, ic_pad=16 if kind == "qs8" else 8)
            rc, p = _query(lib, d)
            assert rc == 0, (layer[0], kind, lib.b200cb_strerror(rc))
            assert p.smem_bytes <= 232448 and p.stages >= 2
            n, oc, oh, ow, ic, f = d.n, d.oc, d.oh, d.ow, d.ic, d.fh
            assert p.flops == 2 * n * oc * oh * ow * ic * f * f   # conv_test_utils.cpp:193
            if layer[0] == "conv1":
                # W-folded stem: 7x7 s2 over IC -> 7x4 s1 over 2*IC; re-packed filter in the workspace
                packed = oc * 7 * 4 * 2 * ic * (1 if kind == "qs8" else 2)
                assert lib.b200cb_preprocessed_filter_bytes(ctypes.byref(p)) == packed
                assert packed <= p.workspace_bytes < packed + 256
                assert lib.b200cb_plan_num_launches(ctypes.byref(p)) == 2
                assert lib.b200cb_plan_use_preprocessed_filter(ctypes.byref(p)) == 0
            assert p.workspace_bytes == 0 and lib.b200cb_plan_num_launches(ctypes.byref(p)) == 1


def test_query_mobilenet_depthwise_supported(lib):
    for kind in ("bf16", "f16", "f32", "qs8"):
        for c, h, w, s in K.MOBILENETV2_DW:
            if kind == "qs8" and c % 16:
                continue
            rc, p = _query(lib, K.dw_desc(kind, 256, c, h, w, s))
            assert rc == 0 and lib.b200cb_algo_name(p.algo) == b"B200_DWCONV_DIRECT"


def test_query_shufflenet_depthwise_channel_counts(lib):
    """ShuffleNetV2-1x depthwise layers: 58 / 116 channels are not 16-byte pixels; the
    16-bit kinds take them (bulk-row / register front ends), QuantizedS8 and fp32 need whole 16-byte vectors."""
    shapes = [(24, 56, 56, 2), (58, 56, 56, 2), (58, 28, 28, 1), (116, 28, 28, 2), (116, 14, 14, 1),
              (232, 14, 14, 2), (232, 7, 7, 1)]
    for kind in ("bf16", "f16"):
        for c, h, w, s in shapes:
            rc, p = _query(lib, K.dw_desc(kind, 256, c, h, w, s))
            assert rc == 0 and lib.b200cb_algo_name(p.algo) == b"B200_DWCONV_DIRECT", (kind, c)
            assert p.workspace_bytes == 0 and lib.b200cb_plan_num_launches(ctypes.byref(p)) == 1
            assert p.bytes == 2 * 256 * c * (h * w + (h // s) * (w // s)) + 2 * c * 9 + 2 * c
    assert _query(lib, K.dw_desc("qs8", 256, 58, 28, 28, 1))[0] > 0
    assert _query(lib, K.dw_desc("f32", 256, 58, 28, 28, 1))[0] == 0   # fp32 3x3: one channel per 32-bit word
    assert _query(lib, K.dw_desc("f32", 256, 58, 28, 28, 1, f=5))[0] > 0  # other filters: 16-byte vectors
    assert _query(lib, K.dw_desc("bf16", 256, 57, 28, 28, 1))[0] > 0   # odd channel count: no channel pairs


def test_unsupported_cells_are_rejected_not_silently_wrong(lib):
    """everything outside the supported-cell matrix must come back > 0 (is_available false)."""
    ok = K.make_desc("bf16", 2, 64, 14, 14, 64, 3, pad=1)
    assert _query(lib, ok)[0] == 0
    cases = []
    d = ok.copy(); d.format = 2; cases.append(d)                                   # format (NHWCD4)
    d = K.make_desc("f32", 2, 64, 14, 14, 62, 3, pad=1, fmt=C.FORMAT_NCHW); cases.append(d)  # fp32 path: OC % 4
    d = K.make_desc("bf16", 2, 64, 14, 14, 60, 3, pad=1, fmt=C.FORMAT_NCHW); cases.append(d)  # NCHW adapter needs OC % 8
    d = ok.copy(); d.src_dtype = d.filter_dtype = C.DT_F32; d.dst_dtype = C.DT_F32; cases.append(d)  # fp32 conv, bf16 bias
    d = ok.copy(); d.ic = 60; cases.append(d)                                      # 16-byte channel alignment
    d = ok.copy(); d.bias_mode = C.BIAS_FULL; d.has_z = 1; d.z_dtype = d.dst_dtype; cases.append(d)  # full bias AND z
    d = K.make_desc("qs8", 2, 64, 14, 14, 64, 3, pad=1, bias_mode=C.BIAS_FULL); cases.append(d)  # QS32 full bias
    d = ok.copy(); d.filter_dtype = C.DT_F16; cases.append(d)                      # src/filter mismatch
    d = ok.copy(); d.oh += 1; cases.append(d)                                      # inconsistent output extent
    d = ok.copy(); d.stride_h = 0; cases.append(d)
    d = K.make_desc("qs8", 2, 64, 14, 14, 64, 3, pad=1, nonline=C.NL_SIGMOID); cases.append(d)
    d = K.make_desc("qs8", 2, 64, 14, 14, 64, 3, pad=1, dst_kind="qs32", nonline=C.NL_RELU); cases.append(d)
    d = K.make_desc("bf16", 2, 24, 14, 14, 64, 3, pad=1, group=2); cases.append(d)  # group slices of 24 bytes
    d = K.make_desc("qs8", 2, 64, 14, 14, 64, 3, pad=1, dst_kind="f32"); d.bias_dtype = C.DT_QS32; cases.append(d)
    for d in cases:
        rc, _p = _query(lib, d)
        assert rc > 0, K.desc_str(d)
        assert lib.b200cb_strerror(rc) not in (b"ok", b"unknown error")


def test_new_cells_of_round2_are_available(lib):
    """a13 cells closed in round 2: general GROUP conv (one dense launch per group on channel-sliced views),
    BiasMode::BIAS through the z ring, QuantizedS8 -> Float32, Float32 NHWC / NCHW through the bf16 x 3 split."""
    d = K.make_desc("bf16", 2, 64, 14, 14, 128, 3, pad=1, group=4, nonline=C.NL_RELU, has_z=True)
    rc, p = _query(lib, d)
    assert rc == 0 and lib.b200cb_plan_num_launches(ctypes.byref(p)) == 4
    assert p.flops == 2 * 2 * 128 * 14 * 14 * 16 * 9
    rc, p = _query(lib, K.make_desc("f16", 2, 64, 14, 14, 64, 3, pad=1, bias_mode=C.BIAS_FULL))
    assert rc == 0 and p.workspace_bytes == 0
    rc, p = _query(lib, K.make_desc("qs8", 2, 64, 14, 14, 64, 3, pad=1, dst_kind="f32", has_z=True, nonline=C.NL_H_SWISH))
    assert rc == 0
    for fmt, nl in ((C.FORMAT_NHWC, 2), (C.FORMAT_NCHW, 6)):
        d = K.make_desc("f32", 2, 20, 14, 14, 64, 3, pad=1, fmt=fmt, has_z=True, nonline=C.NL_RELU)
        rc, p = _query(lib, d)
        assert rc == 0, lib.b200cb_strerror(rc)
        a = lambda v: (v + 255) & ~255
        assert p.workspace_bytes >= a(2 * 14 * 14 * 72 * 2) + a(64 * 9 * 72 * 2)   # 3 x 24 bf16 channels
        assert lib.b200cb_plan_num_launches(ctypes.byref(p)) == nl + 1
    # BASELINE configs[0]: every ResNet-18 fp32 NCHW layer has a plan
    for layer in K.RESNET18_LAYERS:
        assert _query(lib, K.resnet18_desc(layer, 8))[0] == 0, layer[0]


def test_nchw_16bit_plans_carry_the_relayout_workspace(lib):
    """NCHW fp16 / bf16 runs through the relayout adapter: the plan's workspace holds the NHWC copies of
    src / filter / z / dst (channels padded for TMA) and the launch count includes the re-tiling passes."""
    d = K.make_desc("bf16", 2, 20, 14, 14, 64, 3, pad=1, fmt=C.FORMAT_NCHW, has_z=True)
    rc, p = _query(lib, d)
    assert rc == 0
    icp, a = 24, lambda v: (v + 255) & ~255
    need = a(2 * 14 * 14 * icp * 2) + a(64 * 9 * icp * 2) + 2 * a(2 * 14 * 14 * 64 * 2)
    assert p.workspace_bytes >= need
    assert lib.b200cb_plan_num_launches(ctypes.byref(p)) == 5   # src, filter, z, conv, dst
    assert lib.b200cb_preprocessed_filter_bytes(ctypes.byref(p)) == 0


def test_query_algo_enumeration_and_heuristic_pick(lib):
    d = K.make_desc("bf16", 256, 256, 14, 14, 256, 3, pad=1)
    avail = [a for a in range(lib.b200cb_algo_count()) if _query(lib, d, a)[0] == 0]
    assert avail == [0, 1, 2]
    rc, p = _query(lib, d)
    assert rc == 0 and p.algo in avail
    # narrow OC never gets a tile more than 2x wider than OC
    d = K.make_desc("bf16", 4, 64, 14, 14, 64, 1)
    assert [a for a in range(3) if _query(lib, d, a)[0] == 0] == [2]
    assert _query(lib, d, 99)[0] == 6


def test_launch_argument_validation_without_gpu(lib):
    """argument errors are reported as negative codes before anything touches CUDA."""
    d = K.make_desc("bf16", 2, 64, 14, 14, 64, 3, pad=1)
    rc, p = _query(lib, d)
    assert rc == 0
    assert lib.b200cb_launch(None, None, None, None, None, None, None, 0, None) == -5
    bad = C.Plan()
    assert lib.b200cb_launch(ctypes.byref(bad), None, None, None, None, None, None, 0, None) == -4
    assert lib.b200cb_launch(ctypes.byref(p), None, None, None, None, None, None, 0, None) == -5  # null tensors
    assert lib.b200cb_query(None, ctypes.byref(p)) == -5
    assert lib.b200cb_host_scratch_bytes(ctypes.byref(p)) >= 2 * (2 * 14 * 14 * 64) * 2
    assert lib.b200cb_preprocessed_filter_bytes(ctypes.byref(p)) == 0
    for code in (0, 1, 2, 3, 4, 5, 6, -1, -2, -3, -4, -5):
        assert lib.b200cb_strerror(code)


def test_product_has_no_oracle_or_cpu_fallback():
    """the shipped package must not import / link anything under oracle/ (task rule 3)."""
    pkg = os.path.join(ROOT, "megengine_b200")
    for dirpath, _dn, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".sh", ".mk", ".txt", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("liboracle", "ref_driver", "_oracle", "oracle/", "oracle_convbias", "libmegdnn"):
                    assert needle not in text, (f, needle)


def test_int8_epilogue_sass_has_no_contracted_packed_ops():
    """Bit-exactness guard.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (see
    megengine_b200/csrc/epilogue.cuh), so the packed requantize keeps no product that feeds an add.  In
    every QuantizedS8 igemm kernel the machine code must then hold exactly the explicit FMAs of the
    division (2 FFMA2 per FADD2.RM pair rounding) and of the conversion-free int32 -> float (one FFMA2 per pair in
    the full-range variants) plus the one explicit FMA-by-one of the z dequantisation; the only round-to-nearest FADD2s are
    the exact magic-number subtractions -- an add never follows a product."""
    import re
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "megengine_b200", "lib", "libb200cb.so")
    if shutil.which("cuobjdump") is None or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library not available")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    checked = 0
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        is_igemm8 = "igemm_kernelILi2E" in name and "ELi1ELi" not in name.split("igemm_kernelILi2E")[1][:12]
        is_dw8 = "dwconv_nhwc_kernelILi3E" in name
        if not (is_igemm8 or is_dw8):
            continue  # QuantizedS8 -> QuantizedS8 kernels only (igemm KIND 2 / OUT32 0, depthwise KIND 3)
        ffma2, rm, rz = f.count("FFMA2"), f.count("FADD2.RM"), f.count("FADD2.RZ")
        plain = len(re.findall(r"FADD2 ", f))
        # one FADD2.RZ (the +-0.5 of the rounding) per output pair; the clamp and the floor are a saturating
        # conversion, so no FADD2.RM is left (epilogue.cuh, B200CB_RQ_CVT)
        assert rz > 0 and rm == 0, (name, rm, rz)
        # per pair: the two FMAs of the division; the sites that convert int32 -> float on the FMA / ALU pipes over
        # the full range (rows-mode epilogue: 2 of the 14 igemm sites, 16 pairs each) add one FMA and two exact
        # FADD2s, the range-proven magic sites (4 igemm sites) one exact FADD2.  A contracted mul + add would add an
        # FFMA2 and remove a FADD2.
        n_split, n_magic = (32, 64) if is_igemm8 else (0, 0)
        # sites with a residual z (6 of the 14 igemm sites incl. the two NCHW4 ones; every second depthwise site):
        # the dequantisation RN(RN(a * s_bias) + RN(z * s_z)) is two FMUL2s and ONE explicit FFMA2 by a run-time 1.0
        # (rounds like the add, cannot be contracted), plus the exact FADD2 that turns the z bytes into floats
        n_z = 96 if is_igemm8 else rz // 2
        fmul2 = f.count("FMUL2")
        assert ffma2 == 2 * rz + n_split + n_z, (name, ffma2, rz)
        assert plain == n_magic + 2 * n_split + n_z, (name, plain)
        assert fmul2 == 2 * rz + n_z, (name, fmul2)   # a * s_bias and x * (1 / s_dst) per pair, z * s_z per z pair
        checked += 1
    assert checked >= 11, checked


def test_adjacent_operator_argument_validation_without_gpu(lib):
    """include/b200cb_nn.h: argument errors are reported before anything touches CUDA -- null pointers, unknown
    modes, shapes the reference's deduction rejects, windows entirely in the padding, unaligned tensors."""
    d = C.PoolDesc()
    d.format, d.mode, d.dtype = C.FORMAT_NHWC, C.POOL_MAX, C.DT_BF16
    d.n, d.c, d.ih, d.iw, d.oh, d.ow = 1, 8, 8, 8, 4, 4
    d.window_h = d.window_w = d.stride_h = d.stride_w = 2
    assert lib.b200cb_pooling(ctypes.byref(d), None, None, None) == -5
    assert lib.b200cb_pooling(ctypes.byref(d), 4096, 8192 + 2, None) == 5          # dst not 16-byte aligned
    d.oh = 5
    assert lib.b200cb_pooling(ctypes.byref(d), 4096, 8192, None) == 3              # not the deduced extent
    d.oh, d.pad_h, d.pad_w = 6, 2, 2
    d.ow = 6
    assert lib.b200cb_pooling(ctypes.byref(d), 4096, 8192, None) == 4              # a window lies in the padding
    assert lib.b200cb_elemwise(12345, C.DT_F32, 16, 4096, 4096, 8192, None) == 4   # unknown mode
    assert lib.b200cb_elemwise(C.ELEM_ADD, C.DT_F32, 16, 4096, None, 8192, None) == -5
    assert lib.b200cb_elemwise(C.ELEM_ADD, C.DT_QS8, 16, 4096, 4096, 8192, None) == 2
    assert lib.b200cb_elemwise(C.ELEM_RELU, C.DT_F32, 0, 4096, None, 8192, None) == 0  # empty tensor: nothing to do
    assert lib.b200cb_elemwise_q(C.ELEM_MUL, 16, 4096, 1.0, 4096, 1.0, 8192, 1.0, None) == 4
    assert lib.b200cb_typecvt(C.DT_F16, 0.0, C.DT_BF16, 0.0, 16, 4096, 8192, None) == 2   # pair not implemented
    assert lib.b200cb_typecvt(C.DT_F32, 0.0, C.DT_BF16, 0.0, 16, 4096 + 4, 8192, None) == 5
    # deconvolution: the descriptor names the forward convolution; extents must agree with it
    f = K.make_desc("bf16", 2, 16, 12, 12, 32, 3, stride=2, pad=1, bias_mode=C.BIAS_NONE)
    f.bias_dtype = C.DT_NONE
    p = C.Plan()
    assert lib.b200cb_deconv_query(ctypes.byref(f), ctypes.byref(p)) == 0
    assert p.workspace_bytes > 0 and p.reserved[0] == 1           # stride 2: phase decomposition
    assert p.flops == 2 * 2 * 32 * 6 * 6 * 16 * 9
    f.oh = 7
    assert lib.b200cb_deconv_query(ctypes.byref(f), ctypes.byref(p)) == 3
    g = K.make_desc("bf16", 2, 16, 12, 12, 32, 1, pad=1, bias_mode=C.BIAS_NONE)
    g.bias_dtype = C.DT_NONE
    assert lib.b200cb_deconv_query(ctypes.byref(g), ctypes.byref(p)) == 4  # pad > dilated filter - 1: cropping, not implemented
    # ConvolutionForward ignores the epilogue fields; BatchConvBias plans one image
    c = K.make_desc("bf16", 4, 64, 14, 14, 128, 3, pad=1, nonline=C.NL_RELU, has_z=True)
    assert lib.b200cb_convolution_query(ctypes.byref(c), ctypes.byref(p)) == 0
    assert p.desc.has_z == 0 and p.desc.bias_mode == C.BIAS_NONE and p.desc.nonline_mode == C.NL_IDENTITY
    assert lib.b200cb_batch_convbias_query(ctypes.byref(c), ctypes.byref(p)) == 0
    assert p.desc.n == 4 and p.flops == 2 * 4 * 128 * 14 * 14 * 64 * 9

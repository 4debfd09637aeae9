"""Graph-level composition on the GPU: a QuantizedS8 NHWC ResNet-style network -- stem conv
7x7/2 -> max pool 3x3/2 -> bottleneck with a projection shortcut (the residual enters the last conv as z) ->
identity bottleneck -> QuantizedS8 -> global average pool -- built with megengine_b200.graph.Sequential, every
ConvBias algorithm chosen by profiling (fastrun), executed eagerly and as one captured CUDA graph, and compared
bit for bit with the chain of oracles (each consuming the previous oracle's output)."""
import ctypes

import numpy as np
import pytest
import torch

import megengine_b200 as M
from megengine_b200 import _capi as C
from megengine_b200.graph import AlgoChooser, Sequential
from . import _cases as K
from . import _gpu as G
from . import _nn as N
from . import _oracle as O

pytestmark = pytest.mark.gpu

Q8, Q32 = M.dtype.QuantizedS8, M.dtype.QuantizedS32


def _conv_ref(x, xs, w, ws, b, z, zs, ds, n, ic, h, wd, oc, f, stride, pad, nl):
    d = K.make_desc("qs8", n, ic, h, wd, oc, f, stride=stride, pad=pad, nonline=nl, has_z=z is not None,
                    scales=(xs, ws, None, zs if z is not None else 1.0, ds))
    return O.oracle_convbias(d, x, w, b, z), d


def test_int8_resnet_style_graph_bit_exact_with_fastrun_and_cuda_graph():
    rng = np.random.default_rng(7)
    n, H = 2, 64
    handle = M.Handle("cuda:0")
    net = Sequential(handle, fastrun=True)
    ref = {}
    scales = {"x": 0.05}

    def tensor(name, arr, shape, dt):
        ref[name] = arr
        t = G.to_dev(arr).view(dt.torch_dtype).reshape(shape)
        net.add_tensor(name, M.TensorND(t, M.TensorLayout(tuple(shape), dt)))

    def weights(name, oc, f, ic, scale):
        tensor(name + ".w", rng.integers(-20, 21, oc * f * f * ic).astype(np.int8), (oc, f, f, ic), Q8(scale))
        scales[name + ".w"] = scale

    def bias(name, oc, xs, ws):
        tensor(name + ".b", rng.integers(-3000, 3000, oc).astype(np.int32), (1, 1, 1, oc), Q32(K.f32(xs) * K.f32(ws)))

    shapes = {}

    def conv(name, src, ic, oc, f, stride, pad, nl, dscale, z=None):
        h, w = shapes[src]
        xs, ws = scales[src], 0.02
        weights(name, oc, f, ic, ws)
        bias(name, oc, xs, ws)
        net.conv_bias(name, src, name + ".w", name + ".b", z or "", dst_dtype=Q8(dscale), format=C.FORMAT_NHWC,
                      pad_h=pad, pad_w=pad, stride_h=stride, stride_w=stride, nonlineMode=nl)
        out, d = _conv_ref(ref[src], xs, ref[name + ".w"], ws, ref[name + ".b"], ref[z] if z else None,
                           scales[z] if z else None, dscale, n, ic, h, w, oc, f, stride, pad, nl)
        ref[name], scales[name], shapes[name] = out, dscale, (d.oh, d.ow)

    # network input: 8-byte pixels (3 channels padded to 8), as the stem kernel wants
    x = rng.integers(-100, 101, n * H * H * 8).astype(np.int8).reshape(n, H, H, 8)
    x[..., 3:] = 0
    tensor("x", x.reshape(-1), (n, H, H, 8), Q8(scales["x"]))
    shapes["x"] = (H, H)
    conv("stem", "x", 8, 64, 7, 2, 3, C.NL_RELU, 0.9)
    # max pool 3x3 / 2
    net.pooling("pool", "stem", mode=C.POOL_MAX, format=C.FORMAT_NHWC, window_h=3, window_w=3, stride_h=2, stride_w=2,
                pad_h=1, pad_w=1)
    h, w = shapes["stem"]
    pd = N.pool_desc(C.FORMAT_NHWC, C.POOL_MAX, C.DT_QS8, n, 64, h, w, 3, 2, 1)
    ref["pool"], scales["pool"], shapes["pool"] = N.pooling("oracle", pd, ref["stem"]), scales["stem"], (pd.oh, pd.ow)
    # bottleneck with projection shortcut
    conv("b0.c1", "pool", 64, 64, 1, 1, 0, C.NL_RELU, 0.7)
    conv("b0.c2", "b0.c1", 64, 64, 3, 1, 1, C.NL_RELU, 0.8)
    conv("b0.ds", "pool", 64, 256, 1, 1, 0, C.NL_IDENTITY, 1.1)
    conv("b0.c3", "b0.c2", 64, 256, 1, 1, 0, C.NL_RELU, 1.3, z="b0.ds")
    # identity bottleneck, stride 2 in the 3x3 (ResNet v1.5) needs a strided shortcut: keep stride 1 here
    conv("b1.c1", "b0.c3", 256, 64, 1, 1, 0, C.NL_RELU, 0.9)
    conv("b1.c2", "b1.c1", 64, 64, 3, 1, 1, C.NL_RELU, 0.8)
    conv("b1.c3", "b1.c2", 64, 256, 1, 1, 0, C.NL_RELU, 1.7, z="b0.c3")
    # residual add outside a conv (what gopt leaves when z-fusion is not legal) + global average pool
    net.elemwise("sum", M.ElemwiseMultiType.Mode.QFUSE_ADD_RELU, "b1.c3", "b0.c3", dst_dtype=Q8(2.1))
    ref["sum"] = N.elemwise_q("oracle", C.ELEM_FUSE_ADD_RELU, ref["b1.c3"], scales["b1.c3"], ref["b0.c3"], scales["b0.c3"], 2.1)
    h, w = shapes["b1.c3"]
    net.pooling("feat", "sum", mode=C.POOL_AVERAGE, format=C.FORMAT_NHWC, window_h=h, window_w=w, stride_h=h, stride_w=w)
    gd = N.pool_desc(C.FORMAT_NHWC, C.POOL_AVERAGE, C.DT_QS8, n, 256, h, w, (h, w), (h, w), 0)
    ref["feat"] = N.pooling("oracle", gd, ref["sum"])
    net.typecvt("feat_f32", "feat", M.dtype.Float32())
    ref["feat_f32"] = N.typecvt("oracle", C.DT_QS8, 2.1, C.DT_F32, 0.0, ref["feat"])

    def check(tag):
        torch.cuda.synchronize()
        for name in ("stem", "pool", "b0.c3", "b1.c3", "sum", "feat", "feat_f32"):
            got = net.tensors[name].tensor.reshape(-1).cpu().numpy()
            bad = int((got.view(np.uint8) != ref[name].view(np.uint8)).sum())
            assert bad == 0, (tag, name, bad)

    # every conv node carries a profiled algorithm (fastrun) and the results are still bit-exact
    convs = [nd for nd in net.nodes if nd.kind == "conv_bias"]
    assert len(convs) == 8 and all(nd.algo is not None for nd in convs)
    assert net.chooser.profiled >= 5            # distinct (param, layouts) keys were profiled, repeats hit the cache
    net.exec()
    check("eager")
    # one CUDA graph for the whole network, replayed on fresh contents of the intermediate tensors
    net.capture()
    for name in ("stem", "b0.c3", "feat"):
        net.tensors[name].tensor.zero_()
    net.exec()
    check("cuda graph")
    # profile() ranks every available tile variant of a layer; the winner is what setup_algo pinned
    nd = convs[3]
    T = net.tensors
    dt = T[nd.output].layout.dtype
    ranked = net.chooser.profile(nd.opr, T[nd.inputs[0]], T[nd.inputs[1]], net._get(nd.inputs[2], dt),
                                 net._get(nd.inputs[3], dt), T[nd.output])
    assert len(ranked) >= 2 and all(ms > 0 for _a, ms in ranked)
    check("after profiling")

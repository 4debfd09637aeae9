"""Host-side mirror of megdnn's ConvBiasForward operator interface for the B200 path.

Names, argument meaning and error behaviour follow the reference so that the parity tests read
like dnn/test/cuda/conv_bias*.cpp:

  Handle                      dnn/include/megdnn/handle.h:15-176 (stream + device, no cudnn/cublas)
  DType / TensorLayout        dnn/include/megdnn/dtype.h, basic_types.h:129-300
  Param                       param::ConvBias, dnn/scripts/opr_param_defs.py:199-215
  ConvBiasForward             dnn/include/megdnn/oprs/nn.h:358-536
    .deduce_dtype/.deduce_layout   dnn/src/common/convolution.cpp:545-631,633-1049, conv_bias.cpp:142-147
    .check_exec                    dnn/src/common/conv_bias.cpp:8-133
    .get_workspace_in_bytes/.exec  dnn/src/cuda/conv_bias/opr_impl.cpp:16-28,244-259
    .get_all_algorithms_info / .get_algorithm_info_heuristic / .execution_policy
                                   dnn/include/megdnn/oprs/base.h:545-615

Device memory and streams come from torch (plumbing); all compute goes through the C-ABI
(include/b200cb.h) into hand-written sm_100a kernels. There is no fallback path.
"""
import ctypes
import threading
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import _capi as C


class MegDNNError(RuntimeError):
    """megdnn_assert / megdnn_throw equivalent (dnn/include/megdnn/basic_types.h:21-58)."""


def megdnn_assert(cond, msg=""):
    if not cond:
        raise MegDNNError("assertion failed: " + msg)


# ---------------------------------------------------------------------------------- dtypes
@dataclass(frozen=True)
class DType:
    enumv: int
    scale: float = 0.0

    @property
    def size(self):
        return C.DT_SIZE[self.enumv]

    @property
    def name(self):
        return C.DT_NAME[self.enumv]

    @property
    def is_quantized(self):
        return self.enumv in (C.DT_QS8, C.DT_QS32)

    @property
    def torch_dtype(self):
        return {C.DT_F32: torch.float32, C.DT_F16: torch.float16, C.DT_BF16: torch.bfloat16,
                C.DT_QS8: torch.int8, C.DT_QS32: torch.int32}[self.enumv]


class dtype:
    """megdnn::dtype factory namespace."""
    @staticmethod
    def Float32():
        return DType(C.DT_F32)

    @staticmethod
    def Float16():
        return DType(C.DT_F16)

    @staticmethod
    def BFloat16():
        return DType(C.DT_BF16)

    @staticmethod
    def QuantizedS8(scale):
        return DType(C.DT_QS8, _f32(scale))

    @staticmethod
    def QuantizedS32(scale):
        return DType(C.DT_QS32, _f32(scale))


def _f32(x):
    """round a python float to fp32 (scales are `float` in megdnn)."""
    return ctypes.c_float(x).value


def mul_scale(a: DType, b: DType) -> float:
    """fp32 product of two scales (dnn/src/common/utils.cpp:218-236)."""
    return _f32(ctypes.c_float(a.scale).value * ctypes.c_float(b.scale).value)


@dataclass
class TensorLayout:
    shape: Tuple[int, ...] = ()
    dtype: Optional[DType] = None
    stride: Optional[Tuple[int, ...]] = None  # elements; None = contiguous

    @property
    def ndim(self):
        return len(self.shape)

    def contiguous_stride(self):
        s, acc = [], 1
        for d in reversed(self.shape):
            s.append(acc)
            acc *= d
        return tuple(reversed(s))

    def get_stride(self):
        return self.stride if self.stride is not None else self.contiguous_stride()

    def is_contiguous(self):
        return self.stride is None or tuple(self.stride) == self.contiguous_stride()

    def total_nr_elems(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    def eq_shape(self, other):
        return tuple(self.shape) == tuple(other.shape)


@dataclass
class TensorND:
    """raw device pointer + layout (dnn/include/megdnn/basic_types.h:462-530)."""
    tensor: Optional[torch.Tensor]
    layout: TensorLayout

    @property
    def raw_ptr(self):
        return self.tensor.data_ptr() if self.tensor is not None else None


def make_tensor(t: torch.Tensor, dt: DType) -> TensorND:
    megdnn_assert(t.dtype == dt.torch_dtype, "torch dtype %s does not store %s" % (t.dtype, dt.name))
    stride = tuple(t.stride())
    lay = TensorLayout(tuple(t.shape), dt, None if t.is_contiguous() else stride)
    return TensorND(t, lay)


# ---------------------------------------------------------------------------------- param
@dataclass
class Param:
    """param::ConvBias (opr_param_defs.py:199-215)."""
    nonlineMode: int = C.NL_IDENTITY
    mode: int = C.MODE_CROSS_CORRELATION
    sparse: int = C.SPARSE_DENSE
    format: int = C.FORMAT_NCHW
    pad_h: int = 0
    pad_w: int = 0
    stride_h: int = 1
    stride_w: int = 1
    dilate_h: int = 1
    dilate_w: int = 1
    compute_mode: int = C.COMPUTE_DEFAULT

    class NonlineMode:
        IDENTITY, RELU, SIGMOID, H_SWISH = C.NL_IDENTITY, C.NL_RELU, C.NL_SIGMOID, C.NL_H_SWISH

    class Mode:
        CROSS_CORRELATION, CONVOLUTION = C.MODE_CROSS_CORRELATION, C.MODE_CONVOLUTION

    class Sparse:
        DENSE, GROUP = C.SPARSE_DENSE, C.SPARSE_GROUP

    class Format:
        NCHW, NHWC, NCHW4 = C.FORMAT_NCHW, C.FORMAT_NHWC, C.FORMAT_NCHW4

    class ComputeMode:
        DEFAULT, FLOAT32 = C.COMPUTE_DEFAULT, C.COMPUTE_FLOAT32


@dataclass
class CanonizedFilterMeta:
    """dnn/include/megdnn/oprs/nn.h:82-135 (subset)."""
    group: int
    icpg: int
    ocpg: int
    spatial: Tuple[int, int]


@dataclass
class AlgorithmInfo:
    """Algorithm::Info::Desc (dnn/include/megdnn/oprs/base.h:135-158), name + index."""
    name: str
    algo_id: int
    handle_type: str = "CUDA"


@dataclass
class ExecutionPolicy:
    algo: Optional[AlgorithmInfo] = None


class Handle:
    """megdnn::Handle for a CUDA computing handle = (device, stream); owns neither."""

    def __init__(self, device="cuda:0", stream: Optional[torch.cuda.Stream] = None):
        if not torch.cuda.is_available():
            raise MegDNNError("megengine_b200.Handle needs a CUDA device (no CPU fallback exists)")
        self.device = torch.device(device)
        self._stream = stream
        C.load()

    @property
    def stream(self):
        return self._stream if self._stream is not None else torch.cuda.current_stream(self.device)

    def alignment_requirement(self):
        return 16

    def create_operator(self, cls=None):
        return ConvBiasForward(self)


def infer_conv_shape(inp, flt, stride, pad, dilate=1):
    """dnn/src/common/utils.cpp:130-138."""
    sd = (flt - 1) * dilate + 1
    megdnn_assert(inp + 2 * pad >= sd, "input %d pad %d smaller than dilated filter %d" % (inp, pad, sd))
    return (inp + 2 * pad - sd) // stride + 1


class ConvBiasForward:
    """megdnn::ConvBiasForward on the B200 handle."""

    Param = Param
    BIAS_NONE, BIAS_CHANNEL, BIAS_FULL = C.BIAS_NONE, C.BIAS_CHANNEL, C.BIAS_FULL

    def __init__(self, handle: Optional[Handle]):
        self.handle = handle
        self.param = Param()
        self.execution_policy = ExecutionPolicy()
        self._lib = C.load()

    # ------------------------------------------------------------------ filter meta / deduction
    def make_canonized_filter_meta(self, src_ndim, flt: TensorLayout) -> CanonizedFilterMeta:
        """dnn/src/common/convolution.cpp:478-543 for NCHW / NHWC / NCHW4."""
        p = self.param
        grp = p.sparse == C.SPARSE_GROUP
        s = flt.shape
        if p.format == C.FORMAT_NHWC:
            megdnn_assert(len(s) == (5 if grp else 4), "bad NHWC filter ndim")
            if grp:
                g, ocpg, fh, fw, icpg = s
            else:
                g, (ocpg, fh, fw, icpg) = 1, s
        elif p.format == C.FORMAT_NCHW:
            megdnn_assert(len(s) == (5 if grp else 4), "bad NCHW filter ndim")
            if grp:
                g, ocpg, icpg, fh, fw = s
            else:
                g, (ocpg, icpg, fh, fw) = 1, s
        elif p.format == C.FORMAT_NCHW4:
            megdnn_assert(len(s) == (6 if grp else 5) and s[-1] == 4, "bad NCHW4 filter ndim")
            megdnn_assert(p.dilate_h == 1 and p.dilate_w == 1, "NCHW4 requires dilation 1")
            if grp:
                g, ocpg, icpg4, fh, fw, _ = s
            else:
                g, (ocpg, icpg4, fh, fw, _) = 1, s
            icpg = icpg4 * 4
        else:
            raise MegDNNError("unsupported format %d" % p.format)
        return CanonizedFilterMeta(g, icpg, ocpg, (fh, fw))

    def deduce_dtype(self, src: DType, flt: DType, dst: Optional[DType] = None) -> DType:
        """check_or_deduce_dtype_fwd, dnn/src/common/convolution.cpp:545-631."""
        megdnn_assert(src.enumv == flt.enumv, "src and filter dtype must match")
        if src.enumv in (C.DT_F32, C.DT_F16, C.DT_BF16):
            if dst is not None:
                megdnn_assert(dst.enumv == src.enumv, "float conv: dst dtype must equal src dtype")
            return DType(src.enumv)
        if src.enumv == C.DT_QS8:
            if dst is None:
                return dtype.QuantizedS32(mul_scale(src, flt))
            megdnn_assert(dst.enumv in (C.DT_QS8, C.DT_QS32, C.DT_F32), "bad dst dtype for QuantizedS8")
            return dst
        raise MegDNNError("unsupported src dtype " + src.name)

    def deduce_layout(self, src: TensorLayout, flt: TensorLayout, bias: TensorLayout,
                      z: TensorLayout, dst_dtype: Optional[DType] = None) -> TensorLayout:
        """ConvBiasForward::deduce_layout (conv_bias.cpp:142-147 -> deduce_layout_fwd)."""
        p = self.param
        fm = self.make_canonized_filter_meta(src.ndim, flt)
        out_dt = self.deduce_dtype(src.dtype, flt.dtype, dst_dtype)
        fh, fw = fm.spatial
        if p.format == C.FORMAT_NHWC:
            n, ih, iw, ic = src.shape
        elif p.format == C.FORMAT_NCHW:
            n, ic, ih, iw = src.shape
        else:
            n, ic4, ih, iw, _ = src.shape
            ic = ic4 * 4
        megdnn_assert(fm.icpg * fm.group == ic, "icpg*group (%d*%d) != C_in %d" % (fm.icpg, fm.group, ic))
        oh = infer_conv_shape(ih, fh, p.stride_h, p.pad_h, p.dilate_h)
        ow = infer_conv_shape(iw, fw, p.stride_w, p.pad_w, p.dilate_w)
        oc = fm.ocpg * fm.group
        if p.format == C.FORMAT_NHWC:
            shape = (n, oh, ow, oc)
        elif p.format == C.FORMAT_NCHW:
            shape = (n, oc, oh, ow)
        else:
            megdnn_assert(oc % 4 == 0, "NCHW4 needs OC % 4 == 0")
            shape = (n, oc // 4, oh, ow, 4)
        return TensorLayout(shape, out_dt)

    # ------------------------------------------------------------------ checks
    def _bias_mode(self, bias: TensorLayout, dst: TensorLayout):
        """BiasMode deduction + shape asserts of do_check_exec_common (conv_bias.cpp:38-103)."""
        p = self.param
        if bias.ndim == 0:
            return C.BIAS_NONE
        if bias.eq_shape(dst):
            return C.BIAS_FULL
        if p.format == C.FORMAT_NHWC:
            ok = bias.ndim == 4 and bias.shape[:3] == (1, 1, 1) and bias.shape[3] == dst.shape[3]
        elif p.format == C.FORMAT_NCHW:
            ok = bias.ndim == 4 and bias.shape[0] == 1 and bias.shape[2:] == (1, 1) and bias.shape[1] == dst.shape[1]
        else:
            ok = (bias.ndim == 5 and bias.shape[0] == 1 and bias.shape[2:4] == (1, 1) and
                  bias.shape[4] == 4 and bias.shape[1] == dst.shape[1])
        megdnn_assert(ok, "bias shape %s is neither dst-shaped nor per-channel" % (bias.shape,))
        return C.BIAS_CHANNEL

    def check_exec(self, src, flt, bias, z, dst, workspace_in_bytes=None):
        """ConvBiasForward::check_exec_allow_noncontiguous (conv_bias.cpp:8-133,159-170)."""
        megdnn_assert(src.dtype.enumv == flt.dtype.enumv, "src/filter dtype mismatch")
        if src.dtype.enumv == C.DT_QS8:
            if bias.ndim != 0 or bias.dtype is not None:
                bd = bias.dtype
                megdnn_assert(bd is not None and bd.enumv in (C.DT_QS32, C.DT_F32),
                              "quantized conv needs QuantizedS32 or Float32 bias dtype")
                if bd.enumv == C.DT_QS32:
                    expect = mul_scale(src.dtype, flt.dtype)
                    megdnn_assert(abs(bd.scale - expect) < 1e-6,
                                  "bias scale %g != src*filter scale %g" % (bd.scale, expect))
        elif bias.ndim != 0:
            megdnn_assert(bias.dtype.enumv == src.dtype.enumv, "float conv: bias dtype must equal src dtype")
        expect = self.deduce_layout(src, flt, bias, z, dst.dtype)
        megdnn_assert(expect.eq_shape(dst), "dst shape %s, expected %s" % (dst.shape, expect.shape))
        megdnn_assert(bias.ndim == 0 or bias.is_contiguous(), "bias must be contiguous")
        bias_mode = self._bias_mode(bias, dst)
        if z.ndim != 0:
            megdnn_assert(z.dtype.enumv == dst.dtype.enumv, "z dtype must equal dst dtype")
            megdnn_assert(z.eq_shape(dst), "z shape must equal dst shape")
        if self.param.nonlineMode == C.NL_SIGMOID:
            megdnn_assert(not src.dtype.is_quantized, "SIGMOID is not allowed on quantized ConvBias")
        return bias_mode

    # ------------------------------------------------------------------ C-ABI descriptor
    def _batch_stride(self, lay: TensorLayout):
        if lay.ndim == 0 or lay.is_contiguous():
            return 0
        st, cs = lay.get_stride(), lay.contiguous_stride()
        megdnn_assert(tuple(st[1:]) == tuple(cs[1:]), "only the batch dimension may be non-contiguous")
        return int(st[0])

    def make_desc(self, src, flt, bias, z, dst) -> C.Desc:
        p = self.param
        bias_mode = self.check_exec(src, flt, bias, z, dst)
        fm = self.make_canonized_filter_meta(src.ndim, flt)
        d = C.Desc()
        d.format, d.sparse, d.mode = p.format, p.sparse, p.mode
        d.nonline_mode, d.compute_mode = p.nonlineMode, p.compute_mode
        d.bias_mode, d.has_z = bias_mode, 1 if z.ndim else 0
        d.src_dtype, d.filter_dtype, d.dst_dtype = src.dtype.enumv, flt.dtype.enumv, dst.dtype.enumv
        d.bias_dtype = bias.dtype.enumv if bias.dtype is not None else C.DT_NONE
        d.z_dtype = z.dtype.enumv if (z.ndim and z.dtype is not None) else C.DT_NONE
        d.src_scale, d.filter_scale, d.dst_scale = src.dtype.scale, flt.dtype.scale, dst.dtype.scale
        d.bias_scale = bias.dtype.scale if bias.dtype is not None else 0.0
        d.z_scale = z.dtype.scale if (z.ndim and z.dtype is not None) else 0.0
        if p.format == C.FORMAT_NHWC:
            d.n, d.ih, d.iw, d.ic = src.shape
            _, d.oh, d.ow, d.oc = dst.shape
        elif p.format == C.FORMAT_NCHW:
            d.n, d.ic, d.ih, d.iw = src.shape
            _, d.oc, d.oh, d.ow = dst.shape
        else:
            d.n, ic4, d.ih, d.iw, _ = src.shape
            d.ic = ic4 * 4
            _, oc4, d.oh, d.ow, _ = dst.shape
            d.oc = oc4 * 4
        d.fh, d.fw = fm.spatial
        d.group = fm.group
        d.pad_h, d.pad_w, d.stride_h, d.stride_w = p.pad_h, p.pad_w, p.stride_h, p.stride_w
        d.dilate_h, d.dilate_w = p.dilate_h, p.dilate_w
        d.src_batch_stride = self._batch_stride(src)
        d.z_batch_stride = self._batch_stride(z)
        d.dst_batch_stride = self._batch_stride(dst)
        return d

    # ------------------------------------------------------------------ algorithms
    def _plan(self, desc: C.Desc) -> C.Plan:
        plan = C.Plan()
        algo = self.execution_policy.algo
        if algo is not None:
            rc = self._lib.b200cb_query_algo(ctypes.byref(desc), algo.algo_id, ctypes.byref(plan))
        else:
            rc = self._lib.b200cb_query(ctypes.byref(desc), ctypes.byref(plan))
        if rc != 0:
            # AlgoChooser: "no usable conv_bias fwd algorithm" (dnn/src/common/algo_chooser.h:107-154)
            raise MegDNNError("no usable B200 ConvBias algorithm: " + C.strerror(rc)
                              + (" [%s]" % self._lib.b200cb_last_error_detail().decode() if rc < 0 else ""))
        return plan

    def get_all_algorithms_info(self, src, flt, bias, z, dst) -> List[AlgorithmInfo]:
        """every registered algo whose is_available() holds (oprs/base.h:545-615)."""
        desc = self.make_desc(src, flt, bias, z, dst)
        out, plan = [], C.Plan()
        for a in range(self._lib.b200cb_algo_count()):
            if self._lib.b200cb_query_algo(ctypes.byref(desc), a, ctypes.byref(plan)) == 0:
                out.append(AlgorithmInfo(self._lib.b200cb_algo_name(a).decode(), a))
        return out

    def get_algorithm_info_heuristic(self, src, flt, bias, z, dst) -> AlgorithmInfo:
        desc = self.make_desc(src, flt, bias, z, dst)
        saved, self.execution_policy = self.execution_policy, ExecutionPolicy()
        try:
            plan = self._plan(desc)
        finally:
            self.execution_policy = saved
        return AlgorithmInfo(self._lib.b200cb_algo_name(plan.algo).decode(), plan.algo)

    @staticmethod
    def get_algorithm_set_name():
        return "CONV_BIAS_B200"

    # ------------------------------------------------------------------ algorithm cache
    # megdnn resolves an algorithm in three steps: execution_policy().algo, the process-wide AlgorithmCache
    # keyed by the serialised layouts + param (dnn/include/megdnn/algorithm_cache.h:12-80,
    # dnn/src/common/algo_chooser.h:24-50), then the heuristic.  Same here: steady-state exec() /
    # get_workspace_in_bytes() calls cost one dictionary lookup instead of check_exec + descriptor + query.
    _algorithm_cache = {}
    _algorithm_cache_lock = threading.Lock()
    cache_hits = 0
    cache_misses = 0

    @staticmethod
    def _layout_key(lay: TensorLayout):
        return (tuple(lay.shape), lay.dtype, None if lay.stride is None else tuple(lay.stride))

    def _cached_plan(self, src, flt, bias, z, dst) -> C.Plan:
        p = self.param
        algo = self.execution_policy.algo
        key = (p.nonlineMode, p.mode, p.sparse, p.format, p.pad_h, p.pad_w, p.stride_h, p.stride_w, p.dilate_h,
               p.dilate_w, p.compute_mode, -1 if algo is None else algo.algo_id,
               self._layout_key(src), self._layout_key(flt), self._layout_key(bias), self._layout_key(z),
               self._layout_key(dst))
        cls = ConvBiasForward
        plan = cls._algorithm_cache.get(key)
        if plan is not None:
            cls.cache_hits += 1
            return plan
        plan = self._plan(self.make_desc(src, flt, bias, z, dst))
        with cls._algorithm_cache_lock:
            if len(cls._algorithm_cache) >= 4096:
                cls._algorithm_cache.clear()
            cls._algorithm_cache[key] = plan
            cls.cache_misses += 1
        return plan

    def get_workspace_in_bytes(self, src, flt, bias, z, dst, preprocessed_filter=None) -> int:
        return int(self._cached_plan(src, flt, bias, z, dst).workspace_bytes)

    def get_plan(self, src, flt, bias, z, dst) -> C.Plan:
        return self._cached_plan(src, flt, bias, z, dst)

    # ------------------------------------------------------------------ exec
    def exec(self, src: TensorND, flt: TensorND, bias: TensorND, z: TensorND, dst: TensorND,
             preprocessed_filter=None, workspace: Optional[torch.Tensor] = None, plan: Optional[C.Plan] = None):
        """Enqueue on the handle's stream and return (no synchronisation), like
        cuda::ConvBiasForwardImpl::exec (dnn/src/cuda/conv_bias/opr_impl.cpp:16-28)."""
        if plan is None:
            plan = self._cached_plan(src.layout, flt.layout, bias.layout, z.layout, dst.layout)
        ws_bytes = int(plan.workspace_bytes)
        if ws_bytes:
            megdnn_assert(workspace is not None and workspace.numel() * workspace.element_size() >= ws_bytes,
                          "workspace too small: need %d bytes" % ws_bytes)
        dev = self.handle.device
        megdnn_assert(src.tensor.device == dev and dst.tensor.device == dev,
                      "tensors must live on the handle's device %s" % dev)
        if torch.cuda.current_device() != (dev.index or 0):
            # the C side keys function attributes / SM count on the current device
            with torch.cuda.device(dev):
                return self.exec(src, flt, bias, z, dst, preprocessed_filter, workspace, plan)
        stream = self.handle.stream
        rc = self._lib.b200cb_launch(
            ctypes.byref(plan), src.raw_ptr, flt.raw_ptr,
            bias.raw_ptr if bias.layout.ndim else None, z.raw_ptr if z.layout.ndim else None,
            dst.raw_ptr, workspace.data_ptr() if workspace is not None else None,
            ws_bytes if workspace is not None else 0, stream.cuda_stream)
        if rc != 0:
            raise MegDNNError("b200cb_launch failed: %s (%s)" % (
                C.strerror(rc), self._lib.b200cb_last_error_detail().decode()))
        return plan


EMPTY = TensorND(None, TensorLayout())


def empty_like_dtype(dt: Optional[DType]) -> TensorND:
    """bias/z 'absent' = ndim 0 with the dtype still set (src/opr/impl/dnn/convolution.cpp:838-851)."""
    return TensorND(None, TensorLayout((), dt))

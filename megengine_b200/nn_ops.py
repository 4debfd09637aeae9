"""Host-side mirrors of the megdnn operators either side of ConvBiasForward on the
B200 handle: same names, argument order, deduction rules and error behaviour as the reference, all compute
through the C-ABI of include/b200cb_nn.h (no fallback).

  PoolingForward             dnn/include/megdnn/oprs/nn.h (PoolingBase / PoolingForward), deduce_layout_fwd
                             dnn/src/common/pooling.cpp
  ElemwiseForward            dnn/include/megdnn/oprs/general.h (Elemwise), modes param::Elemwise::Mode
  ElemwiseMultiType          dnn/include/megdnn/oprs/nn_int.h (quantized modes QADD / QFUSE_ADD_RELU / QRELU / QH_SWISH)
  TypeCvt                    dnn/include/megdnn/oprs/general.h (TypeCvt)
  ConvolutionForward         dnn/include/megdnn/oprs/nn.h:196-283   (deduce_dtype dnn/src/common/convolution.cpp:545-631)
  ConvolutionBackwardData    dnn/include/megdnn/oprs/nn.h:285-320   (deduce_layout dnn/src/common/convolution.cpp:1112-1170)
  BatchConvBiasForward       dnn/include/megdnn/oprs/nn.h:1878-1908
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _capi as C
from .conv_bias import (ConvBiasForward, DType, Handle, MegDNNError, Param, TensorLayout, TensorND, dtype,
                        empty_like_dtype, infer_conv_shape, megdnn_assert, mul_scale)


def _check(rc, lib, what):
    if rc != 0:
        raise MegDNNError("%s: %s (%s)" % (what, C.strerror(rc), lib.b200cb_last_error_detail().decode()))


class _Opr:
    def __init__(self, handle: Optional[Handle]):
        self.handle = handle
        self._lib = C.load()

    def _stream(self):
        return self.handle.stream.cuda_stream


# ------------------------------------------------------------------------------------------ pooling
@dataclass
class PoolingParam:
    """param::Pooling (opr_param_defs.py:249-254)."""
    mode: int = C.POOL_MAX
    pad_h: int = 0
    pad_w: int = 0
    stride_h: int = 2
    stride_w: int = 2
    window_h: int = 2
    window_w: int = 2
    format: int = C.FORMAT_NCHW

    class Mode:
        MAX, AVERAGE, AVERAGE_COUNT_EXCLUDE_PADDING = C.POOL_MAX, C.POOL_AVERAGE, C.POOL_AVERAGE_COUNT_EXCLUDE_PADDING


class PoolingForward(_Opr):
    Param = PoolingParam

    def __init__(self, handle=None):
        super().__init__(handle)
        self.param = PoolingParam()

    def _dims(self, lay: TensorLayout):
        f = self.param.format
        if f == C.FORMAT_NHWC:
            megdnn_assert(lay.ndim == 4, "NHWC pooling needs a 4-d tensor")
            n, h, w, c = lay.shape
        elif f == C.FORMAT_NCHW:
            megdnn_assert(lay.ndim == 4, "NCHW pooling needs a 4-d tensor")
            n, c, h, w = lay.shape
        elif f == C.FORMAT_NCHW4:
            megdnn_assert(lay.ndim == 5 and lay.shape[4] == 4, "NCHW4 pooling needs (n, c/4, h, w, 4)")
            n, c4, h, w, _ = lay.shape
            c = c4 * 4
        else:
            raise MegDNNError("unsupported pooling format %d" % f)
        return n, c, h, w

    def deduce_layout(self, src: TensorLayout) -> TensorLayout:
        """PoolingBase::deduce_layout_fwd: infer_conv_shape per spatial axis, dtype carried over."""
        p = self.param
        n, c, h, w = self._dims(src)
        oh = infer_conv_shape(h, p.window_h, p.stride_h, p.pad_h)
        ow = infer_conv_shape(w, p.window_w, p.stride_w, p.pad_w)
        if p.format == C.FORMAT_NHWC:
            shape = (n, oh, ow, c)
        elif p.format == C.FORMAT_NCHW:
            shape = (n, c, oh, ow)
        else:
            shape = (n, c // 4, oh, ow, 4)
        return TensorLayout(shape, src.dtype)

    def get_workspace_in_bytes(self, src, dst):
        return 0

    def make_desc(self, src: TensorLayout, dst: TensorLayout) -> C.PoolDesc:
        p = self.param
        megdnn_assert(src.dtype == dst.dtype, "pooling keeps the dtype (and scale)")
        megdnn_assert(src.is_contiguous() and dst.is_contiguous(), "pooling needs contiguous tensors")
        want = self.deduce_layout(src)
        megdnn_assert(want.eq_shape(dst), "dst shape %s, deduced %s" % (dst.shape, want.shape))
        n, c, h, w = self._dims(src)
        _, _, oh, ow = self._dims(dst)
        d = C.PoolDesc()
        d.format, d.mode, d.dtype = p.format, p.mode, src.dtype.enumv
        d.n, d.c, d.ih, d.iw, d.oh, d.ow = n, c, h, w, oh, ow
        d.window_h, d.window_w, d.stride_h, d.stride_w, d.pad_h, d.pad_w = (
            p.window_h, p.window_w, p.stride_h, p.stride_w, p.pad_h, p.pad_w)
        return d

    def exec(self, src: TensorND, dst: TensorND, workspace=None):
        d = self.make_desc(src.layout, dst.layout)
        _check(self._lib.b200cb_pooling(ctypes.byref(d), src.raw_ptr, dst.raw_ptr, self._stream()), self._lib,
               "b200cb_pooling")


# ------------------------------------------------------------------------------------------ elementwise
class ElemwiseForward(_Opr):
    class Mode:
        RELU, ADD, MAX, MUL, FUSE_ADD_RELU, H_SWISH = (C.ELEM_RELU, C.ELEM_ADD, C.ELEM_MAX, C.ELEM_MUL,
                                                        C.ELEM_FUSE_ADD_RELU, C.ELEM_H_SWISH)

    ARITY = {C.ELEM_RELU: 1, C.ELEM_H_SWISH: 1, C.ELEM_ADD: 2, C.ELEM_MAX: 2, C.ELEM_MUL: 2, C.ELEM_FUSE_ADD_RELU: 2}

    def __init__(self, handle=None):
        super().__init__(handle)
        self.mode = C.ELEM_ADD

    def exec(self, srcs, dst: TensorND):
        megdnn_assert(self.mode in self.ARITY, "elemwise mode %d not implemented on the B200 handle" % self.mode)
        megdnn_assert(len(srcs) == self.ARITY[self.mode], "mode %d takes %d operands" % (self.mode, self.ARITY[self.mode]))
        for s in srcs:
            # (broadcasting operands are the reference's job to expand; the conv graph's residual add is same-shape)
            megdnn_assert(s.layout.eq_shape(dst.layout) and s.layout.dtype == dst.layout.dtype and s.layout.is_contiguous(),
                          "operands must match dst in shape and dtype and be contiguous")
        megdnn_assert(dst.layout.is_contiguous(), "dst must be contiguous")
        b = srcs[1].raw_ptr if len(srcs) > 1 else None
        _check(self._lib.b200cb_elemwise(self.mode, dst.layout.dtype.enumv, dst.layout.total_nr_elems(), srcs[0].raw_ptr,
                                         b, dst.raw_ptr, self._stream()), self._lib, "b200cb_elemwise")


class ElemwiseMultiType(_Opr):
    """quantized modes of ElemwiseMultiType on QuantizedS8 operands."""

    class Mode:
        QADD, QFUSE_ADD_RELU, QRELU, QH_SWISH = 6, 7, 15, 51   # param::ElemwiseMultiType::Mode values

    _TO_ELEM = {6: C.ELEM_ADD, 7: C.ELEM_FUSE_ADD_RELU, 15: C.ELEM_RELU, 51: C.ELEM_H_SWISH}

    def __init__(self, handle=None):
        super().__init__(handle)
        self.mode = self.Mode.QADD

    def exec(self, srcs, dst: TensorND):
        megdnn_assert(self.mode in self._TO_ELEM, "ElemwiseMultiType mode %d not implemented" % self.mode)
        em = self._TO_ELEM[self.mode]
        megdnn_assert(len(srcs) == ElemwiseForward.ARITY[em], "wrong operand count")
        for s in srcs + [dst]:
            megdnn_assert(s.layout.dtype.enumv == C.DT_QS8 and s.layout.is_contiguous(), "QuantizedS8 contiguous tensors")
            megdnn_assert(s.layout.eq_shape(dst.layout), "same-shape operands")
        b = srcs[1] if len(srcs) > 1 else None
        _check(self._lib.b200cb_elemwise_q(
            em, dst.layout.total_nr_elems(), srcs[0].raw_ptr, srcs[0].layout.dtype.scale,
            b.raw_ptr if b else None, b.layout.dtype.scale if b else 0.0, dst.raw_ptr, dst.layout.dtype.scale,
            self._stream()), self._lib, "b200cb_elemwise_q")


class TypeCvt(_Opr):
    def exec(self, src: TensorND, dst: TensorND):
        megdnn_assert(src.layout.eq_shape(dst.layout), "TypeCvt keeps the shape")
        megdnn_assert(src.layout.is_contiguous() and dst.layout.is_contiguous(), "contiguous tensors")
        s, d = src.layout.dtype, dst.layout.dtype
        _check(self._lib.b200cb_typecvt(s.enumv, s.scale, d.enumv, d.scale, src.layout.total_nr_elems(), src.raw_ptr,
                                        dst.raw_ptr, self._stream()), self._lib, "b200cb_typecvt")


# ------------------------------------------------------------------------------------------ convolution front-ends
class ConvolutionForward(_Opr):
    """ConvolutionForward = ConvBias without bias, z and activation: delegates to the ConvBias mirror the way the
    reference's CUDA ConvolutionForwardImpl builds a ConvBias operator (dnn/src/cuda/convolution/opr_impl.cpp)."""
    Param = Param

    def __init__(self, handle=None):
        super().__init__(handle)
        self.param = Param()
        self._cb = ConvBiasForward(handle)

    def _sync_param(self):
        p, q = self.param, self._cb.param
        for k in ("mode", "sparse", "format", "pad_h", "pad_w", "stride_h", "stride_w", "dilate_h", "dilate_w",
                  "compute_mode"):
            setattr(q, k, getattr(p, k))
        q.nonlineMode = C.NL_IDENTITY

    def deduce_dtype(self, src: DType, flt: DType) -> DType:
        """ConvolutionBase::deduce_dtype: QuantizedS8 x QuantizedS8 -> QuantizedS32(scale product), floats keep
        their dtype (dnn/src/common/convolution.cpp:545-631)."""
        if src.enumv == C.DT_QS8:
            return dtype.QuantizedS32(mul_scale(src, flt))
        return src

    def deduce_layout(self, src: TensorLayout, flt: TensorLayout) -> TensorLayout:
        self._sync_param()
        return self._cb.deduce_layout(src, flt, TensorLayout((), None), TensorLayout((), None),
                                      self.deduce_dtype(src.dtype, flt.dtype))

    def _absent(self, dst: TensorLayout):
        return empty_like_dtype(dst.dtype), empty_like_dtype(dst.dtype)

    def get_workspace_in_bytes(self, src, flt, dst, preprocessed_filter=None):
        self._sync_param()
        b, z = self._absent(dst)
        return self._cb.get_workspace_in_bytes(src, flt, b.layout, z.layout, dst)

    def exec(self, src: TensorND, flt: TensorND, dst: TensorND, preprocessed_filter=None, workspace=None):
        self._sync_param()
        b, z = self._absent(dst.layout)
        return self._cb.exec(src, flt, b, z, dst, preprocessed_filter, workspace)


class ConvolutionBackwardData(_Opr):
    """grad = d(conv(src, filter)) / d(src) . diff -- NHWC dense on the B200 handle."""
    Param = Param

    def __init__(self, handle=None):
        super().__init__(handle)
        self.param = Param()

    def deduce_dtype(self, flt: DType, diff: DType) -> DType:
        """dnn/src/common/convolution.cpp (ConvolutionBackwardData::deduce_dtype): QuantizedS8 pairs give
        QuantizedS8 only when the caller names the scale, so quantized grads must be supplied; floats keep dtype."""
        megdnn_assert(diff.enumv != C.DT_QS8, "quantized grad dtype carries its own scale: supply the grad layout")
        return diff

    def deduce_layout(self, flt: TensorLayout, diff: TensorLayout) -> TensorLayout:
        """smallest input extent that produces diff: (o - 1) * stride + dilated filter - 2 * pad."""
        p = self.param
        megdnn_assert(p.format == C.FORMAT_NHWC and p.sparse == C.SPARSE_DENSE, "NHWC dense only")
        n, oh, ow, oc = diff.shape
        oc2, fh, fw, ic = flt.shape
        megdnn_assert(oc == oc2, "diff / filter channel mismatch")
        ih = (oh - 1) * p.stride_h + (fh - 1) * p.dilate_h + 1 - 2 * p.pad_h
        iw = (ow - 1) * p.stride_w + (fw - 1) * p.dilate_w + 1 - 2 * p.pad_w
        return TensorLayout((n, ih, iw, ic), self.deduce_dtype(flt.dtype, diff.dtype))

    def make_desc(self, flt: TensorLayout, diff: TensorLayout, grad: TensorLayout) -> C.Desc:
        p = self.param
        megdnn_assert(p.format == C.FORMAT_NHWC and p.sparse == C.SPARSE_DENSE, "NHWC dense only")
        for l in (flt, diff, grad):
            megdnn_assert(l.ndim == 4 and l.is_contiguous(), "4-d contiguous tensors")
        d = C.Desc()
        d.format, d.sparse, d.mode, d.compute_mode = p.format, p.sparse, p.mode, p.compute_mode
        d.nonline_mode, d.bias_mode, d.has_z = C.NL_IDENTITY, C.BIAS_NONE, 0
        d.src_dtype, d.filter_dtype, d.dst_dtype = grad.dtype.enumv, flt.dtype.enumv, diff.dtype.enumv
        d.bias_dtype = d.z_dtype = C.DT_NONE
        d.src_scale, d.filter_scale, d.dst_scale = grad.dtype.scale, flt.dtype.scale, diff.dtype.scale
        d.n, d.ih, d.iw, d.ic = grad.shape
        n2, d.oh, d.ow, d.oc = diff.shape
        oc2, d.fh, d.fw, ic2 = flt.shape
        megdnn_assert(n2 == d.n and oc2 == d.oc and ic2 == d.ic, "filter / diff / grad shapes disagree")
        d.group = 1
        d.pad_h, d.pad_w, d.stride_h, d.stride_w = p.pad_h, p.pad_w, p.stride_h, p.stride_w
        d.dilate_h, d.dilate_w = p.dilate_h, p.dilate_w
        megdnn_assert(infer_conv_shape(d.ih, d.fh, d.stride_h, d.pad_h, d.dilate_h) == d.oh and
                      infer_conv_shape(d.iw, d.fw, d.stride_w, d.pad_w, d.dilate_w) == d.ow,
                      "diff extent is not what the forward convolution of grad's extent produces")
        return d

    def _plan(self, flt, diff, grad) -> C.Plan:
        plan = C.Plan()
        d = self.make_desc(flt, diff, grad)
        rc = self._lib.b200cb_deconv_query(ctypes.byref(d), ctypes.byref(plan))
        if rc != 0:
            raise MegDNNError("no usable B200 ConvolutionBackwardData algorithm: " + C.strerror(rc))
        return plan

    def get_workspace_in_bytes(self, flt, diff, grad):
        return int(self._plan(flt, diff, grad).workspace_bytes)

    def exec(self, flt: TensorND, diff: TensorND, grad: TensorND, workspace: Optional[torch.Tensor] = None):
        plan = self._plan(flt.layout, diff.layout, grad.layout)
        ws = int(plan.workspace_bytes)
        megdnn_assert(workspace is not None and workspace.numel() * workspace.element_size() >= ws,
                      "workspace too small: need %d bytes" % ws)
        _check(self._lib.b200cb_deconv_launch(ctypes.byref(plan), flt.raw_ptr, diff.raw_ptr, grad.raw_ptr,
                                              workspace.data_ptr(), ws, self._stream()), self._lib, "b200cb_deconv_launch")
        return plan


class BatchConvBiasForward(_Opr):
    """ConvBias with one filter per image: filter layout = (n,) + the ConvBias filter layout."""
    Param = Param

    def __init__(self, handle=None):
        super().__init__(handle)
        self.param = Param()
        self._cb = ConvBiasForward(handle)

    def _desc(self, src, flt, bias, z, dst) -> C.Desc:
        self._cb.param = self.param
        megdnn_assert(flt.ndim >= 5 and flt.shape[0] == src.shape[0], "filter needs a leading batch dimension")
        one = TensorLayout(tuple(flt.shape[1:]), flt.dtype)
        return self._cb.make_desc(src, one, bias, z, dst)

    def _plan(self, src, flt, bias, z, dst) -> C.Plan:
        plan = C.Plan()
        d = self._desc(src, flt, bias, z, dst)
        rc = self._lib.b200cb_batch_convbias_query(ctypes.byref(d), ctypes.byref(plan))
        if rc != 0:
            raise MegDNNError("no usable B200 BatchConvBias algorithm: " + C.strerror(rc))
        return plan

    def get_workspace_in_bytes(self, src, flt, bias, z, dst):
        return int(self._plan(src, flt, bias, z, dst).workspace_bytes)

    def exec(self, src: TensorND, flt: TensorND, bias: TensorND, z: TensorND, dst: TensorND,
             workspace: Optional[torch.Tensor] = None):
        plan = self._plan(src.layout, flt.layout, bias.layout, z.layout, dst.layout)
        ws = int(plan.workspace_bytes)
        if ws:
            megdnn_assert(workspace is not None and workspace.numel() * workspace.element_size() >= ws,
                          "workspace too small: need %d bytes" % ws)
        _check(self._lib.b200cb_batch_convbias_launch(
            ctypes.byref(plan), src.raw_ptr, flt.raw_ptr, bias.raw_ptr if bias.layout.ndim else None,
            z.raw_ptr if z.layout.ndim else None, dst.raw_ptr, workspace.data_ptr() if workspace is not None else None,
            ws, self._stream()), self._lib, "b200cb_batch_convbias_launch")
        return plan

/*
 * b200cb_nn.h -- C ABI of the operators either side of ConvBiasForward:
 * the steps between convolutions of a ResNet / MobileNet graph and the convolution front-ends that delegate
 * to the ConvBias kernels.  Same conventions as b200cb.h: plain pointers and sizes, enqueue on the caller's
 * stream, never synchronise; 0 = OK, > 0 = not supported, < 0 = CUDA failure.
 *
 * Reference interfaces replaced (paths relative to the MegEngine tree):
 *   b200cb_pooling            <- PoolingForward::exec            dnn/include/megdnn/oprs/nn.h (PoolingForward),
 *                                CUDA: dnn/src/cuda/pooling/opr_impl.cpp; semantics dnn/src/naive/pooling/opr_impl.cpp:17-330
 *   b200cb_elemwise           <- ElemwiseForward::exec           dnn/include/megdnn/oprs/general.h (Elemwise),
 *                                kernels dnn/src/common/elemwise/kern_defs.cuh:169-270
 *   b200cb_elemwise_q         <- ElemwiseMultiType::exec, modes QADD / QFUSE_ADD_RELU / QRELU
 *                                dnn/src/naive/elemwise_multi_type/opr_impl.h:45-120, CUDA dnn/src/cuda/elemwise_multi_type/
 *   b200cb_typecvt            <- TypeCvt::exec                   dnn/src/naive/type_cvt/opr_impl.cpp:24-57
 *   b200cb_convolution_*      <- ConvolutionForward              dnn/include/megdnn/oprs/nn.h:196-283; the reference's CUDA
 *                                ConvolutionForwardImpl delegates to ConvBias the same way (dnn/src/cuda/convolution/opr_impl.cpp)
 *   b200cb_deconv_*           <- ConvolutionBackwardData         dnn/include/megdnn/oprs/nn.h:285-320,
 *                                semantics dnn/src/naive/convolution/opr_impl.cpp:92-167, helper.h:669-681
 *   b200cb_batch_convbias_*   <- BatchConvBiasForward            dnn/include/megdnn/oprs/nn.h:1878-1908,
 *                                semantics dnn/src/naive/batch_conv_bias/opr_impl.cpp
 */
#ifndef B200CB_NN_H_
#define B200CB_NN_H_

#include "b200cb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* param::Pooling::Mode (dnn/scripts/opr_param_defs.py:233-254) */
enum { B200CB_POOL_MAX = 0, B200CB_POOL_AVERAGE = 1, B200CB_POOL_AVERAGE_COUNT_EXCLUDE_PADDING = 2 };

typedef struct b200cb_pool_desc {
    uint32_t format; /* B200CB_FORMAT_NHWC / NCHW / NCHW4 (QuantizedS8, c % 4 == 0) */
    uint32_t mode;   /* B200CB_POOL_* */
    uint32_t dtype;  /* B200CB_DT_F32 / F16 / BF16 / QS8 (src and dst share dtype and scale) */
    uint32_t n, c, ih, iw, oh, ow;
    uint32_t window_h, window_w, stride_h, stride_w, pad_h, pad_w;
} b200cb_pool_desc;

/* oh / ow must equal megdnn's deduction (ih + 2 pad - window) / stride + 1; a window that lies entirely in the
 * padding is rejected (the reference throws). */
int b200cb_pooling(const b200cb_pool_desc* desc, const void* src, void* dst, void* stream);

/* param::Elemwise::Mode values (dnn/scripts/opr_param_defs.py:342-411) of the modes implemented */
enum {
    B200CB_ELEM_RELU = 0,
    B200CB_ELEM_ADD = 16,
    B200CB_ELEM_MAX = 18,
    B200CB_ELEM_MUL = 21,
    B200CB_ELEM_FUSE_ADD_RELU = 37,
    B200CB_ELEM_H_SWISH = 49
};
/* same-shape contiguous operands; dtype F32 / F16 / BF16 (half types compute in the half type, as the
 * reference's ctype does); `b` is ignored by the unary modes */
int b200cb_elemwise(
        uint32_t mode, uint32_t dtype, uint64_t nelems, const void* a, const void* b, void* dst, void* stream);

/* ElemwiseMultiType quantized modes on QuantizedS8: dst = quantize_dst(op(a * scale_a, b * scale_b));
 * mode = B200CB_ELEM_ADD (QADD), B200CB_ELEM_FUSE_ADD_RELU (QFUSE_ADD_RELU), B200CB_ELEM_RELU (QRELU),
 * B200CB_ELEM_H_SWISH (QH_SWISH) */
int b200cb_elemwise_q(
        uint32_t mode, uint64_t nelems, const void* a, float scale_a, const void* b, float scale_b, void* dst,
        float scale_dst, void* stream);

/* TypeCvt between F32 / F16 / BF16 / QS8 / QS32 (scales used by the quantized sides only) */
int b200cb_typecvt(
        uint32_t src_dtype, float src_scale, uint32_t dst_dtype, float dst_scale, uint64_t nelems, const void* src,
        void* dst, void* stream);

/* ConvolutionForward: `desc` as for ConvBias; bias_mode / has_z / nonline_mode are ignored (forced to
 * NONE / 0 / IDENTITY).  The plan is an ordinary ConvBias plan. */
int b200cb_convolution_query(const b200cb_desc* desc, b200cb_plan* plan);
int b200cb_convolution_launch(
        const b200cb_plan* plan, const void* src, const void* filter, void* dst, void* workspace,
        size_t workspace_bytes, void* stream);

/* ConvolutionBackwardData (deconvolution).  `fwd` describes the FORWARD convolution whose data gradient is
 * computed: (n, ic, ih, iw) is the shape of grad, (n, oc, oh, ow) the shape of diff, src_dtype = dtype of grad,
 * dst_dtype = dtype of diff, filter (oc, fh, fw, ic).  NHWC, dense.  BF16 / F16 (compute_mode FLOAT32), F32,
 * and QuantizedS8 (diff, filter) -> QuantizedS8 grad.  Runs as: filter transposed + flipped and, for
 * stride > 1, diff zero-upsampled into the workspace, then the stride-1 ConvBias kernel. */
int b200cb_deconv_query(const b200cb_desc* fwd, b200cb_plan* plan);
int b200cb_deconv_launch(
        const b200cb_plan* plan, const void* filter, const void* diff, void* grad, void* workspace,
        size_t workspace_bytes, void* stream);

/* BatchConvBiasForward: ConvBias with one filter per image, filter (n, oc, ...) -- `desc` as for ConvBias.
 * Runs one ConvBias launch per image. */
int b200cb_batch_convbias_query(const b200cb_desc* desc, b200cb_plan* plan);
int b200cb_batch_convbias_launch(
        const b200cb_plan* plan, const void* src, const void* filter, const void* bias, const void* z, void* dst,
        void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CB_NN_H_ */

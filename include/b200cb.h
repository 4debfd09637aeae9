/*
 * b200cb.h -- C ABI of the B200-native (sm_100a) ConvBiasForward launch layer.
 *
 * This is the drop-in boundary: the thin layer that an
 * `AlgoBase` subclass registered in megdnn's CUDA `ConvBiasForwardImpl::AlgoPack`
 * calls from `is_available()`, `get_workspace_in_bytes()` and `exec()`.
 * Plain pointers and sizes only; no exceptions, no torch / megdnn types.
 *
 * Reference interfaces replaced (paths relative to the MegEngine tree):
 *   b200cb_desc          <- ConvBiasForwardImpl::AlgoBase::SizeArgs
 *                           dnn/src/cuda/conv_bias/algo.h:43-82 (+ BiasForwardSizeArgs,
 *                           dnn/src/cuda/conv_bias/helper.h:27-36) and param::ConvBias
 *                           dnn/scripts/opr_param_defs.py:199-215
 *   b200cb_query         <- AlgoBase::is_available + AlgoBase::get_workspace_in_bytes
 *                           dnn/src/cuda/conv_bias/algo.h:83-165
 *   b200cb_launch        <- AlgoBase::exec(const ExecArgs&)
 *                           dnn/src/cuda/conv_bias/algo.h:153-160,
 *                           e.g. implicit_gemm_int8_nhwc_imma.cpp:164-230
 *   b200cb_preprocess_filter / b200cb_preprocessed_filter_bytes
 *                        <- AlgoBase::exec_preprocess / deduce_preprocessed_filter_layout
 *                           dnn/src/cuda/conv_bias/implicit_gemm_int8_nhwc_imma.cpp:84-107
 *   b200cb_algo_count/name/query_algo
 *                        <- AlgoPack::all_algos + get_all_algorithms
 *                           dnn/src/cuda/conv_bias/algo.cpp:7-108, opr_impl.cpp:30-42
 *   b200cb_plan_use_preprocessed_filter
 *                        <- ConvBiasForward::exec(..., preprocessed_filter, ...)
 *                           dnn/include/megdnn/oprs/nn.h:381-391,400-447
 *   b200cb_strerror      <- megdnn_throw / cuda_check message text
 *                           dnn/src/cuda/utils.cuh:16-29
 *
 * Return convention: 0 = OK; > 0 = configuration not supported by this algo
 * (caller maps it to is_available()==false); < 0 = CUDA / driver error
 * (caller maps it to megdnn_throw).
 */
#ifndef B200CB_H_
#define B200CB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* values equal to megdnn param::ConvBias::Format (opr_param_defs.py:96-113) */
enum { B200CB_FORMAT_NCHW = 0, B200CB_FORMAT_NHWC = 1, B200CB_FORMAT_NCHW4 = 3 };
/* param::ConvBias::NonlineMode (opr_param_defs.py:201) */
enum { B200CB_NL_IDENTITY = 0, B200CB_NL_RELU = 1, B200CB_NL_SIGMOID = 2, B200CB_NL_H_SWISH = 3 };
/* param::Convolution::Mode */
enum { B200CB_MODE_CROSS_CORRELATION = 0, B200CB_MODE_CONVOLUTION = 1 };
/* param::Convolution::Sparse */
enum { B200CB_SPARSE_DENSE = 0, B200CB_SPARSE_GROUP = 1 };
/* param::Convolution::ComputeMode */
enum { B200CB_COMPUTE_DEFAULT = 0, B200CB_COMPUTE_FLOAT32 = 1 };
/* ConvBiasForward::BiasMode (dnn/include/megdnn/oprs/nn.h:450-454) */
enum { B200CB_BIAS_NONE = 0, B200CB_BIAS_CHANNEL = 1, B200CB_BIAS_FULL = 2 };
/* the DTypeEnum subset reachable on this path (dnn/include/megdnn/dtype.h) */
enum {
    B200CB_DT_F32 = 0,
    B200CB_DT_F16 = 1,
    B200CB_DT_BF16 = 2,
    B200CB_DT_QS8 = 3,  /* QuantizedS8,  carries a float scale */
    B200CB_DT_QS32 = 4, /* QuantizedS32, carries a float scale */
    B200CB_DT_NONE = 255
};

/* error codes (>0 unsupported, <0 runtime failure) */
enum {
    B200CB_OK = 0,
    B200CB_UNSUPPORTED_FORMAT = 1,
    B200CB_UNSUPPORTED_DTYPE = 2,
    B200CB_UNSUPPORTED_SHAPE = 3,
    B200CB_UNSUPPORTED_PARAM = 4,
    B200CB_UNSUPPORTED_ALIGNMENT = 5,
    B200CB_UNSUPPORTED_ALGO = 6,
    B200CB_ERR_CUDA = -1,
    B200CB_ERR_DRIVER = -2,
    B200CB_ERR_WORKSPACE = -3,
    B200CB_ERR_BAD_PLAN = -4,
    B200CB_ERR_NULL = -5
};

/* Problem descriptor: everything megdnn's SizeArgs carries, as a POD. */
typedef struct b200cb_desc {
    uint32_t format;        /* B200CB_FORMAT_* */
    uint32_t sparse;        /* B200CB_SPARSE_* ; depthwise = GROUP with icpg==ocpg==1 */
    uint32_t mode;          /* B200CB_MODE_* */
    uint32_t nonline_mode;  /* B200CB_NL_* */
    uint32_t compute_mode;  /* B200CB_COMPUTE_* */
    uint32_t bias_mode;     /* B200CB_BIAS_* */
    uint32_t has_z;         /* z.layout.ndim != 0 */
    uint32_t src_dtype, filter_dtype, bias_dtype, z_dtype, dst_dtype; /* B200CB_DT_* */
    float src_scale, filter_scale, bias_scale, z_scale, dst_scale;    /* quantized dtypes only */
    uint32_t n, ic, ih, iw; /* ic = total input channels  (group * icpg) */
    uint32_t oc, oh, ow;    /* oc = total output channels (group * ocpg) */
    uint32_t fh, fw, group;
    uint32_t pad_h, pad_w, stride_h, stride_w, dilate_h, dilate_w;
    /* batch strides in ELEMENTS (TensorLayout::stride[0]); 0 = contiguous */
    uint64_t src_batch_stride, z_batch_stride, dst_batch_stride;
} b200cb_desc;

/* Execution plan produced by b200cb_query*: POD, copyable, valid for the process lifetime. */
typedef struct b200cb_plan {
    uint32_t magic;
    int32_t algo;             /* index into the algo table */
    uint32_t kernel;          /* internal kernel family id */
    uint32_t block_n;         /* OC tile */
    uint32_t stages;          /* smem pipeline depth */
    uint32_t grid;            /* CTAs launched */
    uint32_t smem_bytes;      /* dynamic shared memory per CTA */
    uint32_t a_mode;          /* 0 tiled TMA (1x1 s1), 1 im2col TMA */
    uint32_t reserved[8];
    uint64_t workspace_bytes; /* AlgoBase::get_workspace_in_bytes */
    uint64_t flops;           /* 2*N*OC*OH*OW*(IC/group)*FH*FW  (conv_test_utils.cpp:193) */
    uint64_t bytes;           /* algorithmic HBM bytes */
    b200cb_desc desc;
} b200cb_plan;

/* number of algorithms this library registers (>= 1) */
int b200cb_algo_count(void);
/* unique, stable name of algorithm `algo` (megdnn Algorithm::name()); NULL if out of range */
const char* b200cb_algo_name(int algo);

/* heuristic choice: availability + tile config + workspace size. Pure host code, reentrant,
 * performs no CUDA calls. */
int b200cb_query(const b200cb_desc* desc, b200cb_plan* plan);
/* same, for one specific algorithm (fastrun enumerates these) */
int b200cb_query_algo(const b200cb_desc* desc, int algo, b200cb_plan* plan);

/* enqueue on `stream` (a cudaStream_t) and return; never synchronises. bias / z may be NULL
 * when the descriptor says they are absent. */
int b200cb_launch(
        const b200cb_plan* plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* workspace, size_t workspace_bytes, void* stream);

/* weight pre-processing (inference): bytes of the packed filter (0 = algo needs none) and the
 * one-off packing launch. */
size_t b200cb_preprocessed_filter_bytes(const b200cb_plan* plan);
int b200cb_preprocess_filter(
        const b200cb_plan* plan, const void* filter, void* packed, void* stream);

/* Marks `plan` as running on a filter already packed by b200cb_preprocess_filter (megdnn's
 * exec(..., preprocessed_filter, ...) path, dnn/include/megdnn/oprs/nn.h:400-447): b200cb_launch then
 * takes the packed buffer as `filter` and the plan's workspace_bytes drops to what that case needs. */
int b200cb_plan_use_preprocessed_filter(b200cb_plan* plan);

/* number of kernels b200cb_launch enqueues for this plan (for launch accounting) */
int b200cb_plan_num_launches(const b200cb_plan* plan);

/* Host-buffer convenience used by the end-to-end measurement: copies src/filter/bias/z from
 * host memory, runs the plan, copies dst back and synchronises `stream`.  Device scratch is
 * owned by the caller (`dev_scratch`, at least b200cb_host_scratch_bytes(plan)). */
size_t b200cb_host_scratch_bytes(const b200cb_plan* plan);
int b200cb_exec_host(
        const b200cb_plan* plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* dev_scratch, size_t dev_scratch_bytes, void* stream);

/* First-class layout operators either side of the path: what megdnn's RelayoutFormat
 * (dnn/include/megdnn/oprs/general.h RelayoutFormat, modes param::RelayoutFormat::Mode in
 * dnn/scripts/opr_param_defs.py; CUDA kernels dnn/src/cuda/relayout_format/) and gopt's channel padding
 * (src/gopt/impl/padding_channel.cpp) put around an NHWC / NCHW4 ConvBias.  `c` = channels of src, `c_out` =
 * channels of dst (>= c: zero padded; NCHW4 sides need multiples of 4); dense tensors; enqueue only.
 *   NCHW_NHWC   (n,c,h,w)      -> (n,h,w,c_out)        NHWC_NCHW   (n,h,w,c)     -> (n,c_out,h,w)
 *   NCHW_NCHW4  (n,c,h,w)      -> (n,c_out/4,h,w,4)    NCHW4_NCHW  (n,c/4,h,w,4) -> (n,c_out,h,w)  (c_out <= c: drops padding)
 *   NCHW4_NHWC  (n,c/4,h,w,4)  -> (n,h,w,c_out)        NHWC_NCHW4  (n,h,w,c)     -> (n,c_out/4,h,w,4)
 *   NHWC_PAD_C  (n,h,w,c)      -> (n,h,w,c_out)
 * elem_bytes = 1, 2 or 4. */
enum {
    B200CB_RELAYOUT_NCHW_NHWC = 0,
    B200CB_RELAYOUT_NHWC_NCHW = 1,
    B200CB_RELAYOUT_NCHW_NCHW4 = 2,
    B200CB_RELAYOUT_NCHW4_NCHW = 3,
    B200CB_RELAYOUT_NCHW4_NHWC = 4,
    B200CB_RELAYOUT_NHWC_NCHW4 = 5,
    B200CB_RELAYOUT_NHWC_PAD_C = 6
};
int b200cb_relayout(
        uint32_t mode, uint32_t elem_bytes, uint32_t n, uint32_t c, uint32_t h, uint32_t w, uint32_t c_out,
        const void* src, void* dst, void* stream);

/* Device self-tests (synchronise `stream`). which == 0: the requantize epilogue's reciprocal-based
 * division against IEEE division over all 2^32 dividends; which == 1: the ALU-only quantize of the
 * fast int8 epilogue against roundf(IEEE division) + clamp; which == 2: the packed fp32x2 form of that
 * epilogue against the scalar one over all 2^32 accumulators.  *mismatches must come back 0. */
int b200cb_selftest(int which, void* stream, uint64_t* mismatches);

const char* b200cb_strerror(int code);
/* text of the last CUDA/driver failure on this thread ("" if none) */
const char* b200cb_last_error_detail(void);

#ifdef __cplusplus
}
#endif
#endif /* B200CB_H_ */

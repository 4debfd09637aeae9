// C-ABI entry points (include/b200cb.h): algorithm table, heuristic, dispatch, error text.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"

namespace b200cb {

namespace {
thread_local char g_last_error[256] = "";

struct AlgoEntry {
    const char* name;
    uint32_t kernel;
    uint32_t block_n;
};
// Order is part of the ABI: megdnn serialises Algorithm::Info::Desc{type, param, name} into its
// persistent fastrun cache (dnn/include/megdnn/oprs/base.h:135-158) -- append only.
const AlgoEntry g_algos[] = {
        {"B200_IGEMM_TCGEN05_128x256", KERNEL_IGEMM, 256},
        {"B200_IGEMM_TCGEN05_128x128", KERNEL_IGEMM, 128},
        {"B200_IGEMM_TCGEN05_128x64", KERNEL_IGEMM, 64},
        {"B200_DWCONV_DIRECT", KERNEL_DWCONV, 0},
};
constexpr int NUM_ALGOS = sizeof(g_algos) / sizeof(g_algos[0]);

void fill_accounting(b200cb_plan* plan) {
    const b200cb_desc& d = plan->desc;
    const uint64_t icpg = d.ic / (d.group ? d.group : 1);
    // flops = 2*N*OC*OH*OW*(IC/group)*FH*FW  (dnn/test/cuda/conv_test_utils.cpp:193)
    plan->flops = 2ull * d.n * d.oc * d.oh * d.ow * icpg * d.fh * d.fw;
    // algorithmic bytes: every tensor touched once; 1x1 strided convs only read
    // the pixels they use
    uint64_t src_pix = (uint64_t)d.ih * d.iw;
    if (d.fh == 1 && d.fw == 1 && d.pad_h == 0 && d.pad_w == 0)
        src_pix = (uint64_t)d.oh * d.ow;
    uint64_t b = (uint64_t)d.n * src_pix * d.ic * dtype_size(d.src_dtype);
    b += (uint64_t)d.oc * icpg * d.fh * d.fw * dtype_size(d.filter_dtype);
    if (d.bias_mode == B200CB_BIAS_CHANNEL)
        b += (uint64_t)d.oc * dtype_size(d.bias_dtype);
    const uint64_t dst_elems = (uint64_t)d.n * d.oh * d.ow * d.oc;
    if (d.bias_mode == B200CB_BIAS_FULL)
        b += dst_elems * dtype_size(d.bias_dtype);
    if (d.has_z)
        b += dst_elems * dtype_size(d.z_dtype);
    b += dst_elems * dtype_size(d.dst_dtype);
    plan->bytes = b;
}

int validate_desc(const b200cb_desc* d) {
    if (!d)
        return B200CB_ERR_NULL;
    if (d->group == 0 || d->ic % d->group != 0 || d->oc % d->group != 0)
        return B200CB_UNSUPPORTED_SHAPE;
    // output extent must match megdnn's infer_conv_shape (dnn/src/common/utils.cpp:130-138)
    auto out_dim = [](uint32_t i, uint32_t f, uint32_t s, uint32_t p, uint32_t dl) -> int64_t {
        int64_t eff = (int64_t)(f - 1) * dl + 1;
        int64_t num = (int64_t)i + 2 * (int64_t)p - eff;
        if (num < 0 || s == 0)
            return -1;
        return num / s + 1;
    };
    if (d->stride_h == 0 || d->stride_w == 0 || d->dilate_h == 0 || d->dilate_w == 0)
        return B200CB_UNSUPPORTED_PARAM;
    if (out_dim(d->ih, d->fh, d->stride_h, d->pad_h, d->dilate_h) != (int64_t)d->oh ||
        out_dim(d->iw, d->fw, d->stride_w, d->pad_w, d->dilate_w) != (int64_t)d->ow)
        return B200CB_UNSUPPORTED_SHAPE;
    return B200CB_OK;
}

// Heuristic rank of an implicit-GEMM tile width (lower is better), from the measured sweep
// (tools/sweep_resnet.py, profiles/): the widest tile that is not wider than OC wins whenever it still
// yields at least one tile per SM -- operand traffic per flop falls with the tile width and the
// kernels are shared-memory / L2 feed bound -- and below that the tile count is what matters.
double igemm_cost(const b200cb_desc& d, const b200cb_plan& plan, uint32_t bn, int sms) {
    const double M = (double)d.n * d.oh * d.ow;
    const double tiles = ceil(M / IGEMM_BLOCK_M) * ceil((double)d.oc / bn);
    if (tiles < sms)
        return 1.0 + 1.0 / tiles;  // small problems: most tiles first
    double cost = 1.0 / bn;        // enough parallelism: widest first ...
    // (round 2 had a rule here that sent ResNet-50 l3.c3 -- residual z, K = 512 bytes, no room for a two-deep z ring next
    // to a 256-wide tile -- to 128x128 tiles; with the merge-free 16-bit epilogue the 256-wide tile wins again:
    // 0.0550 vs 0.0590 ms, profiles/r02/exp_epilogue_tail_ab.txt)
    // QuantizedS8 residual layers whose 256-wide filter tile cannot stay resident (IgemmParams::b_res: K * BN > 64 KiB)
    // run faster on the narrower tile that can (l4.c3: 128x128 resident 0.026 ms, 128x256 streamed 0.030)
    if (d.has_z && plan.kernel == KERNEL_IGEMM && d.src_dtype == B200CB_DT_QS8 && !(plan.reserved[2] & 8u) && bn > 128 &&
        d.format == B200CB_FORMAT_NHWC && d.group <= 1 && d.fh == 1 && d.fw == 1 && d.stride_h == 1 && d.stride_w == 1 &&
        d.pad_h == 0 && d.pad_w == 0 && d.src_batch_stride == 0 && (uint64_t)d.ic * 128 <= 64 * 1024)
        cost *= 2.5;
    return cost;
}
// ---- NCHW adapter (fp16 / bf16): the tensor-core kernels want channels innermost, so NCHW tensors are
// re-tiled into NHWC / KRSC copies in the workspace (channels zero-padded to what TMA needs), the NHWC
// plan of the same algorithm runs on those, and dst is re-tiled back.  Functional coverage of the
// reference's default layout (imperative F.conv2d); the fast path is NHWC.
bool is_nchw16(const b200cb_desc& d) {
    return d.format == B200CB_FORMAT_NCHW && (d.src_dtype == B200CB_DT_BF16 || d.src_dtype == B200CB_DT_F16);
}
bool nchw_depthwise(const b200cb_desc& d) {
    return d.sparse == B200CB_SPARSE_GROUP && d.group == d.ic && d.ic == d.oc;
}
b200cb_desc nchw_equivalent(const b200cb_desc& d) {
    b200cb_desc e = d;
    e.format = B200CB_FORMAT_NHWC;
    if (!nchw_depthwise(d)) {
        // stems (IC <= 4, stride 2) take 8-byte pixels, everything else 16-byte channel granules
        const bool stem_like = d.ic <= 4 && d.stride_w == 2 && d.stride_h == 2 && d.iw % 2 == 0 &&
                               d.dilate_w == 1 && d.dilate_h == 1 && d.mode == B200CB_MODE_CROSS_CORRELATION;
        e.ic = stem_like ? 4 : (d.ic + 7) / 8 * 8;
    }
    e.src_batch_stride = e.z_batch_stride = e.dst_batch_stride = 0;
    return e;
}
size_t a256(size_t v) {
    return (v + 255) & ~size_t(255);
}
struct NchwLayout {
    size_t src, flt, z, dst, inner, total;
};
NchwLayout nchw_layout(const b200cb_desc& d, const b200cb_desc& e, uint64_t inner_ws) {
    NchwLayout l{};
    size_t off = 0;
    l.src = off, off += a256((size_t)e.n * e.ih * e.iw * e.ic * 2);
    l.flt = off, off += nchw_depthwise(d) ? 0 : a256((size_t)e.oc * e.fh * e.fw * e.ic * 2);
    l.z = off, off += d.has_z ? a256((size_t)e.n * e.oh * e.ow * e.oc * 2) : 0;
    l.dst = off, off += a256((size_t)e.n * e.oh * e.ow * e.oc * 2);
    l.inner = off, off += a256(inner_ws);
    l.total = off;
    return l;
}
// ---- Float32 adapter (dense NHWC / NCHW): BASELINE configs[0]'s dtype on the tensor-core path.  The
// operands are split into bf16 pieces on the fly (relayout_split3: x = x0 + x1 + O(2^-17)), the bf16 kernel
// runs over 3 x the channels with its fp32 accumulator written out as Float32 (Float32 bias / z added in the
// epilogue): products carry ~2^-16 relative error, well inside the reference's own Float32 tolerance
// (1e-3, dnn/test/cuda/conv_bias.cpp:83-100).  NCHW tensors are re-tiled to NHWC first (b200cb_relayout).
bool is_f32_dense(const b200cb_desc& d) {
    return d.src_dtype == B200CB_DT_F32 && d.filter_dtype == B200CB_DT_F32 && d.dst_dtype == B200CB_DT_F32 &&
           d.group == 1 && d.sparse == B200CB_SPARSE_DENSE &&
           (d.format == B200CB_FORMAT_NHWC || d.format == B200CB_FORMAT_NCHW);
}
b200cb_desc f32_equivalent(const b200cb_desc& d) {
    b200cb_desc e = d;
    e.format = B200CB_FORMAT_NHWC;
    e.src_dtype = e.filter_dtype = B200CB_DT_BF16;
    e.compute_mode = B200CB_COMPUTE_FLOAT32;
    e.ic = 3 * ((d.ic + 7) / 8 * 8);
    e.src_batch_stride = e.z_batch_stride = e.dst_batch_stride = 0;
    return e;
}
struct F32Layout {
    size_t src_t, src3, flt_t, flt3, z_t, dst_t, inner, total;
};
F32Layout f32_layout(const b200cb_desc& d, const b200cb_desc& e, uint64_t inner_ws) {
    const bool nchw = d.format == B200CB_FORMAT_NCHW;
    const size_t dst_elems = (size_t)d.n * d.oh * d.ow * d.oc;
    F32Layout l{};
    size_t off = 0;
    l.src_t = off, off += nchw ? a256((size_t)d.n * d.ih * d.iw * d.ic * 4) : 0;
    l.src3 = off, off += a256((size_t)d.n * d.ih * d.iw * e.ic * 2);
    l.flt_t = off, off += nchw ? a256((size_t)d.oc * d.fh * d.fw * d.ic * 4) : 0;
    l.flt3 = off, off += a256((size_t)d.oc * d.fh * d.fw * e.ic * 2);
    l.z_t = off, off += (nchw && d.has_z) ? a256(dst_elems * 4) : 0;
    l.dst_t = off, off += nchw ? a256(dst_elems * 4) : 0;
    l.inner = off, off += a256(inner_ws);
    l.total = off;
    return l;
}

// ---- pixel-pair fold of narrow 1x1 layers (MobileNetV2 / ShuffleNetV2 projections such as 32 -> 16 at 112x112): the
// layer is bound by the number of short TMA rows per tile and stores 16 of a 64-column tile.  Two horizontally adjacent
// pixels are viewed as one pixel of 2 * IC channels -- src, z and dst are the SAME memory under that view -- and the
// filter becomes the block-diagonal (2 * OC) x (2 * IC) matrix [[W, 0], [0, W]] (built in the workspace by one tiny
// launch per exec, with the bias repeated).  Half the tiles, activation rows twice as long, every product of the
// original sum in the same order (the added terms are exact zeros): results are bit-identical.
bool pw_fold_enabled() {
#ifdef B200CB_TOOLS
    if (const char* e = getenv("B200CB_PWFOLD"))
        return atoi(e) != 0;
#endif
    return true;
}
bool is_pw_fold(const b200cb_desc& d) {
    const bool s16 = (d.src_dtype == B200CB_DT_BF16 || d.src_dtype == B200CB_DT_F16) && d.dst_dtype == d.src_dtype;
    const bool s8 = d.src_dtype == B200CB_DT_QS8 && d.dst_dtype == B200CB_DT_QS8;
    if (!(s16 || s8) || d.filter_dtype != d.src_dtype || d.format != B200CB_FORMAT_NHWC || d.group > 1 ||
        d.sparse != B200CB_SPARSE_DENSE)
        return false;
    if (d.fh != 1 || d.fw != 1 || d.stride_h != 1 || d.stride_w != 1 || d.pad_h || d.pad_w || d.iw % 2 != 0)
        return false;
    if (d.src_batch_stride || d.z_batch_stride || d.dst_batch_stride || d.bias_mode == B200CB_BIAS_FULL)
        return false;
    const uint32_t es = dtype_size(d.src_dtype);
    // narrow on both sides: the folded layer still fits one 64-column tile and one 128-byte k-block
    return 2 * d.oc <= 64 && 2 * d.ic * es <= 128 && (d.ic * es) % 16 == 0 && (d.oc * es) % 16 == 0 && pw_fold_enabled();
}
b200cb_desc pw_fold_equivalent(const b200cb_desc& d) {
    b200cb_desc e = d;
    e.ic = 2 * d.ic, e.oc = 2 * d.oc, e.iw = d.iw / 2, e.ow = d.ow / 2;
    return e;
}
struct PwFoldLayout {
    size_t flt, bias, inner, total;
};
PwFoldLayout pw_fold_layout(const b200cb_desc& d, uint64_t inner_ws) {
    PwFoldLayout l{};
    size_t off = 0;
    l.flt = off, off += a256((size_t)4 * d.oc * d.ic * dtype_size(d.filter_dtype));
    l.bias = off, off += a256((size_t)2 * d.oc * 4);
    l.inner = off, off += a256(inner_ws);
    l.total = off;
    return l;
}
// filter (OC, IC) -> [[W, 0], [0, W]] (2 OC, 2 IC), bias (OC) -> (2 OC); one thread per 16-bit / 8-bit / 32-bit element
template <typename T>
__global__ void pw_fold_kernel(
        const T* __restrict__ w, T* __restrict__ wf, uint32_t oc, uint32_t ic, const void* __restrict__ bias,
        void* __restrict__ bf, uint32_t bias_es) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * oc * ic) {
        const uint32_t r = i / (2 * ic), c = i % (2 * ic);   // row of the folded filter, column
        const uint32_t ro = r % oc, half_r = r / oc, half_c = c / ic;
        wf[i] = half_r == half_c ? w[ro * ic + c % ic] : T(0);
    }
    if (bias && i < 2 * oc) {
        if (bias_es == 2)
            static_cast<uint16_t*>(bf)[i] = static_cast<const uint16_t*>(bias)[i % oc];
        else
            static_cast<uint32_t*>(bf)[i] = static_cast<const uint32_t*>(bias)[i % oc];
    }
}

int query_one(const b200cb_desc& d, int algo, b200cb_plan* plan) {
    const AlgoEntry& a = g_algos[algo];
    return a.kernel == KERNEL_IGEMM ? igemm_query(d, a.block_n, plan) : dwconv_query(d, plan);
}

}  // namespace

void set_last_error(const char* what, const char* detail) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what ? what : "", detail ? detail : "");
}

}  // namespace b200cb

using namespace b200cb;

extern "C" {

int b200cb_algo_count(void) {
    return NUM_ALGOS;
}

const char* b200cb_algo_name(int algo) {
    if (algo < 0 || algo >= NUM_ALGOS)
        return nullptr;
    return g_algos[algo].name;
}

int b200cb_query_algo(const b200cb_desc* desc, int algo, b200cb_plan* plan) {
    if (!desc || !plan)
        return B200CB_ERR_NULL;
    if (algo < 0 || algo >= NUM_ALGOS)
        return B200CB_UNSUPPORTED_ALGO;
    int rc = validate_desc(desc);
    if (rc != B200CB_OK)
        return rc;
    memset(plan, 0, sizeof(*plan));
    if (is_f32_dense(*desc)) {
        if (desc->bias_mode == B200CB_BIAS_FULL || (desc->bias_mode != B200CB_BIAS_NONE && desc->bias_dtype != B200CB_DT_F32) ||
            (desc->has_z && desc->z_dtype != B200CB_DT_F32) || desc->oc % 4 != 0)
            return desc->oc % 4 ? B200CB_UNSUPPORTED_ALIGNMENT : B200CB_UNSUPPORTED_DTYPE;
        if (desc->src_batch_stride || desc->z_batch_stride || desc->dst_batch_stride)
            return B200CB_UNSUPPORTED_SHAPE;
        const b200cb_desc e = f32_equivalent(*desc);
        rc = query_one(e, algo, plan);
        if (rc != B200CB_OK)
            return rc;
        plan->workspace_bytes = f32_layout(*desc, e, plan->workspace_bytes).total;
        plan->desc = *desc;
    } else if (is_nchw16(*desc)) {
        if (desc->bias_mode == B200CB_BIAS_FULL || (desc->oc % 8 != 0 && !nchw_depthwise(*desc)))
            return B200CB_UNSUPPORTED_ALIGNMENT;
        const b200cb_desc e = nchw_equivalent(*desc);
        rc = query_one(e, algo, plan);
        if (rc != B200CB_OK)
            return rc;
        plan->workspace_bytes = nchw_layout(*desc, e, plan->workspace_bytes).total;
        plan->desc = *desc;
    } else if (g_algos[algo].kernel == KERNEL_IGEMM && is_pw_fold(*desc)) {
        rc = query_one(pw_fold_equivalent(*desc), algo, plan);
        if (rc != B200CB_OK)
            return rc;
        plan->workspace_bytes = pw_fold_layout(*desc, plan->workspace_bytes).total;
        plan->desc = *desc;
    } else {
        rc = query_one(*desc, algo, plan);
        if (rc != B200CB_OK)
            return rc;
    }
    plan->algo = algo;
    fill_accounting(plan);
    return B200CB_OK;
}

int b200cb_query(const b200cb_desc* desc, b200cb_plan* plan) {
    if (!desc || !plan)
        return B200CB_ERR_NULL;
    int first_err = B200CB_UNSUPPORTED_ALGO;
    int best = -1;
    double best_cost = 0;
    b200cb_plan tmp;
    for (int a = 0; a < NUM_ALGOS; ++a) {
        int rc = b200cb_query_algo(desc, a, &tmp);
        if (rc != B200CB_OK) {
            if (rc < 0)
                return rc;
            if (first_err == B200CB_UNSUPPORTED_ALGO)
                first_err = rc;
            continue;
        }
        double cost = g_algos[a].kernel == KERNEL_IGEMM ? igemm_cost(*desc, tmp, g_algos[a].block_n, 148) : 0.0;
        if (best < 0 || cost < best_cost) {
            best = a;
            best_cost = cost;
            *plan = tmp;
        }
    }
    return best >= 0 ? B200CB_OK : first_err;
}

int b200cb_launch(
        const b200cb_plan* plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* workspace, size_t workspace_bytes, void* stream) {
    if (!plan)
        return B200CB_ERR_NULL;
    if (plan->magic != PLAN_MAGIC)
        return B200CB_ERR_BAD_PLAN;
    if (workspace_bytes < plan->workspace_bytes || (plan->workspace_bytes && !workspace))
        return B200CB_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_f32_dense(plan->desc)) {
        const b200cb_desc& d = plan->desc;
        const b200cb_desc e = f32_equivalent(d);
        b200cb_plan inner;
        int rc = b200cb_query_algo(&e, plan->algo, &inner);
        if (rc != B200CB_OK)
            return rc > 0 ? B200CB_ERR_BAD_PLAN : rc;
        const F32Layout l = f32_layout(d, e, inner.workspace_bytes);
        if (!workspace || workspace_bytes < l.total)
            return B200CB_ERR_WORKSPACE;
        if (!src || !filter || !dst || (d.has_z && !z) || (d.bias_mode != B200CB_BIAS_NONE && !bias))
            return B200CB_ERR_NULL;
        uint8_t* ws = static_cast<uint8_t*>(workspace);
        const bool nchw = d.format == B200CB_FORMAT_NCHW;
        const uint32_t cp = e.ic / 3;
        const void *src_nhwc = src, *flt_krsc = filter, *z_nhwc = z;
        if (nchw) {
            rc = relayout_generic(B200CB_RELAYOUT_NCHW_NHWC, 4, d.n, d.ic, d.ih, d.iw, d.ic, src, ws + l.src_t, s);
            if (rc == B200CB_OK)
                rc = relayout_generic(B200CB_RELAYOUT_NCHW_NHWC, 4, d.oc, d.ic, d.fh, d.fw, d.ic, filter, ws + l.flt_t, s);
            if (rc == B200CB_OK && d.has_z)
                rc = relayout_generic(B200CB_RELAYOUT_NCHW_NHWC, 4, d.n, d.oc, d.oh, d.ow, d.oc, z, ws + l.z_t, s);
            if (rc != B200CB_OK)
                return rc;
            src_nhwc = ws + l.src_t, flt_krsc = ws + l.flt_t, z_nhwc = d.has_z ? ws + l.z_t : nullptr;
        }
        rc = relayout_split3(src_nhwc, ws + l.src3, (uint64_t)d.n * d.ih * d.iw, d.ic, cp, 0, s);
        if (rc == B200CB_OK)
            rc = relayout_split3(flt_krsc, ws + l.flt3, (uint64_t)d.oc * d.fh * d.fw, d.ic, cp, 1, s);
        if (rc == B200CB_OK)
            rc = b200cb_launch(
                    &inner, ws + l.src3, ws + l.flt3, bias, z_nhwc, nchw ? static_cast<void*>(ws + l.dst_t) : dst,
                    ws + l.inner, inner.workspace_bytes, stream);
        if (rc == B200CB_OK && nchw)
            rc = relayout_generic(B200CB_RELAYOUT_NHWC_NCHW, 4, d.n, d.oc, d.oh, d.ow, d.oc, ws + l.dst_t, dst, s);
        return rc;
    }
    if (is_nchw16(plan->desc)) {
        const b200cb_desc& d = plan->desc;
        const b200cb_desc e = nchw_equivalent(d);
        b200cb_plan inner;
        int rc = b200cb_query_algo(&e, plan->algo, &inner);
        if (rc != B200CB_OK)
            return rc > 0 ? B200CB_ERR_BAD_PLAN : rc;
        const NchwLayout l = nchw_layout(d, e, inner.workspace_bytes);
        if (!workspace || workspace_bytes < l.total || !src || !filter || !dst || (d.has_z && !z))
            return !workspace || workspace_bytes < l.total ? B200CB_ERR_WORKSPACE : B200CB_ERR_NULL;
        uint8_t* ws = static_cast<uint8_t*>(workspace);
        const uint32_t P = d.ih * d.iw, OP = d.oh * d.ow;
        // src (N, C, P) -> (N, P, Cpad); filter (OC, IC, FHFW) -> (OC, FHFW, ICpad) (depthwise: same order)
        rc = relayout_transpose16(src, ws + l.src, d.n, d.ic, P, e.ic, d.src_batch_stride, 0, s);
        const void* flt = filter;
        if (rc == B200CB_OK && !nchw_depthwise(d)) {
            rc = relayout_transpose16(filter, ws + l.flt, d.oc, d.ic, d.fh * d.fw, e.ic, 0, 0, s);
            flt = ws + l.flt;
        }
        if (rc == B200CB_OK && d.has_z)
            rc = relayout_transpose16(z, ws + l.z, d.n, d.oc, OP, d.oc, d.z_batch_stride, 0, s);
        if (rc == B200CB_OK)
            rc = b200cb_launch(
                    &inner, ws + l.src, flt, bias, d.has_z ? ws + l.z : nullptr, ws + l.dst, ws + l.inner,
                    inner.workspace_bytes, stream);
        // dst (N, OP, OC) -> (N, OC, OP)
        if (rc == B200CB_OK)
            rc = relayout_transpose16(ws + l.dst, dst, d.n, OP, d.oc, OP, 0, d.dst_batch_stride, s);
        return rc;
    }
    if (plan->kernel == KERNEL_IGEMM && is_pw_fold(plan->desc)) {
        const b200cb_desc& d = plan->desc;
        b200cb_plan inner;
        int rc = query_one(pw_fold_equivalent(d), plan->algo, &inner);
        if (rc != B200CB_OK)
            return rc > 0 ? B200CB_ERR_BAD_PLAN : rc;
        const PwFoldLayout l = pw_fold_layout(d, inner.workspace_bytes);
        if (!workspace || workspace_bytes < l.total)
            return B200CB_ERR_WORKSPACE;
        if (!src || !filter || !dst || (d.has_z && !z) || (d.bias_mode != B200CB_BIAS_NONE && !bias))
            return B200CB_ERR_NULL;
        uint8_t* ws = static_cast<uint8_t*>(workspace);
        const uint32_t total = 4 * d.oc * d.ic, blocks = (total + 255) / 256;
        const void* b_in = d.bias_mode != B200CB_BIAS_NONE ? bias : nullptr;
        if (dtype_size(d.filter_dtype) == 1)
            pw_fold_kernel<int8_t><<<blocks, 256, 0, s>>>(
                    static_cast<const int8_t*>(filter), reinterpret_cast<int8_t*>(ws + l.flt), d.oc, d.ic, b_in,
                    ws + l.bias, dtype_size(d.bias_dtype));
        else
            pw_fold_kernel<uint16_t><<<blocks, 256, 0, s>>>(
                    static_cast<const uint16_t*>(filter), reinterpret_cast<uint16_t*>(ws + l.flt), d.oc, d.ic, b_in,
                    ws + l.bias, dtype_size(d.bias_dtype));
        if (cudaGetLastError() != cudaSuccess) {
            set_last_error("pw_fold launch", "cuda error");
            return B200CB_ERR_CUDA;
        }
        inner.algo = plan->algo;
        return igemm_launch(
                inner, src, ws + l.flt, b_in ? ws + l.bias : nullptr, z, dst, ws + l.inner, inner.workspace_bytes, s);
    }
    switch (plan->kernel) {
        case KERNEL_IGEMM:
            return igemm_launch(*plan, src, filter, bias, z, dst, workspace, workspace_bytes, s);
        case KERNEL_DWCONV:
            return dwconv_launch(*plan, src, filter, bias, z, dst, s);
        default:
            return B200CB_ERR_BAD_PLAN;
    }
}

int b200cb_plan_num_launches(const b200cb_plan* plan) {
    if (!plan || plan->magic != PLAN_MAGIC)
        return 0;
    if (is_f32_dense(plan->desc)) {  // split src + filter, the bf16 plan, and for NCHW the re-tiling passes
        const b200cb_desc e = f32_equivalent(plan->desc);
        b200cb_plan inner;
        if (b200cb_query_algo(&e, plan->algo, &inner) != B200CB_OK)
            return 0;
        const bool nchw = plan->desc.format == B200CB_FORMAT_NCHW;
        return 2 + (nchw ? 3 + (plan->desc.has_z ? 1 : 0) : 0) + b200cb_plan_num_launches(&inner);
    }
    if (plan->kernel == KERNEL_IGEMM && plan->desc.group > 1 && plan->desc.format == B200CB_FORMAT_NHWC)
        return static_cast<int>(plan->desc.group);  // one dense launch per group
    if (is_nchw16(plan->desc)) {  // re-tile src (+ filter, + z), the NHWC plan, re-tile dst
        const b200cb_desc e = nchw_equivalent(plan->desc);
        b200cb_plan inner;
        if (b200cb_query_algo(&e, plan->algo, &inner) != B200CB_OK)
            return 0;
        return 2 + (nchw_depthwise(plan->desc) ? 0 : 1) + (plan->desc.has_z ? 1 : 0) + b200cb_plan_num_launches(&inner);
    }
    if (plan->kernel == KERNEL_IGEMM && is_pw_fold(plan->desc))
        return 2;  // block-diagonal filter + conv
    // NCHW4: src re-tile + filter re-tile (+ stem filter re-pack) + conv
    if (plan->kernel == KERNEL_IGEMM && plan->desc.format == B200CB_FORMAT_NCHW4)
        return plan->reserved[4] ? 4 : 3;
    // W-folded stem without weight preprocess: filter re-pack + conv
    if (plan->kernel == KERNEL_IGEMM && plan->reserved[4] && !plan->reserved[3])
        return 2;
    return 1;
}

size_t b200cb_preprocessed_filter_bytes(const b200cb_plan* plan) {
    // KRSC filters are consumed in place by TMA; only the W-folded stem needs a packed copy.
    if (!plan || plan->magic != PLAN_MAGIC || plan->kernel != KERNEL_IGEMM || is_nchw16(plan->desc) ||
        is_f32_dense(plan->desc))
        return 0;
    return igemm_packed_filter_bytes(*plan);
}

int b200cb_preprocess_filter(const b200cb_plan* plan, const void* filter, void* packed, void* stream) {
    if (!plan)
        return B200CB_ERR_NULL;
    if (plan->magic != PLAN_MAGIC)
        return B200CB_ERR_BAD_PLAN;
    if (plan->kernel != KERNEL_IGEMM || !plan->reserved[4] || plan->desc.format != B200CB_FORMAT_NHWC)
        return B200CB_OK;  // nothing to pack (NCHW4 filters are re-tiled per exec)
    if (!filter || !packed)
        return B200CB_ERR_NULL;
    return igemm_pack_filter(*plan, filter, packed, static_cast<cudaStream_t>(stream));
}

int b200cb_plan_use_preprocessed_filter(b200cb_plan* plan) {
    if (!plan)
        return B200CB_ERR_NULL;
    if (plan->magic != PLAN_MAGIC)
        return B200CB_ERR_BAD_PLAN;
    if (plan->kernel == KERNEL_IGEMM && plan->reserved[4] && plan->desc.format == B200CB_FORMAT_NHWC) {
        plan->reserved[3] = 1;
        plan->workspace_bytes = 0;
    }
    return B200CB_OK;
}

int b200cb_relayout(
        uint32_t mode, uint32_t elem_bytes, uint32_t n, uint32_t c, uint32_t h, uint32_t w, uint32_t c_out,
        const void* src, void* dst, void* stream) {
    if (!src || !dst)
        return B200CB_ERR_NULL;
    if (n == 0 || c == 0 || h == 0 || w == 0 || c_out == 0)
        return B200CB_UNSUPPORTED_SHAPE;
    return relayout_generic(mode, elem_bytes, n, c, h, w, c_out, src, dst, static_cast<cudaStream_t>(stream));
}

int b200cb_selftest(int which, void* stream, uint64_t* mismatches) {
    return igemm_selftest(which, static_cast<cudaStream_t>(stream), mismatches);
}

namespace {
struct HostLayout {
    size_t src, filter, bias, z, dst, ws, total;
    size_t src_b, filter_b, bias_b, z_b, dst_b;
};
size_t align256(size_t v) {
    return (v + 255) & ~size_t(255);
}
HostLayout host_layout(const b200cb_plan* plan) {
    const b200cb_desc& d = plan->desc;
    HostLayout l{};
    const size_t src_img = (size_t)d.ih * d.iw * d.ic;
    l.src_b = (size_t)d.n * src_img * dtype_size(d.src_dtype);
    l.filter_b = (size_t)d.oc * (d.ic / d.group) * d.fh * d.fw * dtype_size(d.filter_dtype);
    const size_t dst_elems = (size_t)d.n * d.oh * d.ow * d.oc;
    l.bias_b = d.bias_mode == B200CB_BIAS_CHANNEL ? (size_t)d.oc * dtype_size(d.bias_dtype)
               : d.bias_mode == B200CB_BIAS_FULL  ? dst_elems * dtype_size(d.bias_dtype)
                                                  : 0;
    l.z_b = d.has_z ? dst_elems * dtype_size(d.z_dtype) : 0;
    l.dst_b = dst_elems * dtype_size(d.dst_dtype);
    size_t off = 0;
    l.src = off, off += align256(l.src_b);
    l.filter = off, off += align256(l.filter_b);
    l.bias = off, off += align256(l.bias_b);
    l.z = off, off += align256(l.z_b);
    l.dst = off, off += align256(l.dst_b);
    l.ws = off, off += align256(plan->workspace_bytes);
    l.total = off;
    return l;
}
}  // namespace

size_t b200cb_host_scratch_bytes(const b200cb_plan* plan) {
    if (!plan || plan->magic != PLAN_MAGIC)
        return 0;
    return host_layout(plan).total;
}

int b200cb_exec_host(
        const b200cb_plan* plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    if (!plan || !src || !filter || !dst || !dev_scratch)
        return B200CB_ERR_NULL;
    if (plan->magic != PLAN_MAGIC)
        return B200CB_ERR_BAD_PLAN;
    const b200cb_desc& d = plan->desc;
    if (d.src_batch_stride || d.dst_batch_stride || d.z_batch_stride)
        return B200CB_UNSUPPORTED_SHAPE;  // host convenience path takes dense tensors only
    HostLayout l = host_layout(plan);
    if (dev_scratch_bytes < l.total)
        return B200CB_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(dev_scratch);
    cudaError_t e = cudaMemcpyAsync(base + l.src, src, l.src_b, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(base + l.filter, filter, l.filter_b, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && l.bias_b)
        e = cudaMemcpyAsync(base + l.bias, bias, l.bias_b, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && l.z_b)
        e = cudaMemcpyAsync(base + l.z, z, l.z_b, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
        set_last_error("exec_host H2D", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    int rc = b200cb_launch(
            plan, base + l.src, base + l.filter, l.bias_b ? base + l.bias : nullptr,
            l.z_b ? base + l.z : nullptr, base + l.dst, base + l.ws, plan->workspace_bytes, stream);
    if (rc != B200CB_OK)
        return rc;
    e = cudaMemcpyAsync(dst, base + l.dst, l.dst_b, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        set_last_error("exec_host D2H", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

const char* b200cb_strerror(int code) {
    switch (code) {
        case B200CB_OK: return "ok";
        case B200CB_UNSUPPORTED_FORMAT: return "unsupported tensor format";
        case B200CB_UNSUPPORTED_DTYPE: return "unsupported dtype combination";
        case B200CB_UNSUPPORTED_SHAPE: return "unsupported shape";
        case B200CB_UNSUPPORTED_PARAM: return "unsupported convolution parameter";
        case B200CB_UNSUPPORTED_ALIGNMENT: return "channel count / stride not 16-byte aligned";
        case B200CB_UNSUPPORTED_ALGO: return "no such algorithm / algorithm not applicable";
        case B200CB_ERR_CUDA: return "CUDA runtime error";
        case B200CB_ERR_DRIVER: return "CUDA driver error (tensor map encode)";
        case B200CB_ERR_WORKSPACE: return "workspace too small";
        case B200CB_ERR_BAD_PLAN: return "plan was not produced by b200cb_query";
        case B200CB_ERR_NULL: return "null pointer argument";
        default: return "unknown error";
    }
}

const char* b200cb_last_error_detail(void) {
    return g_last_error;
}

}  // extern "C"

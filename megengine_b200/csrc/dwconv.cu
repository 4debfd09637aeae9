// Depthwise (channel-wise) NHWC ConvBias: memory-bound CUDA-core kernel.
//
//   One thread owns a 16-byte channel vector (8 x bf16/fp16, 4 x fp32, 16 x int8) and TW
//   horizontally adjacent output pixels, so every input vector it loads is reused by up to
//   FW x TW taps from registers.  Consecutive lanes own consecutive channel vectors, i.e. a warp
//   reads / writes contiguous 16-byte pieces of NHWC rows (coalesced, ld.global.nc / 128-bit).
//   The block's filter slice is staged once in shared memory, transposed to [tap][channel] and
//   widened to 32 bit, and the block then grid-strides over pixel groups (persistent-style, grid
//   sized from the SM count).  bias + z + activation (+ requantize) are fused: one pass over HBM.
//
// Replaces the reference's dnn/src/cuda/conv_bias/chanwise/fwd.cu (NCHW only, bias/activation in
// a second Elemwise pass, z unsupported: chanwise.cpp:18-19,88-90).
// Filter layout (NHWC, GROUP, icpg == ocpg == 1): (C, 1, FH, FW, 1)
//   -> element (c, fh, fw) at c*FH*FW + fh*FW + fw   (dnn/src/common/convolution.cpp:33-94).
#include <stdio.h>
#include <string.h>
#include "common.h"
#include "epilogue.cuh"

namespace b200cb {

struct DwParams {
    const void* src;
    const void* filter;
    const void* bias;
    const void* z;
    void* dst;
    uint32_t N, C, IH, IW, OH, OW;
    uint32_t FH, FW;
    int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    uint32_t CV;        // channel vectors in total (C / VEC)
    uint32_t CVB;       // channel vectors per block slice
    uint32_t PGB;       // pixel groups per block iteration (256 / CVB)
    uint32_t OWG;       // pixel groups per output row
    uint64_t total_pg;  // N * OH * OWG
    uint64_t src_bstride, z_bstride, dst_bstride;  // elements
    uint32_t flip, has_bias, has_z, nonline;
    RequantParams rq;
};

namespace {

template <int KIND>
struct DwTraits;
template <>
struct DwTraits<B200CB_DT_BF16> {
    static constexpr int ES = 2, VEC = 8;
    using acc_t = float;
};
template <>
struct DwTraits<B200CB_DT_F16> {
    static constexpr int ES = 2, VEC = 8;
    using acc_t = float;
};
template <>
struct DwTraits<B200CB_DT_F32> {
    static constexpr int ES = 4, VEC = 4;
    using acc_t = float;
};
template <>
struct DwTraits<B200CB_DT_QS8> {
    static constexpr int ES = 1, VEC = 16;
    using acc_t = int32_t;
};

__device__ __forceinline__ uint4 ldg_nc_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// 16-byte vector -> VEC accumulator-typed lanes
template <int KIND>
__device__ __forceinline__ void unpack(const uint4& v, typename DwTraits<KIND>::acc_t* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if constexpr (KIND == B200CB_DT_BF16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out[2 * i] = bf16lo(w[i]);
            out[2 * i + 1] = bf16hi(w[i]);
        }
    } else if constexpr (KIND == B200CB_DT_F16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out[2 * i] = f16lo(w[i]);
            out[2 * i + 1] = f16hi(w[i]);
        }
    } else if constexpr (KIND == B200CB_DT_F32) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            out[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                out[4 * i + b] = static_cast<int32_t>(static_cast<int8_t>((w[i] >> (8 * b)) & 0xff));
        }
    }
}

template <int KIND>
__device__ __forceinline__ uint32_t load_filter_elem(const void* f, size_t idx) {
    if constexpr (KIND == B200CB_DT_BF16)
        return static_cast<uint32_t>(reinterpret_cast<const unsigned short*>(f)[idx]) << 16;
    else if constexpr (KIND == B200CB_DT_F16)
        return __float_as_uint(__half2float(reinterpret_cast<const __half*>(f)[idx]));
    else if constexpr (KIND == B200CB_DT_F32)
        return reinterpret_cast<const uint32_t*>(f)[idx];
    else
        return static_cast<uint32_t>(static_cast<int32_t>(reinterpret_cast<const int8_t*>(f)[idx]));
}

// bias element -> accumulator domain (fp32 for float kinds, int32 for QuantizedS32 bias)
template <int KIND>
__device__ __forceinline__ typename DwTraits<KIND>::acc_t load_bias_elem(const void* b, size_t idx) {
    if constexpr (KIND == B200CB_DT_BF16)
        return __uint_as_float(static_cast<uint32_t>(reinterpret_cast<const unsigned short*>(b)[idx]) << 16);
    else if constexpr (KIND == B200CB_DT_F16)
        return __half2float(reinterpret_cast<const __half*>(b)[idx]);
    else if constexpr (KIND == B200CB_DT_F32)
        return reinterpret_cast<const float*>(b)[idx];
    else
        return reinterpret_cast<const int32_t*>(b)[idx];
}

// F = filter size (square), S = stride, TW = outputs per thread along W. F == 0: generic
// run-time filter / stride / dilation with TW == 1.
template <int KIND, int F, int S, int TW>
__global__ void __launch_bounds__(256) dwconv_nhwc_kernel(const DwParams p) {
    using T = DwTraits<KIND>;
    using acc_t = typename T::acc_t;
    constexpr int VEC = T::VEC;
    constexpr int INW = (F > 0) ? (TW - 1) * S + F : 1;
    extern __shared__ uint32_t sfilt[];  // [FH*FW][CVB*VEC], 32-bit widened

    const uint32_t FH = F > 0 ? F : p.FH, FW = F > 0 ? F : p.FW;
    const int32_t SH = F > 0 ? S : p.stride_h, SW = F > 0 ? S : p.stride_w;
    const uint32_t taps = FH * FW;
    const uint32_t chb = p.CVB * VEC;
    const uint32_t tid = threadIdx.x;
    const uint32_t cvl = tid % p.CVB, pgl = tid / p.CVB;
    const uint32_t slice = blockIdx.y;
    const uint32_t c0 = (slice * p.CVB + cvl) * VEC;

    for (uint32_t i = tid; i < taps * chb; i += blockDim.x) {
        const uint32_t tap = i / chb, ch = i % chb;
        const uint32_t gc = slice * chb + ch;
        const uint32_t ftap = p.flip ? (taps - 1 - tap) : tap;
        sfilt[i] = gc < p.C ? load_filter_elem<KIND>(p.filter, (size_t)gc * taps + ftap) : 0u;
    }
    __syncthreads();
    if (pgl >= p.PGB || c0 >= p.C)
        return;

    acc_t bias[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        bias[i] = p.has_bias ? load_bias_elem<KIND>(p.bias, c0 + i) : acc_t(0);

    const uint8_t* src = static_cast<const uint8_t*>(p.src);
    const uint8_t* zp = static_cast<const uint8_t*>(p.z);
    uint8_t* dst = static_cast<uint8_t*>(p.dst);
    const uint32_t* wbase = sfilt + cvl * VEC;

    for (uint64_t pg = (uint64_t)blockIdx.x * p.PGB + pgl; pg < p.total_pg;
         pg += (uint64_t)gridDim.x * p.PGB) {
        const uint32_t owg = (uint32_t)(pg % p.OWG);
        const uint64_t t1 = pg / p.OWG;
        const uint32_t oh = (uint32_t)(t1 % p.OH);
        const uint32_t n = (uint32_t)(t1 / p.OH);
        const uint32_t ow0 = owg * TW;

        acc_t acc[TW][VEC];
#pragma unroll
        for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                acc[t][i] = acc_t(0);

        const uint8_t* img = src + (size_t)n * p.src_bstride * T::ES;
        if constexpr (F > 0) {
            const int32_t iw0 = (int32_t)ow0 * S - p.pad_w;
#pragma unroll
            for (int fh = 0; fh < F; ++fh) {
                const int32_t ih = (int32_t)oh * S + fh - p.pad_h;
                if ((uint32_t)ih >= p.IH)
                    continue;
                const uint8_t* row = img + ((size_t)ih * p.IW * p.C + c0) * T::ES;
                uint4 raw[INW];
#pragma unroll
                for (int j = 0; j < INW; ++j) {
                    const int32_t iw = iw0 + j;
                    raw[j] = ((uint32_t)iw < p.IW) ? ldg_nc_16(row + (size_t)iw * p.C * T::ES)
                                                  : make_uint4(0, 0, 0, 0);
                }
                acc_t x[INW][VEC];
#pragma unroll
                for (int j = 0; j < INW; ++j)
                    unpack<KIND>(raw[j], x[j]);
#pragma unroll
                for (int fw = 0; fw < F; ++fw) {
                    acc_t w[VEC];
                    const uint32_t* wp = wbase + (fh * F + fw) * chb;
#pragma unroll
                    for (int i = 0; i < VEC; i += 4) {
                        const uint4 wv = *reinterpret_cast<const uint4*>(wp + i);
                        if constexpr (KIND == B200CB_DT_QS8) {
                            w[i] = (int32_t)wv.x, w[i + 1] = (int32_t)wv.y;
                            w[i + 2] = (int32_t)wv.z, w[i + 3] = (int32_t)wv.w;
                        } else {
                            w[i] = __uint_as_float(wv.x), w[i + 1] = __uint_as_float(wv.y);
                            w[i + 2] = __uint_as_float(wv.z), w[i + 3] = __uint_as_float(wv.w);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < TW; ++t)
#pragma unroll
                        for (int i = 0; i < VEC; ++i)
                            acc[t][i] += x[t * S + fw][i] * w[i];
                }
            }
        } else {
            for (uint32_t fh = 0; fh < FH; ++fh) {
                const int32_t ih = (int32_t)oh * SH + (int32_t)fh * p.dil_h - p.pad_h;
                if ((uint32_t)ih >= p.IH)
                    continue;
                for (uint32_t fw = 0; fw < FW; ++fw) {
                    const int32_t iw = (int32_t)ow0 * SW + (int32_t)fw * p.dil_w - p.pad_w;
                    if ((uint32_t)iw >= p.IW)
                        continue;
                    const uint4 raw = ldg_nc_16(img + (((size_t)ih * p.IW + iw) * p.C + c0) * T::ES);
                    acc_t x[VEC];
                    unpack<KIND>(raw, x);
                    const uint32_t* wp = wbase + (fh * FW + fw) * chb;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        if constexpr (KIND == B200CB_DT_QS8)
                            acc[0][i] += x[i] * (int32_t)wp[i];
                        else
                            acc[0][i] += x[i] * __uint_as_float(wp[i]);
                    }
                }
            }
        }

        // ---------------- fused epilogue: bias, z, activation, (re)quantize, 16-byte store
#pragma unroll
        for (int t = 0; t < TW; ++t) {
            const uint32_t ow = ow0 + t;
            if (ow >= p.OW)
                break;
            const size_t pix = ((size_t)oh * p.OW + ow) * p.C + c0;
            uint4 zv = make_uint4(0, 0, 0, 0);
            if (p.has_z)
                zv = ldg_stream_16(zp + ((size_t)n * p.z_bstride + pix) * T::ES);
            uint4 ov;
            if constexpr (KIND == B200CB_DT_QS8) {
                const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
                uint32_t o[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t packed = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int32_t sfb = acc[t][4 * w + b] + bias[4 * w + b];
                        const int32_t zq = (int32_t)(int8_t)((zw[w] >> (8 * b)) & 0xff);
                        const int32_t q = requant_s8(sfb, zq, p.nonline, p.rq);
                        packed |= ((uint32_t)q & 0xff) << (8 * b);
                    }
                    o[w] = packed;
                }
                ov = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                float zf[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    zf[i] = 0.f;
                if (p.has_z)
                    unpack<KIND>(zv, zf);
                float r[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    r[i] = act_f(acc[t][i] + bias[i] + zf[i], p.nonline);
                if constexpr (KIND == B200CB_DT_BF16)
                    ov = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]),
                                    pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
                else if constexpr (KIND == B200CB_DT_F16)
                    ov = make_uint4(pack_f16x2(r[0], r[1]), pack_f16x2(r[2], r[3]),
                                    pack_f16x2(r[4], r[5]), pack_f16x2(r[6], r[7]));
                else
                    ov = make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]),
                                    __float_as_uint(r[2]), __float_as_uint(r[3]));
            }
            stg_stream_16(dst + ((size_t)n * p.dst_bstride + pix) * T::ES, ov);
        }
    }
}

template <int KIND, int F, int S, int TW>
cudaError_t launch_dw(const DwParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
    auto kern = dwconv_nhwc_kernel<KIND, F, S, TW>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess)
            return e;
    }
    kern<<<grid, 256, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int KIND, int TW>
cudaError_t dispatch_dw(const DwParams& p, int f, int s, dim3 grid, size_t smem, cudaStream_t st) {
    if (f == 3 && s == 1)
        return launch_dw<KIND, 3, 1, TW>(p, grid, smem, st);
    if (f == 3 && s == 2)
        return launch_dw<KIND, 3, 2, TW>(p, grid, smem, st);
    if (f == 5 && s == 1)
        return launch_dw<KIND, 5, 1, TW>(p, grid, smem, st);
    if (f == 5 && s == 2)
        return launch_dw<KIND, 5, 2, TW>(p, grid, smem, st);
    return launch_dw<KIND, 0, 1, 1>(p, grid, smem, st);
}

uint32_t dw_vec(uint32_t dt) {
    return 16 / dtype_size(dt);
}
// outputs per thread along W
uint32_t dw_tw(uint32_t dt, bool specialised) {
    if (!specialised)
        return 1;
    return dt == B200CB_DT_QS8 ? 2 : 4;
}
bool dw_specialised(const b200cb_desc& d) {
    return d.fh == d.fw && (d.fh == 3 || d.fh == 5) && d.stride_h == d.stride_w &&
           (d.stride_h == 1 || d.stride_h == 2) && d.dilate_h == 1 && d.dilate_w == 1;
}

}  // namespace

int dwconv_query(const b200cb_desc& d, b200cb_plan* plan) {
    if (d.format != B200CB_FORMAT_NHWC)
        return B200CB_UNSUPPORTED_FORMAT;
    if (d.sparse != B200CB_SPARSE_GROUP || d.group != d.ic || d.ic != d.oc)
        return B200CB_UNSUPPORTED_PARAM;  // depthwise only: icpg == ocpg == 1
    const uint32_t dt = d.src_dtype;
    if (dt != B200CB_DT_BF16 && dt != B200CB_DT_F16 && dt != B200CB_DT_F32 && dt != B200CB_DT_QS8)
        return B200CB_UNSUPPORTED_DTYPE;
    if (d.filter_dtype != dt || d.dst_dtype != dt)
        return B200CB_UNSUPPORTED_DTYPE;  // QS8 -> QS32 depthwise is left to the reference algos
    if (d.has_z && d.z_dtype != dt)
        return B200CB_UNSUPPORTED_DTYPE;
    if (d.bias_mode == B200CB_BIAS_FULL)
        return B200CB_UNSUPPORTED_PARAM;
    if (d.bias_mode == B200CB_BIAS_CHANNEL &&
        d.bias_dtype != (dt == B200CB_DT_QS8 ? (uint32_t)B200CB_DT_QS32 : dt))
        return B200CB_UNSUPPORTED_DTYPE;
    if (dt == B200CB_DT_QS8 && d.nonline_mode == B200CB_NL_SIGMOID)
        return B200CB_UNSUPPORTED_PARAM;
    const uint32_t vec = dw_vec(dt);
    if (d.ic % vec != 0)
        return B200CB_UNSUPPORTED_ALIGNMENT;
    const uint32_t es = dtype_size(dt);
    if ((d.src_batch_stride * es) % 16 || (d.dst_batch_stride * es) % 16 || (d.z_batch_stride * es) % 16)
        return B200CB_UNSUPPORTED_ALIGNMENT;
    if (d.fh * d.fw > 1024)
        return B200CB_UNSUPPORTED_SHAPE;

    const uint32_t cv = d.ic / vec;
    uint32_t cvb = 1;
    for (uint32_t c = 1; c <= 32 && c <= cv; ++c)
        if (cv % c == 0)
            cvb = c;
    const uint32_t smem = d.fh * d.fw * cvb * vec * 4;
    if (smem > MAX_DYN_SMEM)
        return B200CB_UNSUPPORTED_SHAPE;
    plan->magic = PLAN_MAGIC;
    plan->kernel = KERNEL_DWCONV;
    plan->block_n = cvb;
    plan->stages = 0;
    plan->grid = 0;
    plan->smem_bytes = smem;
    plan->a_mode = 0;
    for (auto& r : plan->reserved)
        r = 0;
    plan->workspace_bytes = 0;
    plan->desc = d;
    return B200CB_OK;
}

int dwconv_launch(
        const b200cb_plan& plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, cudaStream_t stream) {
    const b200cb_desc& d = plan.desc;
    if (!src || !filter || !dst || (d.bias_mode != B200CB_BIAS_NONE && !bias) || (d.has_z && !z))
        return B200CB_ERR_NULL;
    const bool spec = dw_specialised(d);
    const uint32_t vec = dw_vec(d.src_dtype);
    const uint32_t tw = dw_tw(d.src_dtype, spec);
    DwParams p;
    memset(&p, 0, sizeof(p));
    p.src = src, p.filter = filter, p.bias = bias, p.z = z, p.dst = dst;
    p.N = d.n, p.C = d.ic, p.IH = d.ih, p.IW = d.iw, p.OH = d.oh, p.OW = d.ow;
    p.FH = d.fh, p.FW = d.fw;
    p.stride_h = d.stride_h, p.stride_w = d.stride_w, p.pad_h = d.pad_h, p.pad_w = d.pad_w;
    p.dil_h = d.dilate_h, p.dil_w = d.dilate_w;
    p.CV = d.ic / vec;
    p.CVB = plan.block_n;
    p.PGB = 256 / p.CVB;
    p.OWG = (d.ow + tw - 1) / tw;
    p.total_pg = (uint64_t)d.n * d.oh * p.OWG;
    p.src_bstride = d.src_batch_stride ? d.src_batch_stride : (uint64_t)d.ih * d.iw * d.ic;
    p.dst_bstride = d.dst_batch_stride ? d.dst_batch_stride : (uint64_t)d.oh * d.ow * d.oc;
    p.z_bstride = d.z_batch_stride ? d.z_batch_stride : (uint64_t)d.oh * d.ow * d.oc;
    p.flip = d.mode == B200CB_MODE_CONVOLUTION;
    p.has_bias = d.bias_mode != B200CB_BIAS_NONE;
    p.has_z = d.has_z;
    p.nonline = d.nonline_mode;
    p.rq = make_requant(d);

    const uint32_t slices = p.CV / p.CVB;
    const uint64_t iters = (p.total_pg + p.PGB - 1) / p.PGB;
    // persistent-style grid: ~8 resident blocks of 256 threads per SM, split over the slices
    uint64_t gx = ((uint64_t)device_sm_count() * 8 + slices - 1) / slices;
    if (gx > iters)
        gx = iters;
    if (gx < 1)
        gx = 1;
    dim3 grid((uint32_t)gx, slices, 1);
    const int f = spec ? (int)d.fh : 0, s = spec ? (int)d.stride_h : 1;
    cudaError_t e;
    switch (d.src_dtype) {
        case B200CB_DT_BF16: e = dispatch_dw<B200CB_DT_BF16, 4>(p, f, s, grid, plan.smem_bytes, stream); break;
        case B200CB_DT_F16: e = dispatch_dw<B200CB_DT_F16, 4>(p, f, s, grid, plan.smem_bytes, stream); break;
        case B200CB_DT_F32: e = dispatch_dw<B200CB_DT_F32, 4>(p, f, s, grid, plan.smem_bytes, stream); break;
        case B200CB_DT_QS8: e = dispatch_dw<B200CB_DT_QS8, 2>(p, f, s, grid, plan.smem_bytes, stream); break;
        default: e = cudaErrorInvalidValue; break;
    }
    if (e != cudaSuccess) {
        set_last_error("dwconv launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

}  // namespace b200cb

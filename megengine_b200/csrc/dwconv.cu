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
#include <type_traits>
#include "common.h"
#include "epilogue.cuh"
#include "ptx.cuh"

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
                // IDENTITY / RELU with ordinary scales: the packed ALU / FMA-pipe requantize of the
                // implicit-GEMM epilogue (epilogue.cuh; no conversion-pipe instructions), else the
                // generic sequence
                const bool fast = p.rq.fast_div == 2 &&
                                  (p.nonline == B200CB_NL_RELU || p.nonline == B200CB_NL_IDENTITY);
                if (fast) {
                    const RequantPacked kp = make_requant_packed(p.rq);
                    const bool relu = p.nonline == B200CB_NL_RELU;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t zx = zw[w] ^ 0x80808080u;
                        uint32_t ub[4];
#pragma unroll
                        for (int b = 0; b < 4; b += 2) {
                            const int32_t s0 = acc[t][4 * w + b] + bias[4 * w + b];
                            const int32_t s1 = acc[t][4 * w + b + 1] + bias[4 * w + b + 1];
                            const uint32_t z0 = __byte_perm(zx, 0x4B000000u, 0x7540u | b);
                            const uint32_t z1 = __byte_perm(zx, 0x4B000000u, 0x7540u | (b + 1));
                            if (p.has_z) {
                                if (relu)
                                    requant_pair_nonneg<true>(s0, s1, z0, z1, kp, ub[b], ub[b + 1]);
                                else
                                    requant_pair_signed<true>(s0, s1, z0, z1, kp, ub[b], ub[b + 1]);
                            } else {
                                if (relu)
                                    requant_pair_nonneg<false>(s0, s1, z0, z1, kp, ub[b], ub[b + 1]);
                                else
                                    requant_pair_signed<false>(s0, s1, z0, z1, kp, ub[b], ub[b + 1]);
                            }
                        }
                        o[w] = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
                    }
                } else {
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

// ------------------------------------------------------------------------------------------------
// 3x3 depthwise, 16-bit floats: the bandwidth kernels (MobileNetV2 / ShuffleNetV2).
//
//   thread   = one channel PAIR (a 32-bit word of the NHWC row) x TW adjacent output columns x a
//              vertical span of output rows;
//   warp     = consecutive channel pairs: every load / store instruction touches whole 128-byte lines;
//   window   = the three input rows of the current output row live UNPACKED (fp32 pairs) in registers
//              and slide down: each output row fetches only its new input row(s), an input element is
//              unpacked once per span;
//   math     = packed FFMA2 (two channels per instruction), the pair's 3x3 weights in registers;
//   epilogue = bias (accumulator init) + z + activation fused, one pass over HBM.
// Two front ends feed the same window code:
//   dw3x3_tma_kernel  (C % 8 == 0)  input tiles (rows x cols x 64 channels, halo and zero padding
//                     included) are staged in shared memory by TMA, 3 tiles in flight per SM, so
//                     memory-level parallelism does not depend on registers or occupancy;
//   dw3x3_kernel      (any even C)  rows are read straight from global memory (ld.global.nc).
struct Dw3Common {
    const void* filter;
    const void* bias;
    const uint32_t* z;
    uint32_t* dst;
    uint32_t N, C, IH, IW, OH, OW, CP;
    int32_t pad_h, pad_w;
    uint64_t z_bwords, dst_bwords;  // batch strides in 32-bit words
    uint32_t flip, has_bias, has_z, nonline;
};

// Float32 rides the same kernels: one fp32 channel is one 32-bit word, i.e. the "pair" (c, c + 1) of 16-bit
// half-channels of a tensor with 2 C half-channels; the second lane of every packed operation is a zero
// (the kernels are HBM bound: instructions per byte are the same as for the 16-bit kinds).
template <int KIND>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
    if constexpr (KIND == B200CB_DT_BF16)
        return make_float2(bf16lo(w), bf16hi(w));
    else if constexpr (KIND == B200CB_DT_F32)
        return make_float2(__uint_as_float(w), 0.f);
    else
        return make_float2(f16lo(w), f16hi(w));
}
template <int KIND>
__device__ __forceinline__ uint32_t pack2(float2 v) {
    if constexpr (KIND == B200CB_DT_BF16)
        return pack_bf16x2(v.x, v.y);
    else if constexpr (KIND == B200CB_DT_F32)
        return __float_as_uint(v.x);
    else
        return pack_f16x2(v.x, v.y);
}
template <int KIND>
__device__ __forceinline__ float elem_f32(const void* p, size_t idx) {
    if constexpr (KIND == B200CB_DT_F32)
        return reinterpret_cast<const float*>(p)[idx];
    else if constexpr (KIND == B200CB_DT_BF16)
        return __uint_as_float(static_cast<uint32_t>(reinterpret_cast<const unsigned short*>(p)[idx]) << 16);
    else
        return __half2float(reinterpret_cast<const __half*>(p)[idx]);
}
template <int KIND>
__device__ __forceinline__ void load_weights(const Dw3Common& p, uint32_t c, float2 (&w)[9], float2& bias) {
    if constexpr (KIND == B200CB_DT_F32) {
        const uint32_t ch = c / 2;  // half-channel index -> fp32 channel
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
            w[tap] = make_float2(reinterpret_cast<const float*>(p.filter)[(size_t)ch * 9 + (p.flip ? 8 - tap : tap)], 0.f);
        bias = make_float2(p.has_bias ? reinterpret_cast<const float*>(p.bias)[ch] : 0.f, 0.f);
        return;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int ft = p.flip ? 8 - tap : tap;
        w[tap] = make_float2(elem_f32<KIND>(p.filter, (size_t)c * 9 + ft), elem_f32<KIND>(p.filter, (size_t)(c + 1) * 9 + ft));
    }
    bias = p.has_bias ? make_float2(elem_f32<KIND>(p.bias, c), elem_f32<KIND>(p.bias, c + 1)) : make_float2(0.f, 0.f);
}

// Output rows [oh0, oh1) x columns [ow0, ow0 + TW) of image n for channel pair cp.  `ld(r, raw)` fetches
// the IN packed words of input row r, r counted from the top tap row of output row oh0.
template <int KIND, int S, int TW, class Loader>
__device__ __forceinline__ void dw3_slide(
        const Dw3Common& p, const Loader& ld, uint32_t n, uint32_t cp, uint32_t oh0, uint32_t oh1,
        uint32_t ow0, const float2 (&w)[9], float2 bias, bool store_ok) {
    constexpr int IN = (TW - 1) * S + 3;
    constexpr int NEW = S;  // new input rows per output row
    const uint32_t* zp = p.z + n * p.z_bwords + cp;
    uint32_t* dst = p.dst + n * p.dst_bwords + cp;
    const float act_lo = p.nonline == B200CB_NL_RELU ? 0.f : -INFINITY;
    const bool generic_act = p.nonline == B200CB_NL_H_SWISH || p.nonline == B200CB_NL_SIGMOID;
    float2 buf[3][IN];
    uint32_t raw[NEW][IN];
    {
        uint32_t r0[IN];
        ld(0, r0);
#pragma unroll
        for (int j = 0; j < IN; ++j)
            buf[0][j] = unpack2<KIND>(r0[j]);
        if constexpr (S == 1) {
            ld(1, r0);
#pragma unroll
            for (int j = 0; j < IN; ++j)
                buf[1][j] = unpack2<KIND>(r0[j]);
        }
    }
    auto fetch = [&](uint32_t oh) {
        const int32_t r = static_cast<int32_t>(oh - oh0) * S + 3 - NEW;
#pragma unroll
        for (int k = 0; k < NEW; ++k)
            ld(r + k, raw[k]);
    };
    fetch(oh0);
    // one output row; T / M / B = which buffer holds the top / middle / bottom input row
    auto step = [&](uint32_t oh, auto TT, auto MM, auto BB) {
        constexpr int T = decltype(TT)::value, M = decltype(MM)::value, B = decltype(BB)::value;
#pragma unroll
        for (int j = 0; j < IN; ++j) {
            if constexpr (S == 1) {
                buf[B][j] = unpack2<KIND>(raw[0][j]);
            } else {
                buf[M][j] = unpack2<KIND>(raw[0][j]);
                buf[B][j] = unpack2<KIND>(raw[1][j]);
            }
        }
        uint32_t zraw[TW];
        const size_t orow = (static_cast<size_t>(oh) * p.OW + ow0) * p.CP;
        if (p.has_z) {
#pragma unroll
            for (int t2 = 0; t2 < TW; ++t2)
                zraw[t2] = (store_ok && ow0 + t2 < p.OW) ? __ldg(zp + orow + static_cast<size_t>(t2) * p.CP) : 0u;
        }
        if (oh + 1 < oh1)
            fetch(oh + 1);  // next row's loads fly while this row is computed
#pragma unroll
        for (int t2 = 0; t2 < TW; ++t2) {
            float2 acc = bias;
#pragma unroll
            for (int fw = 0; fw < 3; ++fw) {
                acc = __ffma2_rn(buf[T][t2 * S + fw], w[fw], acc);
                acc = __ffma2_rn(buf[M][t2 * S + fw], w[3 + fw], acc);
                acc = __ffma2_rn(buf[B][t2 * S + fw], w[6 + fw], acc);
            }
            if (p.has_z) {
                const float2 zf = unpack2<KIND>(zraw[t2]);
                acc.x += zf.x, acc.y += zf.y;
            }
            if (generic_act) {
                acc.x = act_generic(acc.x, p.nonline), acc.y = act_generic(acc.y, p.nonline);
            } else {
                acc.x = acc.x <= act_lo ? act_lo : acc.x;
                acc.y = acc.y <= act_lo ? act_lo : acc.y;
            }
            if (store_ok && ow0 + t2 < p.OW)
                dst[orow + static_cast<size_t>(t2) * p.CP] =
                        pack2<KIND>(acc);
        }
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    uint32_t oh = oh0;
    while (true) {
        step(oh, I0{}, I1{}, I2{});
        if (++oh >= oh1) break;
        if constexpr (S == 1) {
            step(oh, I1{}, I2{}, I0{});
            if (++oh >= oh1) break;
            step(oh, I2{}, I0{}, I1{});
        } else {
            step(oh, I2{}, I0{}, I1{});
            if (++oh >= oh1) break;
            step(oh, I1{}, I2{}, I0{});
        }
        if (++oh >= oh1) break;
    }
}

// ---------------------------------------------------------------- front end 1: rows from global memory
struct Dw3Params {
    Dw3Common c;
    const uint32_t* src;
    uint32_t strips, spans, RH;
    uint64_t total_items;
    uint64_t src_bwords;
};

template <int KIND, int S, int TW>
__global__ void __launch_bounds__(256, 2) dw3x3_kernel(const Dw3Params q) {
    constexpr int IN = (TW - 1) * S + 3;
    const Dw3Common& p = q.c;
    const uint64_t item = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (item >= q.total_items)
        return;
    const uint32_t cp = static_cast<uint32_t>(item % p.CP);
    uint64_t t = item / p.CP;
    const uint32_t strip = static_cast<uint32_t>(t % q.strips);
    t /= q.strips;
    const uint32_t span = static_cast<uint32_t>(t % q.spans);
    const uint32_t n = static_cast<uint32_t>(t / q.spans);
    const uint32_t ow0 = strip * TW, oh0 = span * q.RH;
    const uint32_t oh1 = oh0 + q.RH < p.OH ? oh0 + q.RH : p.OH;
    float2 w[9], bias;
    load_weights<KIND>(p, cp * 2, w, bias);
    const uint32_t* src = q.src + n * q.src_bwords + cp;
    const int32_t iw0 = static_cast<int32_t>(ow0) * S - p.pad_w;
    const int32_t ih0 = static_cast<int32_t>(oh0) * S - p.pad_h;
    uint32_t colmask = 0;
#pragma unroll
    for (int j = 0; j < IN; ++j)
        colmask |= (static_cast<uint32_t>(iw0 + j) < p.IW ? 1u : 0u) << j;
    auto ld = [&](int32_t r, uint32_t (&raw)[IN]) {
        const int32_t ih = ih0 + r;
        const bool rv = static_cast<uint32_t>(ih) < p.IH;
        const uint32_t* row = src + (static_cast<int64_t>(ih) * p.IW + iw0) * static_cast<int64_t>(p.CP);
#pragma unroll
        for (int j = 0; j < IN; ++j)
            raw[j] = (rv && ((colmask >> j) & 1)) ? __ldg(row + static_cast<size_t>(j) * p.CP) : 0u;
    };
    dw3_slide<KIND, S, TW>(p, ld, n, cp, oh0, oh1, ow0, w, bias, true);
}

// ---------------------------------------------------------------- front end 2: TMA-staged input tiles
// Persistent CTAs walk the tiles with the channel block as the FASTEST tile index and a grid that is a
// multiple of the channel-block count: a CTA keeps one channel block for its whole life (weights are
// loaded once) and the channel blocks of one spatial tile run at the same time on neighbouring SMs, so
// every DRAM line of the NHWC rows is consumed while it is still in L2.  Warp 15 is the TMA producer
// (ring of `stages` tiles, full / empty mbarriers); each compute warp owns one unit = (column strip x
// row span) of the tile.  The tile is [IHT][IWT][CB] in shared memory: a warp reading one pixel's
// channel pairs hits consecutive banks, and all tile offsets are immediates or one register.
struct Dw3TmaParams {
    Dw3Common c;
    CUtensorMap tmS;       // src (C, IW, IH, N), box (CB, IWT, IHT, NB), zero fill = conv padding
    uint32_t TH, TWT;      // output rows / columns per tile
    uint32_t IHT, IWT;     // input rows / columns per tile
    uint32_t tiles_h, tiles_w, cblks;
    uint32_t spatial_tiles;      // N * tiles_h * tiles_w
    uint32_t RS, strips, units;  // rows per unit, column strips per tile, units per tile (<= 15)
    uint32_t NB, spans, N;       // images per tile (small images: whole images side by side), row spans
    uint32_t stage_bytes, tile_bytes, stages;  // ring stride (128-byte aligned) / bytes one TMA box delivers
};
constexpr uint32_t DW3_COMPUTE_WARPS = 15;  // + 1 producer warp = 16 warps: 128 registers per thread

template <int KIND, int S, int TW, int CPB>
__global__ void __launch_bounds__((DW3_COMPUTE_WARPS + 1) * 32, 1)
        dw3x3_tma_kernel(const __grid_constant__ Dw3TmaParams q) {
    constexpr int IN = (TW - 1) * S + 3;
    constexpr uint32_t PIX = CPB * 4;  // bytes per pixel in the tile
    constexpr uint32_t CB = CPB * 2;
    using namespace ptx;
    const Dw3Common& p = q.c;
    extern __shared__ __align__(128) uint8_t dsm[];
    const uint32_t sbase = smem_u32(dsm);
    const uint32_t bar_full = sbase, bar_empty = sbase + 64;  // [8] each
    const uint32_t stage0 = sbase + 128;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tensormap(&q.tmS);
        for (uint32_t i = 0; i < q.stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, DW3_COMPUTE_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // programmatic dependent launch: the next kernel's CTAs may take this SM as soon as it is free; this
    // one touches activations (src / z / dst) only after the kernel before it has completed
    ptx::pdl_launch_dependents();
    // gridDim.x is a multiple of cblks: this CTA's channel block never changes
    const uint32_t cblk = blockIdx.x % q.cblks;
    const uint32_t sp0 = blockIdx.x / q.cblks, sp_step = gridDim.x / q.cblks;
    if (warp == DW3_COMPUTE_WARPS) {
        pdl_wait();
        uint32_t st = 0, ph = 0;
        for (uint32_t sp = sp0; sp < q.spatial_tiles; sp += sp_step) {
            const uint32_t tw = sp % q.tiles_w, r = sp / q.tiles_w;
            const uint32_t th = r % q.tiles_h, n = r / q.tiles_h;
            mbar_wait(bar_empty + 8 * st, ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8 * st, q.tile_bytes);
                tma_load_4d(
                        stage0 + st * q.stage_bytes, &q.tmS, bar_full + 8 * st,
                        static_cast<int32_t>(cblk * CB), static_cast<int32_t>(tw * q.TWT) * S - p.pad_w,
                        static_cast<int32_t>(th * q.TH) * S - p.pad_h, static_cast<int32_t>(n * q.NB));
            }
            __syncwarp();
            if (++st == q.stages) {
                st = 0;
                ph ^= 1;
            }
        }
        return;
    }
    // ---- compute warps: lane = channel pair (and, when CPB == 16, one of two adjacent units)
    const uint32_t cpl = lane % CPB, sub = lane / CPB;
    constexpr uint32_t SUBS = 32 / CPB;
    const uint32_t c = cblk * CB + cpl * 2;
    const bool ch_ok = c < p.C;
    float2 w[9], bias;
    // (filter / bias are stream-ordered inputs like any other: a previous kernel of the stream may have
    // written them -- dynamic filters, on-device weight updates -- so they are read after the dependency wait)
    ptx::pdl_wait();
    load_weights<KIND>(p, ch_ok ? c : 0, w, bias);
    const uint32_t u = warp * SUBS + sub;          // this thread's unit in every tile
    const bool has_unit = u < q.units;
    const uint32_t strip = u % q.strips, span = (u / q.strips) % q.spans, img = u / (q.strips * q.spans);
    const uint32_t srow = q.IWT * PIX;                                  // bytes per tile row
    const uint32_t unit_off = ((img * q.IHT + span * q.RS * S) * q.IWT + strip * TW * S) * PIX + cpl * 4;
    const float act_lo = p.nonline == B200CB_NL_RELU ? 0.f : -INFINITY;
    const bool generic_act = p.nonline == B200CB_NL_H_SWISH || p.nonline == B200CB_NL_SIGMOID;
    const uint32_t cstride = p.CP;                 // words between adjacent output pixels
    const uint32_t rstride = p.OW * p.CP;          // words between output rows
    uint32_t st = 0, ph = 0;
    for (uint32_t sp = sp0; sp < q.spatial_tiles; sp += sp_step) {
        const uint32_t tw = sp % q.tiles_w, r = sp / q.tiles_w;
        const uint32_t th = r % q.tiles_h, n = (r / q.tiles_h) * q.NB + img;
        mbar_wait(bar_full + 8 * st, ph);
        const uint32_t oh0 = th * q.TH + span * q.RS;
        uint32_t oh1 = oh0 + q.RS;
        if (oh1 > (th + 1) * q.TH)
            oh1 = (th + 1) * q.TH;
        if (oh1 > p.OH)
            oh1 = p.OH;
        const uint32_t ow0 = tw * q.TWT + strip * TW;
        if (has_unit && oh0 < oh1 && ow0 < p.OW && n < q.N) {
            // next tile row to read (generic pointer into the staged tile; offsets fold into the LDS)
            const uint8_t* a = dsm + 128 + st * q.stage_bytes + unit_off;
            uint32_t cmask = 0;                                    // columns this thread stores
#pragma unroll
            for (int t2 = 0; t2 < TW; ++t2)
                cmask |= (ch_ok && ow0 + t2 < p.OW ? 1u : 0u) << t2;
            uint32_t* drow = p.dst + n * p.dst_bwords + (static_cast<size_t>(oh0) * p.OW + ow0) * p.CP + c / 2;
            const uint32_t* zrow = p.z + n * p.z_bwords + (static_cast<size_t>(oh0) * p.OW + ow0) * p.CP + c / 2;
            float2 buf[3][IN];
            auto lds_row = [&](float2 (&dstrow)[IN]) {
                uint32_t raw[IN];
#pragma unroll
                for (int j = 0; j < IN; ++j)
                    raw[j] = *reinterpret_cast<const uint32_t*>(a + j * PIX);
#pragma unroll
                for (int j = 0; j < IN; ++j)
                    dstrow[j] = unpack2<KIND>(raw[j]);
                a += srow;
            };
            lds_row(buf[0]);
            if constexpr (S == 1)
                lds_row(buf[1]);
            auto step = [&](auto TT, auto MM, auto BB) {
                constexpr int T = decltype(TT)::value, M = decltype(MM)::value, B = decltype(BB)::value;
                if constexpr (S == 2)
                    lds_row(buf[M]);
                lds_row(buf[B]);
                float2 acc[TW];
#pragma unroll
                for (int t2 = 0; t2 < TW; ++t2) {
                    acc[t2] = bias;
#pragma unroll
                    for (int fw = 0; fw < 3; ++fw) {
                        acc[t2] = __ffma2_rn(buf[T][t2 * S + fw], w[fw], acc[t2]);
                        acc[t2] = __ffma2_rn(buf[M][t2 * S + fw], w[3 + fw], acc[t2]);
                        acc[t2] = __ffma2_rn(buf[B][t2 * S + fw], w[6 + fw], acc[t2]);
                    }
                }
                if (p.has_z) {
#pragma unroll
                    for (int t2 = 0; t2 < TW; ++t2)
                        if ((cmask >> t2) & 1) {
                            const float2 zf = unpack2<KIND>(__ldg(zrow + t2 * cstride));
                            acc[t2].x += zf.x, acc[t2].y += zf.y;
                        }
                    zrow += rstride;
                }
                if (generic_act) {
#pragma unroll
                    for (int t2 = 0; t2 < TW; ++t2)
                        acc[t2].x = act_generic(acc[t2].x, p.nonline), acc[t2].y = act_generic(acc[t2].y, p.nonline);
                } else {
#pragma unroll
                    for (int t2 = 0; t2 < TW; ++t2)
                        acc[t2].x = fmaxf(acc[t2].x, act_lo), acc[t2].y = fmaxf(acc[t2].y, act_lo);
                }
#pragma unroll
                for (int t2 = 0; t2 < TW; ++t2)
                    if ((cmask >> t2) & 1)
                        drow[t2 * cstride] = pack2<KIND>(acc[t2]);
                drow += rstride;
            };
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            using I2 = std::integral_constant<int, 2>;
            uint32_t rows = oh1 - oh0;
            while (true) {
                step(I0{}, I1{}, I2{});
                if (--rows == 0) break;
                if constexpr (S == 1) {
                    step(I1{}, I2{}, I0{});
                    if (--rows == 0) break;
                    step(I2{}, I0{}, I1{});
                } else {
                    step(I2{}, I0{}, I1{});
                    if (--rows == 0) break;
                    step(I1{}, I2{}, I0{});
                }
                if (--rows == 0) break;
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(bar_empty + 8 * st);
        if (++st == q.stages) {
            st = 0;
            ph ^= 1;
        }
    }
}

// ---------------------------------------------------------------- front end 3: bulk-copied row blocks
// Channel counts whose pixels are not 16-byte multiples (ShuffleNetV2's 58 / 116) cannot go through a
// tensor map, but full-width input rows of an NHWC image are ONE contiguous byte range, and it is
// 16-byte aligned whenever IW * C * 2 is: a persistent CTA stages the rows of a tile (halo included,
// clipped to the image) with a single cp.async.bulk per tile into a ring of stages, and the same window
// code reads them with the padding masks of the global-memory front end.  Thread = (channel pair, unit):
// the channel pair is fixed for the thread's life (weights in registers), units = column strip x row
// span, walked in passes of floor(480 / CP) units.
struct Dw3BulkParams {
    Dw3Common c;
    const uint8_t* src;
    uint64_t src_bbytes;
    uint32_t row_bytes;          // IW * C * 2: what one bulk copy delivers
    uint32_t row_stride;         // (IW + 4) * C * 2: staged row = 4 zero pixels + the image row
    uint32_t TH, RS, IHT;        // output rows per tile / per unit, input rows per tile
    uint32_t strips, spans, U;   // units per tile = strips * spans, walked U at a time
    uint32_t tiles_h, total_tiles;
    uint32_t zrow_bytes, stage_bytes, stages, smem_bytes;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Shared memory: [barriers 128 B][zero row][stages x IHT staged rows][4 zero pixels].  A staged row is
// 4 zero pixels followed by the IW image pixels: the left padding column of a row is the pixel before
// it, the right padding columns are the zero pixels of the next row, so the window code needs no column
// masks; tap rows above / below the image read the zero row instead.
template <int KIND, int S, int TW>
__global__ void __launch_bounds__((DW3_COMPUTE_WARPS + 1) * 32, 1) dw3x3_bulk_kernel(const Dw3BulkParams q) {
    constexpr int IN = (TW - 1) * S + 3;
    using namespace ptx;
    const Dw3Common& p = q.c;
    extern __shared__ __align__(128) uint8_t dsm[];
    const uint32_t sbase = smem_u32(dsm);
    const uint32_t bar_full = sbase, bar_empty = sbase + 64;
    const uint32_t zrow = sbase + 128, stage0 = zrow + q.zrow_bytes;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pix = p.CP * 4;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < q.stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, DW3_COMPUTE_WARPS);
        }
        fence_mbar_init();
    }
    // zero row, the 4 leading pixels of every staged row, the 4 trailing pixels (never written by copies)
    for (uint32_t i = threadIdx.x * 4; i < q.zrow_bytes; i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(dsm + 128 + i) = 0u;
    {
        const uint32_t rows = q.stages * q.IHT + 1, padw = 4 * p.CP;  // words per pad
        for (uint32_t i = threadIdx.x; i < rows * padw; i += blockDim.x) {
            const uint32_t r = i / padw, wd = i % padw;
            *reinterpret_cast<uint32_t*>(dsm + 128 + q.zrow_bytes + static_cast<size_t>(r) * q.row_stride + wd * 4) = 0u;
        }
    }
    __syncthreads();
    ptx::pdl_launch_dependents();
    if (warp == DW3_COMPUTE_WARPS) {
        pdl_wait();
        uint32_t st = 0, ph = 0;
        for (uint32_t t = blockIdx.x; t < q.total_tiles; t += gridDim.x) {
            const uint32_t th = t % q.tiles_h, n = t / q.tiles_h;
            const int32_t ih_lo = static_cast<int32_t>(th * q.TH) * S - p.pad_h;
            const int32_t lo = ih_lo < 0 ? 0 : ih_lo;
            int32_t hi = ih_lo + static_cast<int32_t>(q.IHT);
            if (hi > static_cast<int32_t>(p.IH))
                hi = static_cast<int32_t>(p.IH);
            mbar_wait(bar_empty + 8 * st, ph ^ 1);
            if (lane == 0)
                mbar_arrive_expect_tx(bar_full + 8 * st, static_cast<uint32_t>(hi - lo) * q.row_bytes);
            __syncwarp();
            const uint8_t* img = q.src + n * q.src_bbytes;
            for (int32_t ih = lo + static_cast<int32_t>(lane); ih < hi; ih += 32)
                bulk_load_1d(
                        stage0 + st * q.stage_bytes + static_cast<uint32_t>(ih - ih_lo) * q.row_stride + 4 * pix,
                        img + static_cast<uint64_t>(ih) * q.row_bytes, q.row_bytes, bar_full + 8 * st);
            __syncwarp();
            if (++st == q.stages) {
                st = 0;
                ph ^= 1;
            }
        }
        return;
    }
    const uint32_t cp = threadIdx.x % p.CP, u0 = threadIdx.x / p.CP;
    const bool active = u0 < q.U;
    float2 w[9], bias;
    ptx::pdl_wait();  // (weights are stream-ordered inputs too: read after the dependency wait)
    load_weights<KIND>(p, cp * 2, w, bias);
    const uint32_t units = q.strips * q.spans;
    const float act_lo = p.nonline == B200CB_NL_RELU ? 0.f : -INFINITY;
    const bool generic_act = p.nonline == B200CB_NL_H_SWISH || p.nonline == B200CB_NL_SIGMOID;
    const uint32_t cstride = p.CP, rstride = p.OW * p.CP;
    uint32_t st = 0, ph = 0;
    for (uint32_t t = blockIdx.x; t < q.total_tiles; t += gridDim.x) {
        const uint32_t th = t % q.tiles_h, n = t / q.tiles_h;
        const int32_t ih_lo = static_cast<int32_t>(th * q.TH) * S - p.pad_h;
        mbar_wait(bar_full + 8 * st, ph);
        if (active) {
            for (uint32_t u = u0; u < units; u += q.U) {
                const uint32_t strip = u % q.strips, span = u / q.strips;
                const uint32_t oh0 = th * q.TH + span * q.RS;
                uint32_t oh1 = oh0 + q.RS;
                if (oh1 > (th + 1) * q.TH)
                    oh1 = (th + 1) * q.TH;
                if (oh1 > p.OH)
                    oh1 = p.OH;
                const uint32_t ow0 = strip * TW;
                if (oh0 >= oh1)
                    continue;
                // column offset of the unit's first tap inside a staged row (4 zero pixels lead the row)
                const uint32_t coff = (4u + ow0 * S - static_cast<uint32_t>(p.pad_w)) * pix + cp * 4;
                uint32_t a = stage0 + st * q.stage_bytes + span * q.RS * S * q.row_stride + coff;
                const uint32_t za = zrow + coff;
                int32_t ih = ih_lo + static_cast<int32_t>(span * q.RS) * S;  // image row of the next tap row
                uint32_t cmask = 0;
#pragma unroll
                for (int t2 = 0; t2 < TW; ++t2)
                    cmask |= (ow0 + t2 < p.OW ? 1u : 0u) << t2;
                uint32_t* drow = p.dst + n * p.dst_bwords + (static_cast<size_t>(oh0) * p.OW + ow0) * p.CP + cp;
                const uint32_t* zrowp = p.z + n * p.z_bwords + (static_cast<size_t>(oh0) * p.OW + ow0) * p.CP + cp;
                float2 buf[3][IN];
                auto lds_row = [&](float2 (&dstrow)[IN]) {
                    const uint32_t src_a = static_cast<uint32_t>(ih) < p.IH ? a : za;
                    uint32_t raw[IN];
#pragma unroll
                    for (int j = 0; j < IN; ++j)
                        raw[j] = lds_u32(src_a + j * pix);
#pragma unroll
                    for (int j = 0; j < IN; ++j)
                        dstrow[j] = unpack2<KIND>(raw[j]);
                    a += q.row_stride;
                    ++ih;
                };
                lds_row(buf[0]);
                if constexpr (S == 1)
                    lds_row(buf[1]);
                auto step = [&](auto TT, auto MM, auto BB) {
                    constexpr int T = decltype(TT)::value, M = decltype(MM)::value, B = decltype(BB)::value;
                    if constexpr (S == 2)
                        lds_row(buf[M]);
                    lds_row(buf[B]);
                    float2 acc[TW];
#pragma unroll
                    for (int t2 = 0; t2 < TW; ++t2) {
                        acc[t2] = bias;
#pragma unroll
                        for (int fw = 0; fw < 3; ++fw) {
                            acc[t2] = __ffma2_rn(buf[T][t2 * S + fw], w[fw], acc[t2]);
                            acc[t2] = __ffma2_rn(buf[M][t2 * S + fw], w[3 + fw], acc[t2]);
                            acc[t2] = __ffma2_rn(buf[B][t2 * S + fw], w[6 + fw], acc[t2]);
                        }
                    }
                    if (p.has_z) {
#pragma unroll
                        for (int t2 = 0; t2 < TW; ++t2)
                            if ((cmask >> t2) & 1) {
                                const float2 zf = unpack2<KIND>(__ldg(zrowp + t2 * cstride));
                                acc[t2].x += zf.x, acc[t2].y += zf.y;
                            }
                        zrowp += rstride;
                    }
                    if (generic_act) {
#pragma unroll
                        for (int t2 = 0; t2 < TW; ++t2)
                            acc[t2].x = act_generic(acc[t2].x, p.nonline), acc[t2].y = act_generic(acc[t2].y, p.nonline);
                    } else {
#pragma unroll
                        for (int t2 = 0; t2 < TW; ++t2)
                            acc[t2].x = fmaxf(acc[t2].x, act_lo), acc[t2].y = fmaxf(acc[t2].y, act_lo);
                    }
#pragma unroll
                    for (int t2 = 0; t2 < TW; ++t2)
                        if ((cmask >> t2) & 1)
                            drow[t2 * cstride] = pack2<KIND>(acc[t2]);
                    drow += rstride;
                };
                using I0 = std::integral_constant<int, 0>;
                using I1 = std::integral_constant<int, 1>;
                using I2 = std::integral_constant<int, 2>;
                uint32_t rows = oh1 - oh0;
                while (true) {
                    step(I0{}, I1{}, I2{});
                    if (--rows == 0) break;
                    if constexpr (S == 1) {
                        step(I1{}, I2{}, I0{});
                        if (--rows == 0) break;
                        step(I2{}, I0{}, I1{});
                    } else {
                        step(I2{}, I0{}, I1{});
                        if (--rows == 0) break;
                        step(I1{}, I2{}, I0{});
                    }
                    if (--rows == 0) break;
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(bar_empty + 8 * st);
        if (++st == q.stages) {
            st = 0;
            ph ^= 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 depthwise, QuantizedS8 (replaces dnn/src/cuda/conv_bias/chanwise_8x8x32.cpp + the TypeCvt / Elemwise
// passes behind it): a window kernel on 4-channel words.
//   thread   = one 32-bit word of the NHWC row (4 channels) x TW adjacent output columns x a vertical span;
//   warp     = consecutive channel words: loads / stores touch whole 128-byte lines;
//   math     = DP4A ACROSS TAPS: the 6 pixel words of an input row are byte-transposed (PRMT) into
//              per-channel words [p0 p1 p2 p3] / [p2 p3 p4 p5]; one dp4a with the filter row packed as
//              [w0 w1 w2 0] or [0 w0 w1 w2] is one output's whole filter row (3 MACs), so an output costs
//              3 dp4a instead of 9 multiply-adds + 9 byte extractions;
//   window   = the transposed rows slide down in registers (one new input row per output row at stride 1);
//   epilogue = bias (accumulator init), z, activation and the bit-exact requantize of epilogue.cuh fused.
struct DwS8Params {
    const uint32_t* src;
    const int8_t* filter;
    const int32_t* bias;
    const uint32_t* z;
    uint32_t* dst;
    uint32_t N, C, IH, IW, OH, OW, CW;  // CW = C / 4 words per pixel
    int32_t pad_h, pad_w;
    uint64_t src_bwords, z_bwords, dst_bwords;
    uint32_t strips, spans, RH;
    uint64_t total_items;
    uint32_t flip, has_bias, has_z, nonline;
    RequantParams rq;
};

__device__ __forceinline__ int32_t sx8(uint32_t w, int b) {
    return static_cast<int32_t>(static_cast<int8_t>((w >> (8 * b)) & 0xffu));
}

// MODE: 0 = generic requantize (any activation / scale), 1 = fast IDENTITY, 2 = fast RELU (epilogue.cuh);
// the variants are separate kernels so that the row loop holds exactly one epilogue (a run-time choice
// inside the unrolled loop quadrupled the code and ran out of the instruction cache)
template <int S, int MODE, bool HAS_Z>
__global__ void __launch_bounds__(256, 2) dw3x3_s8_kernel(const DwS8Params p) {
    constexpr int TW = S == 1 ? 4 : 2;
    const uint64_t item = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (item >= p.total_items)
        return;
    const uint32_t cw = static_cast<uint32_t>(item % p.CW);
    uint64_t t = item / p.CW;
    const uint32_t strip = static_cast<uint32_t>(t % p.strips);
    t /= p.strips;
    const uint32_t span = static_cast<uint32_t>(t % p.spans);
    const uint32_t n = static_cast<uint32_t>(t / p.spans);
    const uint32_t ow0 = strip * TW, oh0 = span * p.RH;
    const uint32_t oh1 = oh0 + p.RH < p.OH ? oh0 + p.RH : p.OH;
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();

    // filter rows packed for dp4a: wa = [w0 w1 w2 0] (bytes 0..3), wb = [0 w0 w1 w2].  The 4 channels' 36
    // filter bytes are 9 aligned words
    uint32_t wa[4][3], wb[4][3];
    int32_t bias[4];
    {
        uint32_t fw[9];
        const uint32_t* f32p = reinterpret_cast<const uint32_t*>(p.filter) + static_cast<size_t>(cw) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            fw[i] = __ldg(f32p + i);
        auto fbyte = [&](int idx) { return (fw[idx >> 2] >> (8 * (idx & 3))) & 0xffu; };
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                uint32_t w = 0;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    const int tap = r * 3 + x;
                    w |= (p.flip ? fbyte(k * 9 + 8 - tap) : fbyte(k * 9 + tap)) << (8 * x);
                }
                wa[k][r] = w, wb[k][r] = w << 8;
            }
            bias[k] = p.has_bias ? p.bias[cw * 4 + k] : 0;
        }
    }
    const uint32_t* img = p.src + n * p.src_bwords + cw;
    const int32_t iw0 = static_cast<int32_t>(ow0) * S - p.pad_w;
    // raw pixel words of one input row (the 6 this thread needs), and their byte transpose into per-channel
    // words A[k] = pixels 0..3, B[k] = pixels 2..5
    auto fetch_row = [&](int32_t ih, uint32_t (&px)[6]) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int32_t iw = iw0 + j;
            px[j] = (static_cast<uint32_t>(ih) < p.IH && static_cast<uint32_t>(iw) < p.IW)
                            ? __ldg(img + (static_cast<size_t>(ih) * p.IW + iw) * p.CW)
                            : 0u;
        }
    };
    auto transpose_row = [&](const uint32_t (&px)[6], uint32_t (&A)[4], uint32_t (&B)[4]) {
        const uint32_t l01 = __byte_perm(px[0], px[1], 0x5140), h01 = __byte_perm(px[0], px[1], 0x7362);
        const uint32_t l23 = __byte_perm(px[2], px[3], 0x5140), h23 = __byte_perm(px[2], px[3], 0x7362);
        const uint32_t l45 = __byte_perm(px[4], px[5], 0x5140), h45 = __byte_perm(px[4], px[5], 0x7362);
        A[0] = __byte_perm(l01, l23, 0x5410), A[1] = __byte_perm(l01, l23, 0x7632);
        A[2] = __byte_perm(h01, h23, 0x5410), A[3] = __byte_perm(h01, h23, 0x7632);
        B[0] = __byte_perm(l23, l45, 0x5410), B[1] = __byte_perm(l23, l45, 0x7632);
        B[2] = __byte_perm(h23, h45, 0x5410), B[3] = __byte_perm(h23, h45, 0x7632);
    };
    uint32_t A[3][4], B[3][4];
    uint32_t nx[S][6];  // the next output row's new input row(s), in flight while this row is computed
    const int32_t ih_top = static_cast<int32_t>(oh0) * S - p.pad_h;
    {
        uint32_t px[6];
        fetch_row(ih_top, px);
        transpose_row(px, A[0], B[0]);
        if (S == 1) {
            fetch_row(ih_top + 1, px);
            transpose_row(px, A[1], B[1]);
        }
    }
    // new rows of output row oh0: ih_top + 2 (stride 1) / ih_top + 1, + 2 (stride 2)
#pragma unroll
    for (int k = 0; k < S; ++k)
        fetch_row(ih_top + 3 - S + k, nx[k]);
    const RequantPacked kp = make_requant_packed(p.rq);
#pragma unroll 1
    for (uint32_t oh = oh0; oh < oh1; ++oh) {
        const int32_t ih = static_cast<int32_t>(oh) * S - p.pad_h;
        if (S == 1) {
            transpose_row(nx[0], A[2], B[2]);
        } else {
            transpose_row(nx[0], A[1], B[1]);
            transpose_row(nx[S - 1], A[2], B[2]);
        }
        if (oh + 1 < oh1) {
#pragma unroll
            for (int k = 0; k < S; ++k)
                fetch_row(ih + S + 3 - S + k, nx[k]);
        }
        const size_t orow = (static_cast<size_t>(oh) * p.OW + ow0) * p.CW + cw;
        uint32_t zw[TW];
#pragma unroll
        for (int t2 = 0; t2 < TW; ++t2)
            zw[t2] = (HAS_Z && ow0 + t2 < p.OW) ? __ldg(p.z + n * p.z_bwords + orow + static_cast<size_t>(t2) * p.CW) : 0u;
#pragma unroll
        for (int t2 = 0; t2 < TW; ++t2) {
            int32_t acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int32_t a = bias[k];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    uint32_t x, w;
                    if (S == 1) {
                        x = (t2 < 2) ? A[r][k] : B[r][k];
                        w = (t2 & 1) ? wb[k][r] : wa[k][r];
                    } else {
                        x = t2 == 0 ? A[r][k] : B[r][k];
                        w = wa[k][r];
                    }
                    a = __dp4a(static_cast<int32_t>(x), static_cast<int32_t>(w), a);
                }
                acc[k] = a;
            }
            uint32_t o;
            if constexpr (MODE != 0) {
                const uint32_t zx = zw[t2] ^ 0x80808080u;
                uint32_t ub[4];
#pragma unroll
                for (int b = 0; b < 4; b += 2) {
                    const uint32_t z0 = __byte_perm(zx, 0x4B000000u, 0x7540u | b);
                    const uint32_t z1 = __byte_perm(zx, 0x4B000000u, 0x7540u | (b + 1));
                    if constexpr (MODE == 2)
                        requant_pair_nonneg<HAS_Z>(acc[b], acc[b + 1], z0, z1, kp, ub[b], ub[b + 1]);
                    else
                        requant_pair_signed<HAS_Z>(acc[b], acc[b + 1], z0, z1, kp, ub[b], ub[b + 1]);
                }
                o = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
            } else {
                o = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    o |= (static_cast<uint32_t>(requant_s8(acc[b], HAS_Z ? sx8(zw[t2], b) : 0, p.nonline, p.rq)) & 0xffu) << (8 * b);
            }
            if (ow0 + t2 < p.OW)
                p.dst[n * p.dst_bwords + orow + static_cast<size_t>(t2) * p.CW] = o;
        }
        // slide the window down
        if (S == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                A[0][k] = A[1][k], B[0][k] = B[1][k];
                A[1][k] = A[2][k], B[1][k] = B[2][k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                A[0][k] = A[2][k], B[0][k] = B[2][k];
        }
    }
}

// The same window math behind the TMA front end of the 16-bit kernels (dw3x3_tma_kernel): input tiles staged
// in shared memory by a producer warp, 3-4 tiles in flight per SM whatever the occupancy.  The int8 tensor
// is seen as 16-bit "channels" (C / 2 of them), so tile geometry and tensor map are launch_dw3_tma's; a lane
// owns one 32-bit word = 4 int8 channels.  IDENTITY / RELU with ordinary scales only (MODE 1 / 2).
struct Dw3S8TmaParams {
    Dw3TmaParams t;
    RequantParams rq;
};

template <int S, int MODE, bool HAS_Z, int CPB>
__global__ void __launch_bounds__((DW3_COMPUTE_WARPS + 1) * 32, 1)
        dw3x3_s8_tma_kernel(const __grid_constant__ Dw3S8TmaParams qq) {
    constexpr int TW = S == 1 ? 4 : 2;
    constexpr int IN = (TW - 1) * S + 3;     // 6 (stride 1) or 5 (stride 2) pixel words per row
    constexpr uint32_t PIX = CPB * 4;        // bytes per pixel in the tile
    constexpr uint32_t CB = CPB * 2;
    using namespace ptx;
    const Dw3TmaParams& q = qq.t;
    const Dw3Common& p = q.c;
    extern __shared__ __align__(128) uint8_t dsm[];
    const uint32_t sbase = smem_u32(dsm);
    const uint32_t bar_full = sbase, bar_empty = sbase + 64;
    const uint32_t stage0 = sbase + 128;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tensormap(&q.tmS);
        for (uint32_t i = 0; i < q.stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, DW3_COMPUTE_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    const uint32_t cblk = blockIdx.x % q.cblks;
    const uint32_t sp0 = blockIdx.x / q.cblks, sp_step = gridDim.x / q.cblks;
    if (warp == DW3_COMPUTE_WARPS) {
        pdl_wait();
        uint32_t st = 0, ph = 0;
        for (uint32_t sp = sp0; sp < q.spatial_tiles; sp += sp_step) {
            const uint32_t tw = sp % q.tiles_w, r = sp / q.tiles_w;
            const uint32_t th = r % q.tiles_h, n = r / q.tiles_h;
            mbar_wait(bar_empty + 8 * st, ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8 * st, q.tile_bytes);
                tma_load_4d(
                        stage0 + st * q.stage_bytes, &q.tmS, bar_full + 8 * st,
                        static_cast<int32_t>(cblk * CB), static_cast<int32_t>(tw * q.TWT) * S - p.pad_w,
                        static_cast<int32_t>(th * q.TH) * S - p.pad_h, static_cast<int32_t>(n * q.NB));
            }
            __syncwarp();
            if (++st == q.stages) {
                st = 0;
                ph ^= 1;
            }
        }
        return;
    }
    const uint32_t cpl = lane % CPB, sub = lane / CPB;
    constexpr uint32_t SUBS = 32 / CPB;
    const uint32_t c = cblk * CB + cpl * 2;  // first 16-bit "channel" of the lane's word
    const bool ch_ok = c < p.C;
    const uint32_t cw = ch_ok ? c / 2 : 0;   // word of the pixel = 4 int8 channels
    pdl_wait();
    uint32_t wa[4][3];
    int32_t bias[4];
    {
        uint32_t fw[9];
        const uint32_t* f32p = reinterpret_cast<const uint32_t*>(p.filter) + static_cast<size_t>(cw) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            fw[i] = __ldg(f32p + i);
        auto fbyte = [&](int idx) { return (fw[idx >> 2] >> (8 * (idx & 3))) & 0xffu; };
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                uint32_t w = 0;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    const int tap = r * 3 + x;
                    w |= (p.flip ? fbyte(k * 9 + 8 - tap) : fbyte(k * 9 + tap)) << (8 * x);
                }
                wa[k][r] = w;
            }
            bias[k] = p.has_bias ? reinterpret_cast<const int32_t*>(p.bias)[cw * 4 + k] : 0;
        }
    }
    const RequantPacked kp = make_requant_packed(qq.rq);
    const uint32_t u = warp * SUBS + sub;
    const bool has_unit = u < q.units;
    const uint32_t strip = u % q.strips, span = (u / q.strips) % q.spans, img = u / (q.strips * q.spans);
    const uint32_t srow = q.IWT * PIX;
    const uint32_t unit_off = ((img * q.IHT + span * q.RS * S) * q.IWT + strip * TW * S) * PIX + cpl * 4;
    const uint32_t cstride = p.CP, rstride = p.OW * p.CP;
    uint32_t st = 0, ph = 0;
    for (uint32_t sp = sp0; sp < q.spatial_tiles; sp += sp_step) {
        const uint32_t tw = sp % q.tiles_w, r = sp / q.tiles_w;
        const uint32_t th = r % q.tiles_h, n = (r / q.tiles_h) * q.NB + img;
        mbar_wait(bar_full + 8 * st, ph);
        const uint32_t oh0 = th * q.TH + span * q.RS;
        uint32_t oh1 = oh0 + q.RS;
        if (oh1 > (th + 1) * q.TH)
            oh1 = (th + 1) * q.TH;
        if (oh1 > p.OH)
            oh1 = p.OH;
        const uint32_t ow0 = tw * q.TWT + strip * TW;
        if (has_unit && oh0 < oh1 && ow0 < p.OW && n < q.N) {
            const uint8_t* a = dsm + 128 + st * q.stage_bytes + unit_off;
            uint32_t cmask = 0;
#pragma unroll
            for (int t2 = 0; t2 < TW; ++t2)
                cmask |= (ch_ok && ow0 + t2 < p.OW ? 1u : 0u) << t2;
            uint32_t* drow = p.dst + n * p.dst_bwords + (static_cast<size_t>(oh0) * p.OW + ow0) * p.CP + c / 2;
            const uint32_t* zrow = p.z + n * p.z_bwords + (static_cast<size_t>(oh0) * p.OW + ow0) * p.CP + c / 2;
            uint32_t A[3][4], B[3][4];
            // one staged row -> per-channel words A[k] = pixels 0..3, B[k] = pixels 2..5
            auto lds_row = [&](uint32_t (&Ar)[4], uint32_t (&Br)[4]) {
                uint32_t px[6];
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    px[j] = j < IN ? *reinterpret_cast<const uint32_t*>(a + j * PIX) : 0u;
                a += srow;
                const uint32_t l01 = __byte_perm(px[0], px[1], 0x5140), h01 = __byte_perm(px[0], px[1], 0x7362);
                const uint32_t l23 = __byte_perm(px[2], px[3], 0x5140), h23 = __byte_perm(px[2], px[3], 0x7362);
                const uint32_t l45 = __byte_perm(px[4], px[5], 0x5140), h45 = __byte_perm(px[4], px[5], 0x7362);
                Ar[0] = __byte_perm(l01, l23, 0x5410), Ar[1] = __byte_perm(l01, l23, 0x7632);
                Ar[2] = __byte_perm(h01, h23, 0x5410), Ar[3] = __byte_perm(h01, h23, 0x7632);
                Br[0] = __byte_perm(l23, l45, 0x5410), Br[1] = __byte_perm(l23, l45, 0x7632);
                Br[2] = __byte_perm(h23, h45, 0x5410), Br[3] = __byte_perm(h23, h45, 0x7632);
            };
            lds_row(A[0], B[0]);
            if constexpr (S == 1)
                lds_row(A[1], B[1]);
#pragma unroll 1
            for (uint32_t rows = oh1 - oh0; rows > 0; --rows) {
                if constexpr (S == 2)
                    lds_row(A[1], B[1]);
                lds_row(A[2], B[2]);
                uint32_t zw[TW];
#pragma unroll
                for (int t2 = 0; t2 < TW; ++t2)
                    zw[t2] = (HAS_Z && ((cmask >> t2) & 1)) ? __ldg(zrow + t2 * cstride) : 0u;
#pragma unroll
                for (int t2 = 0; t2 < TW; ++t2) {
                    int32_t acc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int32_t av = bias[k];
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) {
                            uint32_t x, w;
                            if (S == 1) {
                                x = (t2 < 2) ? A[rr][k] : B[rr][k];
                                w = (t2 & 1) ? (wa[k][rr] << 8) : wa[k][rr];
                            } else {
                                x = t2 == 0 ? A[rr][k] : B[rr][k];
                                w = wa[k][rr];
                            }
                            av = __dp4a(static_cast<int32_t>(x), static_cast<int32_t>(w), av);
                        }
                        acc[k] = av;
                    }
                    const uint32_t zx = zw[t2] ^ 0x80808080u;
                    uint32_t ub[4];
#pragma unroll
                    for (int b = 0; b < 4; b += 2) {
                        const uint32_t z0 = __byte_perm(zx, 0x4B000000u, 0x7540u | b);
                        const uint32_t z1 = __byte_perm(zx, 0x4B000000u, 0x7540u | (b + 1));
                        if constexpr (MODE == 2)
                            requant_pair_nonneg<HAS_Z>(acc[b], acc[b + 1], z0, z1, kp, ub[b], ub[b + 1]);
                        else
                            requant_pair_signed<HAS_Z>(acc[b], acc[b + 1], z0, z1, kp, ub[b], ub[b + 1]);
                    }
                    if ((cmask >> t2) & 1)
                        drow[t2 * cstride] = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
                }
                drow += rstride;
                zrow += rstride;
                if constexpr (S == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        A[0][k] = A[1][k], B[0][k] = B[1][k];
                        A[1][k] = A[2][k], B[1][k] = B[2][k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        A[0][k] = A[2][k], B[0][k] = B[2][k];
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(bar_empty + 8 * st);
        if (++st == q.stages) {
            st = 0;
            ph ^= 1;
        }
    }
}

bool dw3_s8_applicable(const b200cb_desc& d) {
    return d.src_dtype == B200CB_DT_QS8 && d.dst_dtype == B200CB_DT_QS8 && d.fh == 3 && d.fw == 3 &&
           d.stride_h == d.stride_w && (d.stride_h == 1 || d.stride_h == 2) && d.dilate_h == 1 && d.dilate_w == 1 &&
           d.ic % 4 == 0 && d.src_batch_stride % 4 == 0 && d.dst_batch_stride % 4 == 0 && d.z_batch_stride % 4 == 0;
}

// experiment knobs (tools/): integer environment override with a default
int env_flag(const char* name, int dflt) {
#ifdef B200CB_TOOLS  // the product library never reads the environment
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// launch with programmatic stream serialization (the kernels above call griddepcontrol themselves)
template <class Kern, class Params>
cudaError_t launch_pdl(Kern kern, uint32_t grid, uint32_t threads, uint32_t smem, cudaStream_t stream, const Params& q) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    static const bool use_pdl = env_flag("B200CB_NO_PDL", 0) == 0;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, q);
}

// Float32 3x3: the same problem seen as 16-bit half-channels (ic = oc = 2 C; element strides double)
bool dw3_f32(const b200cb_desc& d) {
    return d.src_dtype == B200CB_DT_F32 && d.fh == 3 && d.fw == 3 && d.stride_h == d.stride_w &&
           (d.stride_h == 1 || d.stride_h == 2) && d.dilate_h == 1 && d.dilate_w == 1;
}
b200cb_desc dw3_f32_view(const b200cb_desc& d) {
    b200cb_desc v = d;
    v.src_dtype = v.filter_dtype = v.dst_dtype = B200CB_DT_BF16;  // (16-bit element counting only)
    v.ic = v.oc = v.group = 2 * d.ic;
    v.src_batch_stride *= 2, v.dst_batch_stride *= 2, v.z_batch_stride *= 2;
    return v;
}
bool dw3_applicable(const b200cb_desc& d) {
    return (d.src_dtype == B200CB_DT_BF16 || d.src_dtype == B200CB_DT_F16) && d.fh == 3 && d.fw == 3 &&
           d.stride_h == d.stride_w && (d.stride_h == 1 || d.stride_h == 2) && d.dilate_h == 1 &&
           d.dilate_w == 1 && d.ic % 2 == 0 && d.src_batch_stride % 2 == 0 && d.dst_batch_stride % 2 == 0 &&
           d.z_batch_stride % 2 == 0;
}
bool dw3_tma_applicable(const b200cb_desc& d) {
    return dw3_applicable(d) && d.ic % 8 == 0 && (d.src_batch_stride * 2) % 16 == 0 && d.pad_h <= 8 && d.pad_w <= 8;
}

bool dw3_bulk_applicable(const b200cb_desc& d) {
    const uint64_t row = (uint64_t)d.iw * d.ic * 2;
    return dw3_applicable(d) && row % 16 == 0 && (d.src_batch_stride * 2) % 16 == 0 && d.pad_h <= 1 &&
           d.pad_w <= 1 && d.ic / 2 <= DW3_COMPUTE_WARPS * 32 && 3 * (row + 8 * d.ic) <= 100 * 1024;
}

template <int KIND>
int launch_dw3_bulk(const b200cb_desc& d, const Dw3Common& c, const void* src, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(src) % 16)
        return B200CB_UNSUPPORTED_SHAPE;  // falls through to the global-memory front end
    Dw3BulkParams q;
    memset(&q, 0, sizeof(q));
    q.c = c;
    q.src = static_cast<const uint8_t*>(src);
    q.src_bbytes = (d.src_batch_stride ? d.src_batch_stride : (uint64_t)d.ih * d.iw * d.ic) * 2;
    q.row_bytes = d.iw * d.ic * 2;
    q.row_stride = (d.iw + 4) * d.ic * 2;
    const uint32_t pix = d.ic * 2;
    const uint32_t S = d.stride_h, TW = S == 1 ? 4 : 2;
    q.strips = (d.ow + TW - 1) / TW;
    q.U = DW3_COMPUTE_WARPS * 32 / (d.ic / 2);
    uint32_t spans = q.strips >= q.U ? 1 : q.U / q.strips;
    if (spans > d.oh)
        spans = d.oh;
    // rows per unit from the stage budget (3 stages of <= 64 KiB, else 2 of <= 100 KiB), then balanced
    // over the tiles of one image
    uint32_t budget = 64 * 1024;
    if (3 * q.row_stride > budget)
        budget = 100 * 1024;
    if (((spans - 1) * S + 3) * (uint64_t)q.row_stride > budget)
        spans = 1;
    uint32_t rs_max = 1;
    while (((spans * (rs_max + 1) - 1) * S + 3) * (uint64_t)q.row_stride <= budget && spans * rs_max < d.oh)
        ++rs_max;
    const uint32_t th_max = spans * rs_max;
    q.tiles_h = (d.oh + th_max - 1) / th_max;
    const uint32_t th_need = (d.oh + q.tiles_h - 1) / q.tiles_h;
    q.RS = (th_need + spans - 1) / spans;
    q.spans = spans;
    q.TH = spans * q.RS;
    q.IHT = (q.TH - 1) * S + 3;
    q.tiles_h = (d.oh + q.TH - 1) / q.TH;
    const uint64_t total = (uint64_t)d.n * q.tiles_h;
    if (total > 0x7fffffffull)
        return B200CB_UNSUPPORTED_SHAPE;
    q.total_tiles = (uint32_t)total;
    q.zrow_bytes = q.row_stride;     // followed by the zero pixels of stage 0's first row
    q.stage_bytes = q.IHT * q.row_stride;
    const uint32_t fixed = 128 + q.zrow_bytes + 4 * pix;
    if (fixed + 2 * q.stage_bytes > MAX_DYN_SMEM)
        return B200CB_UNSUPPORTED_SHAPE;
    q.stages = (MAX_DYN_SMEM - fixed) / q.stage_bytes;
    if (q.stages > 4)
        q.stages = 4;
    q.smem_bytes = fixed + q.stages * q.stage_bytes;
    const uint32_t sms = (uint32_t)device_sm_count();
    const uint32_t grid = total < sms ? (uint32_t)total : sms;
    auto go = [&](auto kern) -> cudaError_t {
        // (both stride variants have the same function type: the flag is keyed by device AND stride)
        static bool attr_set[64][2] = {{false}};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev][S - 1]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM);
            if (e != cudaSuccess)
                return e;
            if (dev >= 0 && dev < 64)
                attr_set[dev][S - 1] = true;
        }
        return launch_pdl(kern, grid, (DW3_COMPUTE_WARPS + 1) * 32, q.smem_bytes, stream, q);
    };
    const cudaError_t e = S == 1 ? go(dw3x3_bulk_kernel<KIND, 1, 4>) : go(dw3x3_bulk_kernel<KIND, 2, 2>);
    if (e != cudaSuccess) {
        set_last_error("dw3x3 bulk launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

template <int KIND>
cudaError_t launch_dw3(const b200cb_desc& d, Dw3Params& p, cudaStream_t stream) {
    const uint32_t tw = d.stride_h == 1 ? 4 : 2;
    p.strips = (d.ow + tw - 1) / tw;
    // vertical span: long enough to amortise the 2 halo rows, short enough to fill the machine
    uint32_t rh = 16;
    auto items = [&](uint32_t r) { return (uint64_t)p.c.CP * p.strips * ((d.oh + r - 1) / r) * d.n; };
    while (rh > 4 && items(rh) < 148ull * 2048)
        rh /= 2;
    p.RH = rh;
    p.spans = (d.oh + rh - 1) / rh;
    p.total_items = items(rh);
    const uint64_t blocks = (p.total_items + 255) / 256;
    if (blocks > 0x7fffffffull)
        return cudaErrorInvalidValue;
    if (d.stride_h == 1)
        dw3x3_kernel<KIND, 1, 4><<<(uint32_t)blocks, 256, 0, stream>>>(p);
    else
        dw3x3_kernel<KIND, 2, 2><<<(uint32_t)blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

template <int KIND, int S, int TW, int CPB>
cudaError_t launch_dw3_tma_variant(const Dw3TmaParams& q, uint32_t grid, uint32_t smem, cudaStream_t stream) {
    auto kern = dw3x3_tma_kernel<KIND, S, TW, CPB>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM);
        if (e != cudaSuccess)
            return e;
        attr_set[dev] = true;
    }
    return launch_pdl(kern, grid, (DW3_COMPUTE_WARPS + 1) * 32, smem, stream, q);
}

template <int S, int MODE, bool HAS_Z, int CPB>
cudaError_t launch_s8_tma_variant(const Dw3S8TmaParams& qq, uint32_t grid, uint32_t smem, cudaStream_t stream) {
    auto kern = dw3x3_s8_tma_kernel<S, MODE, HAS_Z, CPB>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM);
        if (e != cudaSuccess)
            return e;
        attr_set[dev] = true;
    }
    return launch_pdl(kern, grid, (DW3_COMPUTE_WARPS + 1) * 32, smem, stream, qq);
}
template <int S, int CPB>
cudaError_t launch_s8_tma(const Dw3S8TmaParams& qq, int mode, bool has_z, uint32_t grid, uint32_t smem, cudaStream_t st) {
    if (mode == 2)
        return has_z ? launch_s8_tma_variant<S, 2, true, CPB>(qq, grid, smem, st)
                     : launch_s8_tma_variant<S, 2, false, CPB>(qq, grid, smem, st);
    return has_z ? launch_s8_tma_variant<S, 1, true, CPB>(qq, grid, smem, st)
                 : launch_s8_tma_variant<S, 1, false, CPB>(qq, grid, smem, st);
}

// tile geometry + tensor map of the TMA front end (KIND == QuantizedS8: `d` is the 16-bit view of the int8
// tensors, `rq8` / `mode8` select the int8 window kernel)
template <int KIND>
int launch_dw3_tma(const b200cb_desc& d, const Dw3Common& c, const void* src, cudaStream_t stream,
                   const RequantParams* rq8 = nullptr, int mode8 = 0) {
    Dw3TmaParams q;
    memset(&q, 0, sizeof(q));
    q.c = c;
    const uint32_t S = d.stride_h, TW = S == 1 ? 4 : 2;
    // channel block: 64 channels per warp, or 32 (two units per warp) when that pads the channel count less
    uint32_t CB = d.ic > 32 ? 64 : 32;
    if (env_flag("B200CB_DW_CB32", 1) && (d.ic + 31) / 32 * 32 < (d.ic + 63) / 64 * 64)
        CB = 32;
    const uint32_t CPB = CB / 2, SUBS = 32 / CPB;
    // column strips per tile: 7 (28 / 14 output columns) divides the MobileNet widths; small images
    // take their whole width.  One unit per compute warp (x SUBS): spans = floor(15 * SUBS / strips).
    const uint32_t strips_full = (d.ow + TW - 1) / TW;
    q.strips = strips_full <= 8 ? strips_full : 7;
    q.TWT = q.strips * TW;
    q.IWT = (q.TWT - 1) * S + 3;
    uint32_t spans = DW3_COMPUTE_WARPS * SUBS / q.strips;
    if (spans < 1)
        spans = 1;
    if (spans > d.oh)
        spans = d.oh;
    // rows per unit: as many as a ~64 KiB stage holds (3 stages in flight), at most the image
    const uint32_t budget = 64 * 1024;
    uint32_t rs = (d.oh + spans - 1) / spans;
    while (rs > 1 && ((spans * rs - 1) * S + 3) * q.IWT * CB * 2 > budget)
        --rs;
    q.NB = 1;
    if (spans * rs >= d.oh && q.TWT >= d.ow) {
        // the whole image fits in one tile: give every unit a full-height strip and put several images
        // side by side in the tile (a 7 x 7 image alone is 10 KiB and one output row per unit)
        const uint32_t img_bytes = ((d.oh - 1) * S + 3) * q.IWT * CB * 2;
        uint32_t nb = DW3_COMPUTE_WARPS * SUBS / q.strips;
        if (nb > budget / img_bytes)
            nb = budget / img_bytes;
        if (nb > d.n)
            nb = d.n;
        if (nb >= 2 && env_flag("B200CB_DW_NB", 1)) {
            q.NB = nb;
            spans = 1;
            rs = d.oh;
        }
    }
    q.RS = rs;
    q.spans = spans;
    q.N = d.n;
    q.TH = spans * rs;
    q.units = q.strips * spans * q.NB;
    q.IHT = (q.TH - 1) * S + 3;
    if (q.IHT > 256 || q.IWT > 256 || q.units > DW3_COMPUTE_WARPS * SUBS)
        return B200CB_UNSUPPORTED_SHAPE;
    q.tiles_h = (d.oh + q.TH - 1) / q.TH;
    q.tiles_w = (d.ow + q.TWT - 1) / q.TWT;
    q.cblks = (d.ic + CB - 1) / CB;
    const uint64_t spatial = (uint64_t)((d.n + q.NB - 1) / q.NB) * q.tiles_h * q.tiles_w;
    if (spatial > 0x7fffffffull)
        return B200CB_UNSUPPORTED_SHAPE;
    q.spatial_tiles = (uint32_t)spatial;
    q.tile_bytes = q.NB * q.IHT * q.IWT * CB * 2;
    q.stage_bytes = (q.tile_bytes + 127) & ~127u;
    q.stages = (MAX_DYN_SMEM - 128) / q.stage_bytes;
    if (q.stages > 4)
        q.stages = 4;
    if (q.stages < 2)
        return B200CB_UNSUPPORTED_SHAPE;
    const uint64_t src_b = d.src_batch_stride ? d.src_batch_stride : (uint64_t)d.ih * d.iw * d.ic;
    uint64_t dims[4] = {d.ic, d.iw, d.ih, d.n};
    uint64_t strides[3] = {(uint64_t)d.ic * 2, (uint64_t)d.iw * d.ic * 2, src_b * 2};
    uint32_t box[4] = {CB, q.IWT, q.IHT, q.NB};
    CUresult cr = encode_tiled(
            &q.tmS, KIND == B200CB_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
            src, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (cr != CUDA_SUCCESS) {
        char buf[64];
        snprintf(buf, sizeof(buf), "CUresult %d (depthwise src map)", (int)cr);
        set_last_error("cuTensorMapEncodeTiled", buf);
        return B200CB_ERR_DRIVER;
    }
    // persistent grid: a multiple of the channel-block count, at most one CTA per SM
    const uint32_t sms = (uint32_t)device_sm_count();
    if (q.cblks > sms)
        return B200CB_UNSUPPORTED_SHAPE;
    uint64_t per = sms / q.cblks;
    if (per > spatial)
        per = spatial;
    const uint32_t grid = (uint32_t)per * q.cblks;
    const uint32_t smem = 128 + q.stages * q.stage_bytes;
    cudaError_t e;
    if constexpr (KIND == B200CB_DT_QS8) {
        Dw3S8TmaParams qq;
        qq.t = q;
        qq.rq = *rq8;
        if (S == 1)
            e = CPB == 32 ? launch_s8_tma<1, 32>(qq, mode8, c.has_z != 0, grid, smem, stream)
                          : launch_s8_tma<1, 16>(qq, mode8, c.has_z != 0, grid, smem, stream);
        else
            e = CPB == 32 ? launch_s8_tma<2, 32>(qq, mode8, c.has_z != 0, grid, smem, stream)
                          : launch_s8_tma<2, 16>(qq, mode8, c.has_z != 0, grid, smem, stream);
    } else if (S == 1)
        e = CPB == 32 ? launch_dw3_tma_variant<KIND, 1, 4, 32>(q, grid, smem, stream)
                      : launch_dw3_tma_variant<KIND, 1, 4, 16>(q, grid, smem, stream);
    else
        e = CPB == 32 ? launch_dw3_tma_variant<KIND, 2, 2, 32>(q, grid, smem, stream)
                      : launch_dw3_tma_variant<KIND, 2, 2, 16>(q, grid, smem, stream);
    if (e != cudaSuccess) {
        set_last_error("dw3x3 tma launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
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
    const uint32_t es = dtype_size(dt);
    // the 3x3 bandwidth kernels work on 32-bit words (16-bit channel pairs / one fp32 channel); everything
    // else on 16-byte vectors
    if (!dw3_applicable(d) && !dw3_f32(d) && !dw3_s8_applicable(d)) {
        if (d.ic % vec != 0)
            return B200CB_UNSUPPORTED_ALIGNMENT;
        if ((d.src_batch_stride * es) % 16 || (d.dst_batch_stride * es) % 16 || (d.z_batch_stride * es) % 16)
            return B200CB_UNSUPPORTED_ALIGNMENT;
    }
    if (d.fh * d.fw > 1024)
        return B200CB_UNSUPPORTED_SHAPE;

    const uint32_t cv = d.ic / vec ? d.ic / vec : 1;
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
    const bool f32_dw3 = dw3_f32(plan.desc);
    // (Float32 3x3: the half-channel view of the same tensors, see dw3_f32_view)
    const b200cb_desc d = f32_dw3 ? dw3_f32_view(plan.desc) : plan.desc;
    if (!src || !filter || !dst || (d.bias_mode != B200CB_BIAS_NONE && !bias) || (d.has_z && !z))
        return B200CB_ERR_NULL;
    if (dw3_applicable(d)) {
        Dw3Common c;
        memset(&c, 0, sizeof(c));
        c.filter = filter, c.bias = bias;
        c.z = static_cast<const uint32_t*>(z), c.dst = static_cast<uint32_t*>(dst);
        c.N = d.n, c.C = d.ic, c.IH = d.ih, c.IW = d.iw, c.OH = d.oh, c.OW = d.ow, c.CP = d.ic / 2;
        c.pad_h = d.pad_h, c.pad_w = d.pad_w;
        c.dst_bwords = (d.dst_batch_stride ? d.dst_batch_stride : (uint64_t)d.oh * d.ow * d.oc) / 2;
        c.z_bwords = (d.z_batch_stride ? d.z_batch_stride : (uint64_t)d.oh * d.ow * d.oc) / 2;
        c.flip = d.mode == B200CB_MODE_CONVOLUTION;
        c.has_bias = d.bias_mode != B200CB_BIAS_NONE;
        c.has_z = d.has_z;
        c.nonline = d.nonline_mode;
        if (dw3_tma_applicable(d) && plan.reserved[2] != 2) {  // reserved[2] == 2: force the register front end
            int rc = f32_dw3                          ? launch_dw3_tma<B200CB_DT_F32>(d, c, src, stream)
                     : d.src_dtype == B200CB_DT_BF16 ? launch_dw3_tma<B200CB_DT_BF16>(d, c, src, stream)
                                                     : launch_dw3_tma<B200CB_DT_F16>(d, c, src, stream);
            if (rc != B200CB_UNSUPPORTED_SHAPE)
                return rc;
        }
        if (dw3_bulk_applicable(d) && plan.reserved[2] != 2 && env_flag("B200CB_DW_BULK", 1)) {
            int rc = f32_dw3                          ? launch_dw3_bulk<B200CB_DT_F32>(d, c, src, stream)
                     : d.src_dtype == B200CB_DT_BF16 ? launch_dw3_bulk<B200CB_DT_BF16>(d, c, src, stream)
                                                     : launch_dw3_bulk<B200CB_DT_F16>(d, c, src, stream);
            if (rc != B200CB_UNSUPPORTED_SHAPE)
                return rc;
        }
        Dw3Params q;
        memset(&q, 0, sizeof(q));
        q.c = c;
        q.src = static_cast<const uint32_t*>(src);
        q.src_bwords = (d.src_batch_stride ? d.src_batch_stride : (uint64_t)d.ih * d.iw * d.ic) / 2;
        cudaError_t e3 = f32_dw3                          ? launch_dw3<B200CB_DT_F32>(d, q, stream)
                         : d.src_dtype == B200CB_DT_BF16 ? launch_dw3<B200CB_DT_BF16>(d, q, stream)
                                                         : launch_dw3<B200CB_DT_F16>(d, q, stream);
        if (e3 != cudaSuccess) {
            set_last_error("dw3x3 launch", cudaGetErrorString(e3));
            return B200CB_ERR_CUDA;
        }
        return B200CB_OK;
    }
    if (dw3_s8_applicable(d) && plan.reserved[2] != 2 && d.ic % 16 == 0) {
        // TMA-staged window kernel: the int8 tensors as C / 2 16-bit channels (fast IDENTITY / RELU only)
        const RequantParams rq = make_requant(d);
        const bool fast = rq.fast_div == 2 && (d.nonline_mode == B200CB_NL_RELU || d.nonline_mode == B200CB_NL_IDENTITY);
        b200cb_desc v = d;
        v.src_dtype = v.filter_dtype = v.dst_dtype = B200CB_DT_BF16;  // (16-bit element counting only)
        v.ic = v.oc = v.group = d.ic / 2;
        v.src_batch_stride /= 2, v.dst_batch_stride /= 2, v.z_batch_stride /= 2;
        if (fast && dw3_tma_applicable(v) && d.src_batch_stride % 2 == 0 && d.dst_batch_stride % 4 == 0 &&
            d.z_batch_stride % 4 == 0) {
            Dw3Common c;
            memset(&c, 0, sizeof(c));
            c.filter = filter, c.bias = bias;
            c.z = static_cast<const uint32_t*>(z), c.dst = static_cast<uint32_t*>(dst);
            c.N = d.n, c.C = v.ic, c.IH = d.ih, c.IW = d.iw, c.OH = d.oh, c.OW = d.ow, c.CP = d.ic / 4;
            c.pad_h = d.pad_h, c.pad_w = d.pad_w;
            c.dst_bwords = (d.dst_batch_stride ? d.dst_batch_stride : (uint64_t)d.oh * d.ow * d.oc) / 4;
            c.z_bwords = (d.z_batch_stride ? d.z_batch_stride : (uint64_t)d.oh * d.ow * d.oc) / 4;
            c.flip = d.mode == B200CB_MODE_CONVOLUTION;
            c.has_bias = d.bias_mode != B200CB_BIAS_NONE, c.has_z = d.has_z, c.nonline = d.nonline_mode;
            const int rc = launch_dw3_tma<B200CB_DT_QS8>(v, c, src, stream, &rq, d.nonline_mode == B200CB_NL_RELU ? 2 : 1);
            if (rc != B200CB_UNSUPPORTED_SHAPE)
                return rc;
        }
    }
    if (dw3_s8_applicable(d) && plan.reserved[2] != 2) {
        DwS8Params q;
        memset(&q, 0, sizeof(q));
        q.src = static_cast<const uint32_t*>(src), q.filter = static_cast<const int8_t*>(filter);
        q.bias = static_cast<const int32_t*>(bias), q.z = static_cast<const uint32_t*>(z);
        q.dst = static_cast<uint32_t*>(dst);
        q.N = d.n, q.C = d.ic, q.IH = d.ih, q.IW = d.iw, q.OH = d.oh, q.OW = d.ow, q.CW = d.ic / 4;
        q.pad_h = d.pad_h, q.pad_w = d.pad_w;
        q.src_bwords = (d.src_batch_stride ? d.src_batch_stride : (uint64_t)d.ih * d.iw * d.ic) / 4;
        q.dst_bwords = (d.dst_batch_stride ? d.dst_batch_stride : (uint64_t)d.oh * d.ow * d.oc) / 4;
        q.z_bwords = (d.z_batch_stride ? d.z_batch_stride : (uint64_t)d.oh * d.ow * d.oc) / 4;
        q.flip = d.mode == B200CB_MODE_CONVOLUTION;
        q.has_bias = d.bias_mode != B200CB_BIAS_NONE, q.has_z = d.has_z, q.nonline = d.nonline_mode;
        q.rq = make_requant(d);
        const uint32_t tw = d.stride_h == 1 ? 4 : 2;
        q.strips = (d.ow + tw - 1) / tw;
        // vertical span: long enough to amortise the halo rows, short enough to fill the machine
        uint32_t rh = 16;
        auto items = [&](uint32_t r) { return (uint64_t)q.CW * q.strips * ((d.oh + r - 1) / r) * d.n; };
        while (rh > 2 && items(rh) < 148ull * 4096)
            rh /= 2;
        q.RH = rh;
        q.spans = (d.oh + rh - 1) / rh;
        q.total_items = items(rh);
        const uint64_t blocks = (q.total_items + 255) / 256;
        if (blocks <= 0x7fffffffull) {
            const bool fast = q.rq.fast_div == 2 && (d.nonline_mode == B200CB_NL_RELU || d.nonline_mode == B200CB_NL_IDENTITY);
            const int mode = !fast ? 0 : (d.nonline_mode == B200CB_NL_RELU ? 2 : 1);
            cudaError_t e8 = cudaErrorInvalidValue;
#define B200CB_DW8(S_, M_, Z_) e8 = launch_pdl(dw3x3_s8_kernel<S_, M_, Z_>, (uint32_t)blocks, 256, 0, stream, q)
#define B200CB_DW8_S(S_)                                              \
    if (d.has_z) {                                                     \
        if (mode == 2) B200CB_DW8(S_, 2, true);                        \
        else if (mode == 1) B200CB_DW8(S_, 1, true);                   \
        else B200CB_DW8(S_, 0, true);                                  \
    } else {                                                           \
        if (mode == 2) B200CB_DW8(S_, 2, false);                       \
        else if (mode == 1) B200CB_DW8(S_, 1, false);                  \
        else B200CB_DW8(S_, 0, false);                                 \
    }
            if (d.stride_h == 1) {
                B200CB_DW8_S(1)
            } else {
                B200CB_DW8_S(2)
            }
#undef B200CB_DW8_S
#undef B200CB_DW8
            if (e8 != cudaSuccess) {
                set_last_error("dw3x3 s8 launch", cudaGetErrorString(e8));
                return B200CB_ERR_CUDA;
            }
            return B200CB_OK;
        }
    }
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

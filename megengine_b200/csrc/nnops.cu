// Operators either side of ConvBiasForward (include/b200cb_nn.h): pooling,
// elementwise (plain and quantized), TypeCvt -- pure HBM-bound kernels, 16-byte vectors, grids sized in
// multiples of the SM count -- and the convolution front-ends (ConvolutionForward, ConvolutionBackwardData,
// BatchConvBiasForward) that run on the ConvBias kernels.
//
// Arithmetic restates the reference's naive operators element for element:
//   pooling      dnn/src/naive/pooling/opr_impl.cpp:17-330 (MaxPooler, MeanInclude / MeanExclude poolers: half
//                types accumulate IN the half type, QuantizedS8 averages round half away from zero)
//   elemwise     dnn/src/common/elemwise/kern_defs.cuh:169-270 (ctype arithmetic: half types round every op)
//   quantized    dnn/src/naive/elemwise_multi_type/opr_impl.h:45-120 (quantize(op(dequantize(a), dequantize(b))))
//   TypeCvt      dnn/src/naive/type_cvt/opr_impl.cpp:24-57, dnn/include/megdnn/dtype.h:845-897
//   deconv       dnn/src/naive/convolution/opr_impl.cpp:92-167 (fp32 / int32 accumulation, one final TypeCvt)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <string.h>
#include "b200cb_nn.h"
#include "common.h"
#include "epilogue.cuh"

namespace b200cb {
namespace {

// ---------------------------------------------------------------- element helpers
template <int DT>
struct Elem;
template <>
struct Elem<B200CB_DT_F32> {
    using T = float;
    static __device__ __forceinline__ float ld(T v) { return v; }
    static __device__ __forceinline__ T st(float v) { return v; }
};
// float -> fp16 the way the reference's half type rounds (third_party half.hpp built with HALF_ROUND_STYLE 1,
// HALF_ROUND_TIES_TO_EVEN 0): to nearest, TIES AWAY FROM ZERO.  Sums of two fp16 values tie all the time, so the
// pooling / elementwise results are bit-exact only with this rounding (cvt.rn.f16.f32 breaks ties to even).
__device__ __forceinline__ __half f2h_ties_away(float v) {
    // cvt.rn breaks ties to even; setting the lowest fp32 mantissa bit moves an exact tie (dropped bits 100..0)
    // just past the midpoint -- away from zero -- and cannot move any other value across a rounding boundary
    // (dropped bits d != tie: d | 1 stays on d's side of the tie).  Infinities keep their bits.
    const uint32_t bits = __float_as_uint(v);
    const uint32_t finite = (bits & 0x7f800000u) != 0x7f800000u;
    return __float2half_rn(__uint_as_float(bits | finite));
}
template <>
struct Elem<B200CB_DT_F16> {
    using T = __half;
    static __device__ __forceinline__ float ld(T v) { return __half2float(v); }
    static __device__ __forceinline__ T st(float v) { return f2h_ties_away(v); }
};
template <>
struct Elem<B200CB_DT_BF16> {
    using T = __nv_bfloat16;
    static __device__ __forceinline__ float ld(T v) { return __bfloat162float(v); }
    static __device__ __forceinline__ T st(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Elem<B200CB_DT_QS8> {
    using T = int8_t;
    static __device__ __forceinline__ float ld(T v) { return static_cast<float>(v); }
    static __device__ __forceinline__ T st(float v) { return static_cast<int8_t>(v); }
};
// a value rounded to the element type and widened again (every ctype operation of the reference rounds)
template <int DT>
__device__ __forceinline__ float rnd(float v) {
    return Elem<DT>::ld(Elem<DT>::st(v));
}

template <typename T, int V>
struct alignas(sizeof(T) * V > 16 ? 16 : sizeof(T) * V) Vec {
    T v[V];
};

inline uint32_t grid_for(uint64_t items, uint32_t threads) {
    const uint64_t blocks = (items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)device_sm_count() * 16;   // persistent-ish: grid-stride beyond 16 CTAs / SM
    return (uint32_t)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// ---------------------------------------------------------------- pooling
struct PoolParams {
    uint32_t n, c, ih, iw, oh, ow, wh, ww, sh, sw, ph, pw;
    uint32_t cg;          // channel groups of V elements a thread handles (NHWC: c / V, NCHW4: c / 4, NCHW: c)
    uint32_t layout;      // 0 NHWC (pixel pitch c), 1 NCHW (planes), 2 NCHW4 (planes of 4-channel pixels)
    float count_all;      // window_h * window_w rounded to the element type (MeanIncludePooler::count)
    float inv_count;      // AVERAGE with a power-of-two count: 1 / count (x * 2^-k == x / 2^k, bit for bit), else 0
    uint32_t tpr;         // threads per output row (power of two <= 256)
};

// lane-wise maximum of two packed words of the element type (MaxPooler: answer > x ? answer : x; NaN inputs aside)
template <int DT>
__device__ __forceinline__ uint32_t packed_max(uint32_t a, uint32_t b) {
    if constexpr (DT == B200CB_DT_QS8) {
        // a, b: two bytes each, biased to unsigned and widened to the fp16 bit patterns 0x0400 + u (normal numbers
        // of one binade: their order is the order of u) -- see the pooling kernel; one native HMNMX2 per byte pair.
        // (__vmaxs4 and max.s16x2 are emulated on sm_100: ~8 instructions per byte)
        uint32_t r;
        asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    } else if constexpr (DT == B200CB_DT_F16) {
        uint32_t r;
        asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    } else if constexpr (DT == B200CB_DT_BF16) {
        uint32_t r;
        asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
        return r;
    } else {
        return __float_as_uint(fmaxf(__uint_as_float(a), __uint_as_float(b)));
    }
}

// one thread = one output pixel x V consecutive channels.  Work item = (output row, position in the row): 32-bit
// index arithmetic, the valid part of the window computed once (no bounds test per tap), loads of a window row
// issued together.
template <int DT, int MODE, int V>
__global__ void __launch_bounds__(256) pool_kernel(
        const typename Elem<DT>::T* __restrict__ src, typename Elem<DT>::T* __restrict__ dst, PoolParams p) {
    using T = typename Elem<DT>::T;
    const uint32_t row_items = p.layout == 0 ? p.ow * p.cg : p.ow;   // items of one output row (NHWC) / one plane row
    const uint32_t rows = p.layout == 0 ? p.n * p.oh : p.n * p.cg * p.oh;
    const uint32_t per_block = blockDim.x / p.tpr * 1u;           // rows a block handles at once (tpr threads per row)
    for (uint32_t r0 = blockIdx.x * per_block; r0 < rows; r0 += gridDim.x * per_block) {
        const uint32_t r = r0 + threadIdx.x / p.tpr;
        if (r >= rows)
            continue;
        for (uint32_t it = threadIdx.x % p.tpr; it < row_items; it += p.tpr) {
            uint32_t g, ow, oh, n;
            const T* sb;
            T* db;
            uint32_t pitch;   // elements between horizontally adjacent pixels
            if (p.layout == 0) {   // NHWC: channel groups fastest
                g = it % p.cg, ow = it / p.cg;
                oh = r % p.oh, n = r / p.oh;
                sb = src + (size_t)n * p.ih * p.iw * p.c + (size_t)g * V;
                db = dst + ((size_t)r * p.ow + ow) * p.c + (size_t)g * V;
                pitch = p.c;
            } else {               // planes: pixels fastest
                ow = it, oh = r % p.oh;
                const uint32_t plane = r / p.oh;   // n * cg + g
                sb = src + (size_t)plane * p.ih * p.iw * V;
                db = dst + ((size_t)r * p.ow + ow) * V;
                pitch = V;
            }
            // valid taps: ih = oh * sh - ph + fh in [0, ih)
            const int32_t h0 = (int32_t)(oh * p.sh) - (int32_t)p.ph, w0 = (int32_t)(ow * p.sw) - (int32_t)p.pw;
            const uint32_t fh_lo = h0 < 0 ? (uint32_t)-h0 : 0u, fw_lo = w0 < 0 ? (uint32_t)-w0 : 0u;
            const uint32_t fh_hi = min(p.wh, p.ih - (uint32_t)min(h0, (int32_t)p.ih)), fw_hi = min(p.ww, p.iw - (uint32_t)min(w0, (int32_t)p.iw));
            constexpr bool WORDS = MODE == B200CB_POOL_MAX && sizeof(T) * V >= 4;   // packed maxima on 32-bit words
            constexpr bool Q8W = WORDS && DT == B200CB_DT_QS8;   // int8: accumulators widened to fp16 pairs (2 regs / word)
            constexpr int NW = WORDS ? (int)(sizeof(T) * V / 4) * (Q8W ? 2 : 1) : 1;
            uint32_t wacc[NW];
            float acc[V];
            int32_t iacc[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
                acc[k] = MODE == B200CB_POOL_MAX ? (DT == B200CB_DT_QS8 ? -128.f : (DT == B200CB_DT_F16 ? -65504.f : -FLT_MAX)) : 0.f;
                iacc[k] = 0;
            }
            if (DT == B200CB_DT_BF16 && MODE == B200CB_POOL_MAX) {
#pragma unroll
                for (int k = 0; k < V; ++k)
                    acc[k] = __uint_as_float(0xff7f0000u);   // bfloat16 lowest
            }
#pragma unroll
            for (int k = 0; k < NW; ++k)   // the lowest value of the element type in every lane
                wacc[k] = DT == B200CB_DT_QS8 ? 0x04000400u : DT == B200CB_DT_F16 ? 0xfbfffbffu : DT == B200CB_DT_BF16 ? 0xff7fff7fu : 0xff7fffffu;
#define B200CB_POOL_FEED(X_)                                                                               \
    do {                                                                                                   \
        if constexpr (Q8W) {                                                                               \
            const uint32_t* xw__ = reinterpret_cast<const uint32_t*>(&(X_));                               \
            _Pragma("unroll") for (int k = 0; k < NW / 2; ++k) {                                           \
                const uint32_t u__ = xw__[k] ^ 0x80808080u;                                                \
                wacc[2 * k] = packed_max<DT>(wacc[2 * k], __byte_perm(u__, 0x04040404u, 0x5140));          \
                wacc[2 * k + 1] = packed_max<DT>(wacc[2 * k + 1], __byte_perm(u__, 0x04040404u, 0x7362));  \
            }                                                                                              \
        } else if constexpr (WORDS) {                                                                      \
            const uint32_t* xw__ = reinterpret_cast<const uint32_t*>(&(X_));                               \
            _Pragma("unroll") for (int k = 0; k < NW; ++k) wacc[k] = packed_max<DT>(wacc[k], xw__[k]);     \
        } else {                                                                                           \
            _Pragma("unroll") for (int k = 0; k < V; ++k) {                                                \
                if constexpr (MODE == B200CB_POOL_MAX) {                                                   \
                    const float xv__ = Elem<DT>::ld((X_).v[k]);                                            \
                    acc[k] = acc[k] > xv__ ? acc[k] : xv__;                                                \
                } else if constexpr (DT == B200CB_DT_QS8) {                                                \
                    iacc[k] += (X_).v[k];                                                                  \
                } else {                                                                                   \
                    acc[k] = rnd<DT>(__fadd_rn(acc[k], Elem<DT>::ld((X_).v[k])));                          \
                }                                                                                          \
            }                                                                                              \
        }                                                                                                  \
    } while (0)
            if (p.wh == 3 && p.ww == 3) {
                // the 3x3 window (ResNet stem pool): all nine loads issued before the first use
                Vec<T, V> x[9];
                bool ok[9];
                // one base pointer (tap (0, 0), possibly outside the image: only dereferenced where ok) and two
                // strides: the kernel is bound by its index arithmetic, not by the loads
                const ptrdiff_t rs = (ptrdiff_t)p.iw * pitch;
                const T* p00 = sb + ((ptrdiff_t)h0 * p.iw + w0) * (ptrdiff_t)pitch;
                const bool rok[3] = {fh_lo == 0 && fh_hi > 0, fh_lo <= 1 && fh_hi > 1, fh_hi > 2};
                const bool cok[3] = {fw_lo == 0 && fw_hi > 0, fw_lo <= 1 && fw_hi > 1, fw_hi > 2};
#pragma unroll
                for (uint32_t t = 0; t < 9; ++t) {
                    const uint32_t fh = t / 3, fw = t % 3;
                    ok[t] = rok[fh] && cok[fw];
                    if (ok[t])
                        x[t] = *reinterpret_cast<const Vec<T, V>*>(p00 + (ptrdiff_t)fh * rs + (ptrdiff_t)fw * pitch);
                }
#pragma unroll
                for (uint32_t t = 0; t < 9; ++t)
                    if (ok[t])
                        B200CB_POOL_FEED(x[t]);
            } else {
                for (uint32_t fh = fh_lo; fh < fh_hi; ++fh) {
                    const T* rowp = sb + ((size_t)(h0 + (int32_t)fh) * p.iw + w0) * pitch;
                    uint32_t fw = fw_lo;
                    for (; fw + 7 <= fw_hi; fw += 7) {   // seven independent loads in flight, fed in tap order
                        Vec<T, V> x[7];
#pragma unroll
                        for (uint32_t t = 0; t < 7; ++t)
                            x[t] = *reinterpret_cast<const Vec<T, V>*>(rowp + (size_t)(fw + t) * pitch);
#pragma unroll
                        for (uint32_t t = 0; t < 7; ++t)
                            B200CB_POOL_FEED(x[t]);
                    }
                    for (; fw + 2 <= fw_hi; fw += 2) {
                        const Vec<T, V> x0 = *reinterpret_cast<const Vec<T, V>*>(rowp + (size_t)fw * pitch);
                        const Vec<T, V> x1 = *reinterpret_cast<const Vec<T, V>*>(rowp + (size_t)(fw + 1) * pitch);
                        B200CB_POOL_FEED(x0);
                        B200CB_POOL_FEED(x1);
                    }
                    for (; fw < fw_hi; ++fw) {
                        const Vec<T, V> x0 = *reinterpret_cast<const Vec<T, V>*>(rowp + (size_t)fw * pitch);
                        B200CB_POOL_FEED(x0);
                    }
                }
            }
#undef B200CB_POOL_FEED
            if constexpr (WORDS) {
                Vec<T, V> o;
                uint32_t* ow__ = reinterpret_cast<uint32_t*>(&o);
                if constexpr (Q8W) {
#pragma unroll
                    for (int k = 0; k < NW / 2; ++k)
                        ow__[k] = __byte_perm(wacc[2 * k], wacc[2 * k + 1], 0x6420) ^ 0x80808080u;
                } else {
#pragma unroll
                    for (int k = 0; k < NW; ++k)
                        ow__[k] = wacc[k];
                }
                *reinterpret_cast<Vec<T, V>*>(db) = o;
                continue;
            }
            const uint32_t cnt = (fh_hi - fh_lo) * (fw_hi - fw_lo);
            Vec<T, V> o;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float rr;
                if constexpr (MODE == B200CB_POOL_MAX) {
                    rr = acc[k];
                } else if constexpr (DT == B200CB_DT_QS8) {
                    // T(std::round(float(sum) / count)): count is the window size (include) or the fed count
                    const float c = MODE == B200CB_POOL_AVERAGE ? static_cast<float>(p.wh * p.ww) : static_cast<float>(cnt);
                    rr = roundf(__fdiv_rn(static_cast<float>(iacc[k]), c));
                } else {
                    const float c = MODE == B200CB_POOL_AVERAGE ? p.count_all : rnd<DT>(static_cast<float>(cnt));
                    // (the IEEE division is ~10 instructions; a power-of-two count divides exactly by multiplication)
                    rr = (MODE == B200CB_POOL_AVERAGE && p.inv_count != 0.f) ? __fmul_rn(acc[k], p.inv_count) : __fdiv_rn(acc[k], c);
                }
                o.v[k] = Elem<DT>::st(rr);
            }
            *reinterpret_cast<Vec<T, V>*>(db) = o;
        }
    }
}

template <int DT, int MODE>
int pool_launch_v(const b200cb_pool_desc& d, const void* src, void* dst, cudaStream_t stream) {
    using T = typename Elem<DT>::T;
    PoolParams p;
    p.n = d.n, p.c = d.c, p.ih = d.ih, p.iw = d.iw, p.oh = d.oh, p.ow = d.ow;
    p.wh = d.window_h, p.ww = d.window_w, p.sh = d.stride_h, p.sw = d.stride_w, p.ph = d.pad_h, p.pw = d.pad_w;
    const float cnt = (float)(d.window_h * d.window_w);
    p.count_all = DT == B200CB_DT_F16   ? __half2float(__float2half_rn(cnt))   /* window sizes < 2048 are exact */
                  : DT == B200CB_DT_BF16 ? __bfloat162float(__float2bfloat16_rn(cnt))
                                         : cnt;
    {
        int ex = 0;
        p.inv_count = (p.count_all > 0.f && frexpf(p.count_all, &ex) == 0.5f) ? 1.0f / p.count_all : 0.f;
    }
    constexpr int VMAX = 16 / (int)sizeof(T);
    int v = 1;
    if (d.format == B200CB_FORMAT_NHWC) {
        p.layout = 0;
        for (v = VMAX; v > 1 && d.c % v != 0; v /= 2) {
        }
    } else if (d.format == B200CB_FORMAT_NCHW) {
        p.layout = 1;
    } else {
        p.layout = 2;
        v = 4;
    }
    p.cg = d.c / v;
    const uint32_t row_items = p.layout == 0 ? d.ow * p.cg : d.ow;
    p.tpr = 256;
    while (p.tpr > 4 && p.tpr / 2 >= row_items)
        p.tpr /= 2;
    const uint64_t rows = p.layout == 0 ? (uint64_t)d.n * d.oh : (uint64_t)d.n * p.cg * d.oh;
    if (rows > 0x7fffffffull || (uint64_t)row_items > 0x7fffffffull)
        return B200CB_UNSUPPORTED_SHAPE;
    const uint32_t grid = grid_for(rows * p.tpr, 256);
    const T* s = static_cast<const T*>(src);
    T* o = static_cast<T*>(dst);
    switch (v) {
        case 1: pool_kernel<DT, MODE, 1><<<grid, 256, 0, stream>>>(s, o, p); break;
        case 2: pool_kernel<DT, MODE, 2><<<grid, 256, 0, stream>>>(s, o, p); break;
        case 4: pool_kernel<DT, MODE, 4><<<grid, 256, 0, stream>>>(s, o, p); break;
        case 8:
            if constexpr (VMAX >= 8) {
                pool_kernel<DT, MODE, 8><<<grid, 256, 0, stream>>>(s, o, p);
                break;
            }
        case 16:
            if constexpr (VMAX >= 16) {
                pool_kernel<DT, MODE, 16><<<grid, 256, 0, stream>>>(s, o, p);
                break;
            }
        default: return B200CB_UNSUPPORTED_ALIGNMENT;
    }
    return B200CB_OK;
}

template <int DT>
int pool_launch(const b200cb_pool_desc& d, const void* src, void* dst, cudaStream_t stream) {
    switch (d.mode) {
        case B200CB_POOL_MAX: return pool_launch_v<DT, B200CB_POOL_MAX>(d, src, dst, stream);
        case B200CB_POOL_AVERAGE: return pool_launch_v<DT, B200CB_POOL_AVERAGE>(d, src, dst, stream);
        case B200CB_POOL_AVERAGE_COUNT_EXCLUDE_PADDING:
            return pool_launch_v<DT, B200CB_POOL_AVERAGE_COUNT_EXCLUDE_PADDING>(d, src, dst, stream);
        default: return B200CB_UNSUPPORTED_PARAM;
    }
}

// ---------------------------------------------------------------- elementwise
template <int DT>
__device__ __forceinline__ float elem_apply(uint32_t mode, float x, float y) {
    switch (mode) {
        case B200CB_ELEM_RELU: return x <= 0.f ? 0.f : x;
        case B200CB_ELEM_ADD: return rnd<DT>(__fadd_rn(x, y));
        case B200CB_ELEM_MAX: return x > y ? x : y;
        case B200CB_ELEM_MUL: return rnd<DT>(__fmul_rn(x, y));
        case B200CB_ELEM_FUSE_ADD_RELU: {
            const float s = rnd<DT>(__fadd_rn(x, y));
            return s <= 0.f ? 0.f : s;
        }
        default:   // H_SWISH: x * min(max(x + 3, 0.f), 6.f) * (1.f / 6.f) -- the float literals promote the whole
                   // expression to fp32 for the half types too: one rounding at the end (kern_defs.cuh:203)
            return rnd<DT>(act_generic(x, B200CB_NL_H_SWISH));
    }
}

template <int DT>
__global__ void __launch_bounds__(256) elemwise_kernel(
        uint32_t mode, bool binary, uint64_t nvec, uint64_t n, const typename Elem<DT>::T* __restrict__ a,
        const typename Elem<DT>::T* __restrict__ b, typename Elem<DT>::T* __restrict__ dst) {
    using T = typename Elem<DT>::T;
    constexpr int V = 16 / (int)sizeof(T);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const Vec<T, V> x = reinterpret_cast<const Vec<T, V>*>(a)[i];
        Vec<T, V> y = x, o;
        if (binary)
            y = reinterpret_cast<const Vec<T, V>*>(b)[i];
#pragma unroll
        for (int k = 0; k < V; ++k)
            o.v[k] = Elem<DT>::st(elem_apply<DT>(mode, Elem<DT>::ld(x.v[k]), Elem<DT>::ld(y.v[k])));
        reinterpret_cast<Vec<T, V>*>(dst)[i] = o;
    }
    // tail (< one vector)
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * V) {
        const uint64_t i = nvec * V + threadIdx.x;
        dst[i] = Elem<DT>::st(elem_apply<DT>(mode, Elem<DT>::ld(a[i]), Elem<DT>::ld(binary ? b[i] : a[i])));
    }
}

__device__ __forceinline__ float q_apply(uint32_t mode, float x, float y) {
    switch (mode) {
        case B200CB_ELEM_RELU: return x <= 0.f ? 0.f : x;
        case B200CB_ELEM_ADD: return __fadd_rn(x, y);
        case B200CB_ELEM_FUSE_ADD_RELU: {
            const float s = __fadd_rn(x, y);
            return s <= 0.f ? 0.f : s;
        }
        default: return act_generic(x, B200CB_NL_H_SWISH);
    }
}

__global__ void __launch_bounds__(256) elemwise_q_kernel(
        uint32_t mode, bool binary, uint64_t nvec, uint64_t n, const int8_t* __restrict__ a, float sa,
        const int8_t* __restrict__ b, float sb, int8_t* __restrict__ dst, RequantParams rq) {
    // fast_div == 2 (host-proven finite quotients): the conversion-free sequence of the ConvBias epilogue
    // (epilogue.cuh) -- bytes to float through PRMT, unguarded Markstein division, magic-number rounding
    const bool fast = rq.fast_div == 2;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 xv = reinterpret_cast<const uint4*>(a)[i];
        const uint4 yv = binary ? reinterpret_cast<const uint4*>(b)[i] : xv;
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
        uint32_t ow[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t xx = xw[w] ^ 0x80808080u, yy = yw[w] ^ 0x80808080u;
            float xf[4] = {s8_to_float<0>(xx), s8_to_float<1>(xx), s8_to_float<2>(xx), s8_to_float<3>(xx)};
            float yf[4] = {s8_to_float<0>(yy), s8_to_float<1>(yy), s8_to_float<2>(yy), s8_to_float<3>(yy)};
            uint32_t ub[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float r = q_apply(mode, __fmul_rn(xf[k], sa), __fmul_rn(yf[k], sb));
                ub[k] = fast ? round_sat_s8_signed(div_by_scale_noguard(r, rq)) : static_cast<uint32_t>(quant_s8(r, rq));
            }
            ow[w] = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
        }
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * 16) {
        const uint64_t i = nvec * 16 + threadIdx.x;
        const float r = q_apply(mode, __fmul_rn((float)a[i], sa), __fmul_rn((float)(binary ? b[i] : a[i]), sb));
        dst[i] = (int8_t)quant_s8(r, rq);
    }
}

// Binary ADD / FUSE_ADD_RELU (the residual joins of a QuantizedS8 network) on the packed requantize of the ConvBias
// epilogue (epilogue.cuh): two outputs per FMUL2 / FFMA2 / FADD2, the sum RN(RN(a * sa) + RN(b * sb)) as an FMA by a
// run-time 1.0, division / rounding / clamp through requant_tail_*.  Same results as elemwise_q_kernel bit for bit
// (fast_div == 2 only); ~9 instead of ~20 instructions per output.
template <bool RELU>
__global__ void __launch_bounds__(256) elemwise_q_add_kernel(
        uint64_t nvec, uint64_t n, const int8_t* __restrict__ a, float sa, const int8_t* __restrict__ b, float sb,
        int8_t* __restrict__ dst, RequantParams rq) {
    const RequantPacked kp = make_requant_packed(rq);
    const f32x2_t sa2 = f2_pack(sa, sa), sb2 = f2_pack(sb, sb);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 xv = reinterpret_cast<const uint4*>(a)[i];
        const uint4 yv = reinterpret_cast<const uint4*>(b)[i];
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
        uint32_t ow[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t xx = xw[w] ^ 0x80808080u, yy = yw[w] ^ 0x80808080u;
            uint32_t ub[4];
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                // bytes -> the bits of 2^23 + (v + 128); one exact packed add removes the offset
                const f32x2_t xf = f2_add(f2_pack(__uint_as_float(__byte_perm(xx, 0x4B000000u, 0x7540u | k)),
                                                  __uint_as_float(__byte_perm(xx, 0x4B000000u, 0x7540u | (k + 1)))),
                                          kp.neg_zmagic);
                const f32x2_t yf = f2_add(f2_pack(__uint_as_float(__byte_perm(yy, 0x4B000000u, 0x7540u | k)),
                                                  __uint_as_float(__byte_perm(yy, 0x4B000000u, 0x7540u | (k + 1)))),
                                          kp.neg_zmagic);
                const f32x2_t r = f2_fma(f2_mul(xf, sa2), kp.one, f2_mul(yf, sb2));
                if constexpr (RELU)
                    requant_tail_nonneg(r, kp, ub[k], ub[k + 1]);
                else
                    requant_tail_signed(r, kp, ub[k], ub[k + 1]);
            }
            ow[w] = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
        }
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * 16) {
        const uint64_t i = nvec * 16 + threadIdx.x;
        const float r = q_apply(RELU ? B200CB_ELEM_FUSE_ADD_RELU : B200CB_ELEM_ADD, __fmul_rn((float)a[i], sa),
                                __fmul_rn((float)b[i], sb));
        dst[i] = (int8_t)quant_s8(r, rq);
    }
}

// ---------------------------------------------------------------- TypeCvt
// from<T>: quantized -> dequantize (int * scale), normal -> the value; to<T>: quantized -> quantize(float(in)),
// normal -> static_cast (float -> half types round to nearest even)
template <int DT>
struct Raw {
    using T = typename Elem<DT>::T;
};
template <>
struct Raw<B200CB_DT_QS32> {
    using T = int32_t;
};
template <int SDT, int DDT>
__device__ __forceinline__ typename Raw<DDT>::T cvt_one(typename Raw<SDT>::T x, float s_src, const RequantParams& rq) {
    float v;
    if constexpr (SDT == B200CB_DT_QS8 || SDT == B200CB_DT_QS32)
        v = __fmul_rn(static_cast<float>(x), s_src);
    else
        v = Elem<SDT>::ld(x);
    if constexpr (DDT == B200CB_DT_QS8) {
        return (int8_t)quant_s8(v, rq);
    } else if constexpr (DDT == B200CB_DT_QS32) {
        float q = roundf(__fdiv_rn(v, rq.s_dst));
        q = fminf(fmaxf(-2147483648.f, q), 2147483520.f);
        return static_cast<int32_t>(q);
    } else {
        return Elem<DDT>::st(v);
    }
}
// E elements per thread and access: 16 bytes of the WIDER side, so that consecutive lanes touch consecutive bytes on
// both sides (a lane writing 64 contiguous bytes makes every store instruction a 64-byte-strided scatter)
template <int SDT, int DDT>
__global__ void __launch_bounds__(256) typecvt_kernel(
        uint64_t n, const void* __restrict__ src_, void* __restrict__ dst_, float s_src, RequantParams rq) {
    using ST = typename Raw<SDT>::T;
    using DT_ = typename Raw<DDT>::T;
    constexpr int E = 16 / (int)(sizeof(ST) > sizeof(DT_) ? sizeof(ST) : sizeof(DT_));
    const ST* src = static_cast<const ST*>(src_);
    DT_* dst = static_cast<DT_*>(dst_);
    const uint64_t nvec = n / E;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const Vec<ST, E> x = reinterpret_cast<const Vec<ST, E>*>(src)[i];
        Vec<DT_, E> o;
        if constexpr (SDT == B200CB_DT_QS8 && E == 4 && DDT != B200CB_DT_QS8 && DDT != B200CB_DT_QS32) {
            // int8 -> float through PRMT + one exact FADD (epilogue.cuh) instead of four I2F.S8
            const uint32_t w = *reinterpret_cast<const uint32_t*>(&x) ^ 0x80808080u;
            const float f[4] = {s8_to_float<0>(w), s8_to_float<1>(w), s8_to_float<2>(w), s8_to_float<3>(w)};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o.v[k] = Elem<DDT>::st(__fmul_rn(f[k], s_src));
        } else {
#pragma unroll
            for (int k = 0; k < E; ++k)
                o.v[k] = cvt_one<SDT, DDT>(x.v[k], s_src, rq);
        }
        reinterpret_cast<Vec<DT_, E>*>(dst)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * E) {
        const uint64_t i = nvec * E + threadIdx.x;
        dst[i] = cvt_one<SDT, DDT>(src[i], s_src, rq);
    }
}

inline size_t al256(size_t v) {
    return (v + 255) & ~size_t(255);
}
// ---------------------------------------------------------------- deconvolution helpers
// KRSC filter (OC, FH, FW, IC) -> (IC, FH, FW, OC), spatially flipped: the filter of the stride-1 convolution
// over diff that equals the data gradient
template <typename T>
__global__ void deconv_filter_kernel(const T* __restrict__ w, T* __restrict__ o, uint32_t OC, uint32_t FH, uint32_t FW, uint32_t IC) {
    const uint64_t total = (uint64_t)OC * FH * FW * IC;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t oc = i % OC;
        const uint32_t fw = (i / OC) % FW;
        const uint32_t fh = (i / ((uint64_t)OC * FW)) % FH;
        const uint32_t ic = i / ((uint64_t)OC * FW * FH);
        o[i] = w[(((uint64_t)oc * FH + (FH - 1 - fh)) * FW + (FW - 1 - fw)) * IC + ic];
    }
}
// diff (N, OH, OW, C) -> zero-upsampled (N, UH, UW, C): up[n, oh * sh, ow * sw, :] = diff[n, oh, ow, :], 16-byte vectors
__global__ void deconv_upsample_kernel(
        const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t N, uint32_t OH, uint32_t OW, uint32_t UH,
        uint32_t UW, uint32_t cv, uint32_t sh, uint32_t sw) {
    const uint64_t total = (uint64_t)N * UH * UW * cv;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = i % cv;
        const uint32_t uw = (i / cv) % UW;
        const uint32_t uh = (i / ((uint64_t)cv * UW)) % UH;
        const uint32_t n = i / ((uint64_t)cv * UW * UH);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (uh % sh == 0 && uw % sw == 0 && uh / sh < OH && uw / sw < OW)
            v = in[(((uint64_t)n * OH + uh / sh) * OW + uw / sw) * cv + c];
        out[i] = v;
    }
}

// ---- stride > 1 (dilation 1): phase decomposition.  Grad rows ih with (ih + pad) % s == a only receive filter rows
// fh = a + s*j, so each of the s*s phases is a dense stride-1 correlation of diff with a J_a x J_b sub-filter
// (J_a = ceil((F - a) / s)), written to a temporary (OH + J_a - 1) x (OW + J_b - 1) image; one interleave pass
// scatters the temporaries into grad.  No multiply by an inserted zero is ever issued.
struct PhaseGeom {
    uint32_t s_h, s_w, nph;            // phases = s_h * s_w (<= 16)
    uint32_t jh[4], jw[4];             // taps per phase along h / w (0: the phase receives nothing)
    uint32_t th[4], tw[4];             // temporary extent per phase
    size_t flt_off[16], tmp_off[16];   // workspace offsets
    size_t flt_total, tmp_total;
    b200cb_desc inner[16];
};
int phase_geometry(const b200cb_desc& f, PhaseGeom* g) {
    if (f.stride_h > 4 || f.stride_w > 4 || f.dilate_h != 1 || f.dilate_w != 1)
        return B200CB_UNSUPPORTED_PARAM;
    g->s_h = f.stride_h, g->s_w = f.stride_w, g->nph = f.stride_h * f.stride_w;
    for (uint32_t a = 0; a < f.stride_h; ++a) {
        g->jh[a] = a < f.fh ? (f.fh - a + f.stride_h - 1) / f.stride_h : 0;
        g->th[a] = g->jh[a] ? f.oh + g->jh[a] - 1 : 0;
    }
    for (uint32_t b = 0; b < f.stride_w; ++b) {
        g->jw[b] = b < f.fw ? (f.fw - b + f.stride_w - 1) / f.stride_w : 0;
        g->tw[b] = g->jw[b] ? f.ow + g->jw[b] - 1 : 0;
    }
    size_t fo = 0, to = 0;
    for (uint32_t a = 0; a < f.stride_h; ++a)
        for (uint32_t b = 0; b < f.stride_w; ++b) {
            const uint32_t ph = a * f.stride_w + b;
            g->flt_off[ph] = fo, g->tmp_off[ph] = to;
            b200cb_desc& e = g->inner[ph];
            memset(&e, 0, sizeof(e));
            if (!g->jh[a] || !g->jw[b])
                continue;
            e.format = B200CB_FORMAT_NHWC, e.sparse = B200CB_SPARSE_DENSE, e.mode = B200CB_MODE_CROSS_CORRELATION;
            e.nonline_mode = B200CB_NL_IDENTITY, e.compute_mode = f.compute_mode, e.bias_mode = B200CB_BIAS_NONE;
            e.src_dtype = f.dst_dtype, e.filter_dtype = f.filter_dtype, e.dst_dtype = f.src_dtype;
            e.bias_dtype = e.z_dtype = B200CB_DT_NONE;
            e.src_scale = f.dst_scale, e.filter_scale = f.filter_scale, e.dst_scale = f.src_scale;
            e.n = f.n, e.ic = f.oc, e.ih = f.oh, e.iw = f.ow, e.oc = f.ic, e.oh = g->th[a], e.ow = g->tw[b];
            e.fh = g->jh[a], e.fw = g->jw[b], e.group = 1;
            e.pad_h = g->jh[a] - 1, e.pad_w = g->jw[b] - 1, e.stride_h = e.stride_w = 1, e.dilate_h = e.dilate_w = 1;
            fo += al256((size_t)f.oc * g->jh[a] * g->jw[b] * f.ic * dtype_size(f.filter_dtype));
            to += al256((size_t)f.n * g->th[a] * g->tw[b] * f.ic * dtype_size(f.src_dtype));
        }
    g->flt_total = fo, g->tmp_total = to;
    return B200CB_OK;
}
// sub-filter of phase (a, b): k[ic][jh'][jw'][oc] = w_eff[oc][a + s_h (J_a - 1 - jh')][b + s_w (J_b - 1 - jw')][ic],
// w_eff = w (CROSS_CORRELATION) or w spatially flipped (CONVOLUTION)
template <typename T>
__global__ void deconv_phase_filter_kernel(
        const T* __restrict__ w, T* __restrict__ o, uint32_t OC, uint32_t FH, uint32_t FW, uint32_t IC, uint32_t a,
        uint32_t b, uint32_t sh, uint32_t sw, uint32_t JH, uint32_t JW, uint32_t flip) {
    const uint64_t total = (uint64_t)IC * JH * JW * OC;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t oc = i % OC;
        const uint32_t jw = (i / OC) % JW;
        const uint32_t jh = (i / ((uint64_t)OC * JW)) % JH;
        const uint32_t ic = i / ((uint64_t)OC * JW * JH);
        uint32_t fh = a + sh * (JH - 1 - jh), fw = b + sw * (JW - 1 - jw);
        if (flip)
            fh = FH - 1 - fh, fw = FW - 1 - fw;
        o[i] = w[(((uint64_t)oc * FH + fh) * FW + fw) * IC + ic];
    }
}
struct PhasePtrs {
    const uint4* tmp[16];
    uint32_t th[4], tw[4];
};
// grad[n, ih, iw, :] = temporary of phase ((ih + p) % s_h, (iw + p) % s_w) at ((ih + p) / s_h, (iw + p) / s_w);
// rows / columns no diff element reaches (beyond the temporary) are zero
__global__ void deconv_interleave_kernel(
        PhasePtrs pp, uint4* __restrict__ grad, uint32_t N, uint32_t IH, uint32_t IW, uint32_t cv, uint32_t sh,
        uint32_t sw, uint32_t ph, uint32_t pw) {
    const uint64_t total = (uint64_t)N * IH * IW * cv;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = i % cv;
        const uint32_t iw = (i / cv) % IW;
        const uint32_t ih = (i / ((uint64_t)cv * IW)) % IH;
        const uint32_t n = i / ((uint64_t)cv * IW * IH);
        const uint32_t a = (ih + ph) % sh, y = (ih + ph) / sh, b = (iw + pw) % sw, x = (iw + pw) / sw;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (y < pp.th[a] && x < pp.tw[b])
            v = pp.tmp[a * sw + b][(((uint64_t)n * pp.th[a] + y) * pp.tw[b] + x) * cv + c];
        grad[i] = v;
    }
}
bool use_phases(const b200cb_desc& f) {
    return (f.stride_h > 1 || f.stride_w > 1) && f.dilate_h == 1 && f.dilate_w == 1 && f.stride_h <= 4 && f.stride_w <= 4 &&
           (f.ic * dtype_size(f.src_dtype)) % 16 == 0;
}

constexpr uint32_t KERNEL_DECONV = 16, KERNEL_BATCH_CONVBIAS = 17;

struct DeconvGeom {
    b200cb_desc inner;       // the stride-1 forward convolution over (upsampled) diff
    uint32_t uh, uw;         // upsampled diff extent (== oh, ow for stride 1)
    bool upsample;
    size_t flt_bytes, up_bytes;
};
int deconv_geometry(const b200cb_desc& f, DeconvGeom* g) {
    if (f.format != B200CB_FORMAT_NHWC || f.sparse != B200CB_SPARSE_DENSE || f.group != 1)
        return B200CB_UNSUPPORTED_FORMAT;
    const uint32_t eh = (f.fh - 1) * f.dilate_h, ew = (f.fw - 1) * f.dilate_w;
    if (f.pad_h > eh || f.pad_w > ew)
        return B200CB_UNSUPPORTED_PARAM;   // (negative padding of the equivalent convolution: cropping, not implemented)
    if (f.stride_h == 0 || f.stride_w == 0)
        return B200CB_UNSUPPORTED_PARAM;
    // diff extent must be what the forward convolution produces from grad's extent
    if ((f.ih + 2 * f.pad_h < eh + 1) || (f.iw + 2 * f.pad_w < ew + 1) ||
        (f.ih + 2 * f.pad_h - eh - 1) / f.stride_h + 1 != f.oh || (f.iw + 2 * f.pad_w - ew - 1) / f.stride_w + 1 != f.ow)
        return B200CB_UNSUPPORTED_SHAPE;
    g->upsample = f.stride_h > 1 || f.stride_w > 1;
    g->uh = f.ih + 2 * f.pad_h - eh;   // == (oh - 1) * stride + 1 + the rows the forward stride skipped
    g->uw = f.iw + 2 * f.pad_w - ew;
    b200cb_desc e;
    memset(&e, 0, sizeof(e));
    e.format = B200CB_FORMAT_NHWC, e.sparse = B200CB_SPARSE_DENSE, e.mode = f.mode;
    e.nonline_mode = B200CB_NL_IDENTITY, e.compute_mode = f.compute_mode;
    e.bias_mode = B200CB_BIAS_NONE, e.has_z = 0;
    e.src_dtype = f.dst_dtype, e.filter_dtype = f.filter_dtype, e.dst_dtype = f.src_dtype;   // diff -> grad
    e.bias_dtype = e.z_dtype = B200CB_DT_NONE;
    e.src_scale = f.dst_scale, e.filter_scale = f.filter_scale, e.dst_scale = f.src_scale;
    e.n = f.n, e.ic = f.oc, e.ih = g->uh, e.iw = g->uw;
    e.oc = f.ic, e.oh = f.ih, e.ow = f.iw;
    e.fh = f.fh, e.fw = f.fw, e.group = 1;
    e.pad_h = eh - f.pad_h, e.pad_w = ew - f.pad_w;
    e.stride_h = e.stride_w = 1;
    e.dilate_h = f.dilate_h, e.dilate_w = f.dilate_w;
    g->inner = e;
    const size_t es = dtype_size(f.filter_dtype);
    g->flt_bytes = al256((size_t)f.oc * f.fh * f.fw * f.ic * es);
    g->up_bytes = g->upsample ? al256((size_t)f.n * g->uh * g->uw * f.oc * dtype_size(f.dst_dtype)) : 0;
    return B200CB_OK;
}

// the kernels move 16-byte vectors: tensors must start on 16-byte boundaries (megdnn allocations are 256-byte
// aligned; an offset view is reported as unsupported rather than faulting)
inline bool aligned16(const void* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

int cuda_rc(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error(what, cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

}  // namespace
}  // namespace b200cb

using namespace b200cb;

extern "C" {

int b200cb_pooling(const b200cb_pool_desc* d, const void* src, void* dst, void* stream) {
    if (!d || !src || !dst)
        return B200CB_ERR_NULL;
    if (d->window_h == 0 || d->window_w == 0 || d->stride_h == 0 || d->stride_w == 0 || d->n == 0 || d->c == 0)
        return B200CB_UNSUPPORTED_PARAM;
    if (d->ih + 2 * d->pad_h < d->window_h || d->iw + 2 * d->pad_w < d->window_w ||
        (d->ih + 2 * d->pad_h - d->window_h) / d->stride_h + 1 != d->oh ||
        (d->iw + 2 * d->pad_w - d->window_w) / d->stride_w + 1 != d->ow)
        return B200CB_UNSUPPORTED_SHAPE;
    // every window must see at least one pixel (the reference throws "The pooling window lies outside completely")
    if (d->pad_h >= d->window_h || d->pad_w >= d->window_w ||
        (d->oh - 1) * d->stride_h >= d->ih + d->pad_h || (d->ow - 1) * d->stride_w >= d->iw + d->pad_w)
        return B200CB_UNSUPPORTED_PARAM;
    if (d->format == B200CB_FORMAT_NCHW4 && (d->dtype != B200CB_DT_QS8 || d->c % 4 != 0))
        return B200CB_UNSUPPORTED_FORMAT;
    if (d->format != B200CB_FORMAT_NHWC && d->format != B200CB_FORMAT_NCHW && d->format != B200CB_FORMAT_NCHW4)
        return B200CB_UNSUPPORTED_FORMAT;
    if (!aligned16(src) || !aligned16(dst))
        return B200CB_UNSUPPORTED_ALIGNMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc;
    switch (d->dtype) {
        case B200CB_DT_F32: rc = pool_launch<B200CB_DT_F32>(*d, src, dst, s); break;
        case B200CB_DT_F16: rc = pool_launch<B200CB_DT_F16>(*d, src, dst, s); break;
        case B200CB_DT_BF16: rc = pool_launch<B200CB_DT_BF16>(*d, src, dst, s); break;
        case B200CB_DT_QS8: rc = pool_launch<B200CB_DT_QS8>(*d, src, dst, s); break;
        default: return B200CB_UNSUPPORTED_DTYPE;
    }
    return rc != B200CB_OK ? rc : cuda_rc("pooling launch");
}

int b200cb_elemwise(uint32_t mode, uint32_t dtype, uint64_t n, const void* a, const void* b, void* dst, void* stream) {
    const bool binary = mode == B200CB_ELEM_ADD || mode == B200CB_ELEM_MAX || mode == B200CB_ELEM_MUL ||
                        mode == B200CB_ELEM_FUSE_ADD_RELU;
    if (!binary && mode != B200CB_ELEM_RELU && mode != B200CB_ELEM_H_SWISH)
        return B200CB_UNSUPPORTED_PARAM;
    if (!a || !dst || (binary && !b))
        return B200CB_ERR_NULL;
    if (n == 0)
        return B200CB_OK;
    if (!aligned16(a) || !aligned16(dst) || (binary && !aligned16(b)))
        return B200CB_UNSUPPORTED_ALIGNMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t es = dtype_size(dtype);
    if (es != 2 && es != 4)
        return B200CB_UNSUPPORTED_DTYPE;
    const uint64_t nvec = n / (16 / es);
    const uint32_t grid = grid_for(nvec ? nvec : 1, 256);
    switch (dtype) {
        case B200CB_DT_F32:
            elemwise_kernel<B200CB_DT_F32><<<grid, 256, 0, s>>>(
                    mode, binary, nvec, n, static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(dst));
            break;
        case B200CB_DT_F16:
            elemwise_kernel<B200CB_DT_F16><<<grid, 256, 0, s>>>(
                    mode, binary, nvec, n, static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(dst));
            break;
        case B200CB_DT_BF16:
            elemwise_kernel<B200CB_DT_BF16><<<grid, 256, 0, s>>>(
                    mode, binary, nvec, n, static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
                    static_cast<__nv_bfloat16*>(dst));
            break;
        default: return B200CB_UNSUPPORTED_DTYPE;
    }
    return cuda_rc("elemwise launch");
}

int b200cb_elemwise_q(
        uint32_t mode, uint64_t n, const void* a, float sa, const void* b, float sb, void* dst, float sd, void* stream) {
    const bool binary = mode == B200CB_ELEM_ADD || mode == B200CB_ELEM_FUSE_ADD_RELU;
    if (!binary && mode != B200CB_ELEM_RELU && mode != B200CB_ELEM_H_SWISH)
        return B200CB_UNSUPPORTED_PARAM;
    if (!a || !dst || (binary && !b))
        return B200CB_ERR_NULL;
    if (n == 0)
        return B200CB_OK;
    if (!aligned16(a) || !aligned16(dst) || (binary && !aligned16(b)))
        return B200CB_UNSUPPORTED_ALIGNMENT;
    b200cb_desc t;
    memset(&t, 0, sizeof(t));
    t.dst_scale = sd, t.src_scale = t.filter_scale = 1.f, t.bias_dtype = B200CB_DT_NONE;
    RequantParams rq = make_requant(t);
    if (rq.fast_div) {
        // |op(a * sa, b * sb)| <= 128 (|sa| + |sb|) (RELU / H_SWISH do not amplify): the unguarded division is
        // exact whenever that bound over s_dst stays far below overflow (the test of make_requant)
        const double rmax = 128.0 * (fabs((double)sa) + fabs((double)sb));
        rq.fast_div = (isfinite(rmax) && rmax * (double)rq.inv_dst < 1152921504606846976.0) ? 2u : 1u;
    }
    const uint64_t nvec = n / 16;
    if (binary && rq.fast_div == 2 && (mode == B200CB_ELEM_ADD || mode == B200CB_ELEM_FUSE_ADD_RELU)) {
        const uint32_t grid = grid_for(nvec ? nvec : 1, 256);
        if (mode == B200CB_ELEM_ADD)
            elemwise_q_add_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                    nvec, n, static_cast<const int8_t*>(a), sa, static_cast<const int8_t*>(b), sb,
                    static_cast<int8_t*>(dst), rq);
        else
            elemwise_q_add_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                    nvec, n, static_cast<const int8_t*>(a), sa, static_cast<const int8_t*>(b), sb,
                    static_cast<int8_t*>(dst), rq);
        return cuda_rc("elemwise_q_add launch");
    }
    elemwise_q_kernel<<<grid_for(nvec ? nvec : 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            mode, binary, nvec, n, static_cast<const int8_t*>(a), sa, static_cast<const int8_t*>(b), sb,
            static_cast<int8_t*>(dst), rq);
    return cuda_rc("elemwise_q launch");
}

int b200cb_typecvt(
        uint32_t sdt, float s_src, uint32_t ddt, float s_dst, uint64_t n, const void* src, void* dst, void* stream) {
    if (!src || !dst)
        return B200CB_ERR_NULL;
    if (n == 0)
        return B200CB_OK;
    if (!aligned16(src) || !aligned16(dst))
        return B200CB_UNSUPPORTED_ALIGNMENT;
    b200cb_desc t;
    memset(&t, 0, sizeof(t));
    t.dst_scale = s_dst, t.src_scale = t.filter_scale = 1.f, t.bias_dtype = B200CB_DT_NONE;
    RequantParams rq = make_requant(t);
    if (rq.fast_div == 2)
        rq.fast_div = 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t grid = grid_for(n / 4 + 1, 256);   // (>= one vector per thread for every pair)
#define CVT(S_, D_)                                                                      \
    if (sdt == S_ && ddt == D_) {                                                        \
        typecvt_kernel<S_, D_><<<grid, 256, 0, s>>>(n, src, dst, s_src, rq);             \
        return cuda_rc("typecvt launch");                                                \
    }
    CVT(B200CB_DT_F32, B200CB_DT_F16) CVT(B200CB_DT_F32, B200CB_DT_BF16) CVT(B200CB_DT_F32, B200CB_DT_QS8)
    CVT(B200CB_DT_F32, B200CB_DT_QS32) CVT(B200CB_DT_F32, B200CB_DT_F32)
    CVT(B200CB_DT_F16, B200CB_DT_F32) CVT(B200CB_DT_BF16, B200CB_DT_F32) CVT(B200CB_DT_F16, B200CB_DT_QS8)
    CVT(B200CB_DT_BF16, B200CB_DT_QS8)
    CVT(B200CB_DT_QS8, B200CB_DT_F32) CVT(B200CB_DT_QS8, B200CB_DT_F16) CVT(B200CB_DT_QS8, B200CB_DT_BF16)
    CVT(B200CB_DT_QS8, B200CB_DT_QS8) CVT(B200CB_DT_QS8, B200CB_DT_QS32)
    CVT(B200CB_DT_QS32, B200CB_DT_QS8) CVT(B200CB_DT_QS32, B200CB_DT_F32) CVT(B200CB_DT_QS32, B200CB_DT_QS32)
#undef CVT
    return B200CB_UNSUPPORTED_DTYPE;
}

// ---- ConvolutionForward: ConvBias without bias / z / activation (the reference's CUDA ConvolutionForwardImpl
// builds the same ConvBias operator internally)
int b200cb_convolution_query(const b200cb_desc* desc, b200cb_plan* plan) {
    if (!desc || !plan)
        return B200CB_ERR_NULL;
    b200cb_desc d = *desc;
    d.bias_mode = B200CB_BIAS_NONE, d.bias_dtype = B200CB_DT_NONE;
    d.has_z = 0, d.z_dtype = B200CB_DT_NONE;
    d.nonline_mode = B200CB_NL_IDENTITY;
    return b200cb_query(&d, plan);
}
int b200cb_convolution_launch(
        const b200cb_plan* plan, const void* src, const void* filter, void* dst, void* workspace, size_t ws_bytes,
        void* stream) {
    if (!plan)
        return B200CB_ERR_NULL;
    if (plan->desc.bias_mode != B200CB_BIAS_NONE || plan->desc.has_z)
        return B200CB_ERR_BAD_PLAN;
    return b200cb_launch(plan, src, filter, nullptr, nullptr, dst, workspace, ws_bytes, stream);
}

// ---- ConvolutionBackwardData
int b200cb_deconv_query(const b200cb_desc* fwd, b200cb_plan* plan) {
    if (!fwd || !plan)
        return B200CB_ERR_NULL;
    DeconvGeom g;
    int rc = deconv_geometry(*fwd, &g);
    if (rc != B200CB_OK)
        return rc;
    b200cb_plan inner;
    uint64_t ws = 0;
    bool phases = use_phases(*fwd);
    if (phases) {
        PhaseGeom pg;
        if (phase_geometry(*fwd, &pg) != B200CB_OK)
            phases = false;
        uint64_t inner_ws = 0;
        for (uint32_t ph = 0; phases && ph < pg.nph; ++ph) {
            if (!pg.inner[ph].n)
                continue;
            if (b200cb_query(&pg.inner[ph], &inner) != B200CB_OK)
                phases = false;   // (a phase the ConvBias kernels reject: fall back to the zero-upsampled form)
            else
                inner_ws = inner.workspace_bytes > inner_ws ? inner.workspace_bytes : inner_ws;
        }
        if (phases)
            ws = pg.flt_total + pg.tmp_total + inner_ws;
    }
    if (!phases) {
        rc = b200cb_query(&g.inner, &inner);
        if (rc != B200CB_OK)
            return rc;
        ws = g.flt_bytes + g.up_bytes + inner.workspace_bytes;
    }
    *plan = inner;
    plan->kernel = KERNEL_DECONV;
    plan->desc = *fwd;
    plan->workspace_bytes = ws;
    plan->reserved[0] = phases ? 1u : 0u;
    // accounting of the gradient itself
    plan->flops = 2ull * fwd->n * fwd->oc * fwd->oh * fwd->ow * fwd->ic * fwd->fh * fwd->fw;
    plan->bytes = (uint64_t)fwd->n * fwd->oh * fwd->ow * fwd->oc * dtype_size(fwd->dst_dtype) +
                  (uint64_t)fwd->oc * fwd->fh * fwd->fw * fwd->ic * dtype_size(fwd->filter_dtype) +
                  (uint64_t)fwd->n * fwd->ih * fwd->iw * fwd->ic * dtype_size(fwd->src_dtype);
    return B200CB_OK;
}
int b200cb_deconv_launch(
        const b200cb_plan* plan, const void* filter, const void* diff, void* grad, void* workspace, size_t ws_bytes,
        void* stream) {
    if (!plan || !filter || !diff || !grad)
        return B200CB_ERR_NULL;
    if (plan->magic != PLAN_MAGIC || plan->kernel != KERNEL_DECONV)
        return B200CB_ERR_BAD_PLAN;
    const b200cb_desc& f = plan->desc;
    DeconvGeom g;
    if (deconv_geometry(f, &g) != B200CB_OK)
        return B200CB_ERR_BAD_PLAN;
    b200cb_plan inner;
    if (plan->reserved[0] == 1) {
        PhaseGeom pg;
        if (phase_geometry(f, &pg) != B200CB_OK || !workspace || ws_bytes < plan->workspace_bytes)
            return pg.nph ? B200CB_ERR_WORKSPACE : B200CB_ERR_BAD_PLAN;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        uint8_t* ws = static_cast<uint8_t*>(workspace);
        uint8_t* inner_ws = ws + pg.flt_total + pg.tmp_total;
        const uint32_t es = dtype_size(f.filter_dtype);
        PhasePtrs pp;
        memset(&pp, 0, sizeof(pp));
        for (uint32_t a = 0; a < pg.s_h; ++a)
            pp.th[a] = pg.th[a];
        for (uint32_t b = 0; b < pg.s_w; ++b)
            pp.tw[b] = pg.tw[b];
        for (uint32_t a = 0; a < pg.s_h; ++a)
            for (uint32_t b = 0; b < pg.s_w; ++b) {
                const uint32_t ph = a * pg.s_w + b;
                if (!pg.inner[ph].n)
                    continue;
                void* k = ws + pg.flt_off[ph];
                void* t = ws + pg.flt_total + pg.tmp_off[ph];
                pp.tmp[ph] = static_cast<const uint4*>(t);
                const uint64_t fl = (uint64_t)f.oc * pg.jh[a] * pg.jw[b] * f.ic;
                const uint32_t flip = f.mode == B200CB_MODE_CONVOLUTION;
                if (es == 1)
                    deconv_phase_filter_kernel<uint8_t><<<grid_for(fl, 256), 256, 0, s>>>(
                            static_cast<const uint8_t*>(filter), static_cast<uint8_t*>(k), f.oc, f.fh, f.fw, f.ic, a, b,
                            pg.s_h, pg.s_w, pg.jh[a], pg.jw[b], flip);
                else if (es == 2)
                    deconv_phase_filter_kernel<uint16_t><<<grid_for(fl, 256), 256, 0, s>>>(
                            static_cast<const uint16_t*>(filter), static_cast<uint16_t*>(k), f.oc, f.fh, f.fw, f.ic, a, b,
                            pg.s_h, pg.s_w, pg.jh[a], pg.jw[b], flip);
                else
                    deconv_phase_filter_kernel<uint32_t><<<grid_for(fl, 256), 256, 0, s>>>(
                            static_cast<const uint32_t*>(filter), static_cast<uint32_t*>(k), f.oc, f.fh, f.fw, f.ic, a, b,
                            pg.s_h, pg.s_w, pg.jh[a], pg.jw[b], flip);
                int rc = cuda_rc("deconv phase filter");
                if (rc != B200CB_OK)
                    return rc;
                if (b200cb_query(&pg.inner[ph], &inner) != B200CB_OK)
                    return B200CB_ERR_BAD_PLAN;
                rc = b200cb_launch(&inner, diff, k, nullptr, nullptr, t, inner_ws, inner.workspace_bytes, stream);
                if (rc != B200CB_OK)
                    return rc;
            }
        const uint32_t cv = f.ic * dtype_size(f.src_dtype) / 16;
        const uint64_t items = (uint64_t)f.n * f.ih * f.iw * cv;
        deconv_interleave_kernel<<<grid_for(items, 256), 256, 0, s>>>(
                pp, static_cast<uint4*>(grad), f.n, f.ih, f.iw, cv, pg.s_h, pg.s_w, f.pad_h, f.pad_w);
        return cuda_rc("deconv interleave");
    }
    if (b200cb_query_algo(&g.inner, plan->algo, &inner) != B200CB_OK)
        return B200CB_ERR_BAD_PLAN;
    if (!workspace || ws_bytes < g.flt_bytes + g.up_bytes + inner.workspace_bytes)
        return B200CB_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    void* wt = ws;
    void* up = ws + g.flt_bytes;
    void* inner_ws = ws + g.flt_bytes + g.up_bytes;
    const uint64_t fl = (uint64_t)f.oc * f.fh * f.fw * f.ic;
    const uint32_t es = dtype_size(f.filter_dtype);
    if (es == 1)
        deconv_filter_kernel<uint8_t><<<grid_for(fl, 256), 256, 0, s>>>(
                static_cast<const uint8_t*>(filter), static_cast<uint8_t*>(wt), f.oc, f.fh, f.fw, f.ic);
    else if (es == 2)
        deconv_filter_kernel<uint16_t><<<grid_for(fl, 256), 256, 0, s>>>(
                static_cast<const uint16_t*>(filter), static_cast<uint16_t*>(wt), f.oc, f.fh, f.fw, f.ic);
    else
        deconv_filter_kernel<uint32_t><<<grid_for(fl, 256), 256, 0, s>>>(
                static_cast<const uint32_t*>(filter), static_cast<uint32_t*>(wt), f.oc, f.fh, f.fw, f.ic);
    int rc = cuda_rc("deconv filter transform");
    if (rc != B200CB_OK)
        return rc;
    const void* a = diff;
    if (g.upsample) {
        const uint32_t row = f.oc * dtype_size(f.dst_dtype);
        if (row % 16 != 0)
            return B200CB_ERR_BAD_PLAN;   // (the inner query already demands 16-byte pixels)
        const uint64_t items = (uint64_t)f.n * g.uh * g.uw * (row / 16);
        deconv_upsample_kernel<<<grid_for(items, 256), 256, 0, s>>>(
                static_cast<const uint4*>(diff), static_cast<uint4*>(up), f.n, f.oh, f.ow, g.uh, g.uw, row / 16,
                f.stride_h, f.stride_w);
        rc = cuda_rc("deconv upsample");
        if (rc != B200CB_OK)
            return rc;
        a = up;
    }
    return b200cb_launch(&inner, a, wt, nullptr, nullptr, grad, inner_ws, inner.workspace_bytes, stream);
}

// ---- BatchConvBiasForward: per-image filters, one ConvBias launch per image
int b200cb_batch_convbias_query(const b200cb_desc* desc, b200cb_plan* plan) {
    if (!desc || !plan)
        return B200CB_ERR_NULL;
    b200cb_desc one = *desc;
    one.n = 1;
    one.src_batch_stride = one.z_batch_stride = one.dst_batch_stride = 0;
    const int rc = b200cb_query(&one, plan);
    if (rc != B200CB_OK)
        return rc;
    plan->kernel = KERNEL_BATCH_CONVBIAS;
    plan->desc = *desc;
    plan->flops *= desc->n;
    plan->bytes = plan->bytes * desc->n;
    return B200CB_OK;
}
int b200cb_batch_convbias_launch(
        const b200cb_plan* plan, const void* src, const void* filter, const void* bias, const void* z, void* dst,
        void* workspace, size_t ws_bytes, void* stream) {
    if (!plan || !src || !filter || !dst)
        return B200CB_ERR_NULL;
    if (plan->magic != PLAN_MAGIC || plan->kernel != KERNEL_BATCH_CONVBIAS)
        return B200CB_ERR_BAD_PLAN;
    const b200cb_desc& d = plan->desc;
    b200cb_desc one = d;
    one.n = 1;
    one.src_batch_stride = one.z_batch_stride = one.dst_batch_stride = 0;
    b200cb_plan inner;
    if (b200cb_query_algo(&one, plan->algo, &inner) != B200CB_OK)
        return B200CB_ERR_BAD_PLAN;
    const size_t src_img = (d.src_batch_stride ? d.src_batch_stride : (size_t)d.ih * d.iw * d.ic) * dtype_size(d.src_dtype);
    const size_t dst_img = (d.dst_batch_stride ? d.dst_batch_stride : (size_t)d.oh * d.ow * d.oc) * dtype_size(d.dst_dtype);
    const size_t z_img = (d.z_batch_stride ? d.z_batch_stride : (size_t)d.oh * d.ow * d.oc) * dtype_size(d.z_dtype);
    const size_t flt_img = (size_t)d.oc * (d.ic / d.group) * d.fh * d.fw * dtype_size(d.filter_dtype);
    for (uint32_t i = 0; i < d.n; ++i) {
        const int rc = b200cb_launch(
                &inner, static_cast<const uint8_t*>(src) + i * src_img, static_cast<const uint8_t*>(filter) + i * flt_img,
                bias, z ? static_cast<const uint8_t*>(z) + i * z_img : nullptr, static_cast<uint8_t*>(dst) + i * dst_img,
                workspace, ws_bytes, stream);
        if (rc != B200CB_OK)
            return rc;
    }
    return B200CB_OK;
}

}  // extern "C"

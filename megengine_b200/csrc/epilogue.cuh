// Epilogue arithmetic shared by the implicit-GEMM and depthwise kernels.
// Restates the naive oracle's z-add / activation / requantize sequence
// (dnn/src/naive/conv_bias/opr_impl.cpp:21-98, dnn/include/megdnn/dtype.h:845-854,888-897,
//  dnn/src/common/elemwise/kern_defs.cuh:176,203); see SURVEY.md Appendix A.3/A.4.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "b200cb.h"

namespace b200cb {

// H_SWISH / SIGMOID (the rare modes); RELU / IDENTITY are handled branch-free by the callers.
__device__ __forceinline__ float act_generic(float x, uint32_t nl) {
    if (nl == B200CB_NL_H_SWISH)
        return __fmul_rn(__fmul_rn(x, fminf(fmaxf(__fadd_rn(x, 3.f), 0.f), 6.f)), 1.f / 6.f);
    if (nl == B200CB_NL_SIGMOID)
        return 1.f / (expf(-x) + 1.f);
    return x;
}

__device__ __forceinline__ float act_f(float x, uint32_t nl) {
    if (nl == B200CB_NL_RELU)
        return x <= 0.f ? 0.f : x;
    if (nl == B200CB_NL_IDENTITY)
        return x;
    return act_generic(x, nl);
}

// Scalars of the QuantizedS8 requantize epilogue, prepared once on the host.
struct RequantParams {
    float s_bias;   // scale of the QuantizedS32 bias / accumulator tensor
    float s_z;      // scale of z
    float s_dst;    // scale of dst
    float inv_dst;  // RN(1 / s_dst) when `fast_div`, unused otherwise
    uint32_t fast_div;
};

// host: fills RequantParams from the descriptor's scales (defined in igemm_sm100.cu)
RequantParams make_requant(const b200cb_desc& d);

// r / s, correctly rounded (== IEEE fp32 division, what the oracle's `x / scale` computes).
// Fast path: Markstein's sequence with y = RN(1/s) precomputed -- q0 = RN(r*y); e = r - q0*s (exact,
// one FMA); q = RN(q0 + e*y) is the correctly rounded quotient whenever nothing overflows or
// underflows.  Quotients too large for that guarantee saturate in the clamp that follows anyway, so
// q0 is returned for them; b200cb_selftest(0) sweeps all 2^32 dividends against __fdiv_rn on the GPU.
__device__ __forceinline__ float div_by_scale(float r, const RequantParams& q) {
    if (q.fast_div) {
        const float q0 = __fmul_rn(r, q.inv_dst);
        const float e = __fmaf_rn(-q0, q.s_dst, r);
        const float q1 = __fmaf_rn(e, q.inv_dst, q0);
        return fabsf(q0) < 1048576.f ? q1 : q0;
    }
    return __fdiv_rn(r, q.s_dst);
}

// QuantizedS8 requantize, bit-exact restatement of the oracle: every step is one
// IEEE fp32 RN operation, no FMA contraction across steps, true division, roundf = half away.
//   r = float(acc + bias) * s_bias  [+ float(z) * s_z]  ->  activation  ->  sat_s8(roundf(r / s_dst))
// Without z the callers pass zq == 0 with s_z == 0: r + (+0) == r (a -0 becomes +0, which rounds to
// the same int8).
__device__ __forceinline__ float dequant_acc(int32_t sfb, int32_t zq, const RequantParams& q) {
    const float r = __fmul_rn(__int2float_rn(sfb), q.s_bias);
    return __fadd_rn(r, __fmul_rn(__int2float_rn(zq), q.s_z));
}
__device__ __forceinline__ int32_t quant_s8(float r, const RequantParams& q) {
    float v = roundf(div_by_scale(r, q));
    v = fminf(fmaxf(-128.f, v), 127.f);
    return __float2int_rn(v);
}
// ---- the same arithmetic on the FMA / ALU pipes only ------------------------------------------
// The generic sequence above costs four conversion-pipe instructions per element (I2F.S8, FRND,
// F2I, ...), and that pipe issues 16 lanes / clk / SM: it, not HBM, bounded the int8 epilogue.  The
// helpers below produce bit-identical results with FADD / FMNMX / PRMT / LOP3 only; they require
// `fast_div == 2` (host-checked: every quotient is finite, see make_requant).

// byte `b` of `w_xor80` (an int8x4 word XOR 0x80808080) -> exact float of the signed byte:
// PRMT builds the bits of 2^23 + (z + 128), one exact FADD removes the offset.
template <int B>
__device__ __forceinline__ float s8_to_float(uint32_t w_xor80) {
    return __uint_as_float(__byte_perm(w_xor80, 0x4B000000u, 0x7540u | B)) - 8388736.f;
}
// r / s_dst without the overflow guard (fast_div == 2)
__device__ __forceinline__ float div_by_scale_noguard(float r, const RequantParams& q) {
    const float q0 = __fmul_rn(r, q.inv_dst);
    const float e = __fmaf_rn(-q0, q.s_dst, r);
    return __fmaf_rn(e, q.inv_dst, q0);
}
// sat_s8(roundf(v)) for v >= 0 (RELU output), result in the LOW BYTE of the return value.
// roundf(v) == floor(v + 0.5) for v >= 0; the add is done round-toward-zero so that it can never
// round up across an integer, and 2^23 + t rounded down leaves floor(t) in the low mantissa bits.
__device__ __forceinline__ uint32_t round_sat_s8_nonneg(float v) {
    const float c = fminf(v, 127.f);
    const float t = __fadd_rz(c, 0.5f);
    return __float_as_uint(__fadd_rd(t, 8388608.f));
}
// any sign: magnitude as above, then a two's-complement negate of the low byte
__device__ __forceinline__ uint32_t round_sat_s8_signed(float v) {
    const float c = fminf(fmaxf(v, -128.f), 127.f);
    const float t = __fadd_rz(fabsf(c), 0.5f);
    const uint32_t m = __float_as_uint(__fadd_rd(t, 8388608.f));
    const uint32_t s = static_cast<uint32_t>(static_cast<int32_t>(__float_as_uint(c)) >> 31);
    return (m ^ s) - s;
}
// low bytes of four words -> one int8x4 word
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

__device__ __forceinline__ int32_t requant_s8(int32_t sfb, int32_t zq, uint32_t nl, const RequantParams& q) {
    return quant_s8(act_f(dequant_acc(sfb, zq, q), nl), q);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float f16lo(uint32_t v) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xffff)));
}
__device__ __forceinline__ float f16hi(uint32_t v) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16)));
}

}  // namespace b200cb

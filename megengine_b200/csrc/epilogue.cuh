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
    float one;      // 1.0f as a RUN-TIME value (see the z dequantisation of requant_pair_*)
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
// ---- the same rounding + clamp through ONE saturating conversion (B200CB_RQ_CVT): the int8 epilogue is bound by
// the ALU pipe and the issue slots while the conversion pipe idles (ncu: ALU 57-69 %, XU 3 %), so the clamp
// (FMNMX) and the second magic add move there.  roundf(v) = floor(v + 0.5) for v >= 0 and trunc(v -+ 0.5) in
// general; the add is round-toward-zero (can never cross an integer), cvt.{rmi,rzi}.sat.s8 floors / truncates
// exactly and saturates to [-128, 127] like the oracle's fmin(fmax(-128, roundf(v)), 127).
#ifndef B200CB_RQ_CVT
#define B200CB_RQ_CVT 1
#endif
__device__ __forceinline__ uint32_t cvt_floor_sat_s8(float t) {
    int32_t r;
    asm("cvt.rmi.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(t));
    return static_cast<uint32_t>(r);
}
__device__ __forceinline__ uint32_t cvt_trunc_sat_s8(float t) {
    int32_t r;
    asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(t));
    return static_cast<uint32_t>(r);
}
__device__ __forceinline__ uint32_t round_sat_s8_nonneg_cvt(float v) {
    return cvt_floor_sat_s8(__fadd_rz(v, 0.5f));
}
__device__ __forceinline__ uint32_t round_sat_s8_signed_cvt(float v) {
    const float h = __uint_as_float(0x3f000000u | (__float_as_uint(v) & 0x80000000u));   // copysign(0.5, v)
    return cvt_trunc_sat_s8(__fadd_rz(v, h));
}
// low bytes of four words -> one int8x4 word
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// ---- packed fp32x2 forms (sm_100 FMUL2 / FFMA2 / FADD2 with explicit rounding): two elements per
// instruction on the FMA pipe, each lane rounded exactly as the scalar op
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t f2_add_rz(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t f2_add_rm(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// loop-invariant operands of the packed requantize
struct RequantPacked {
    f32x2_t s_bias, inv, neg_s, half, magic, neg_i2f, neg_2p23, c4096, s_z, one, neg_zmagic;
};
// int32 -> float without the conversion pipe (I2F issues 16 lanes / clk / SM and was the busiest pipe of the
// int8 epilogue): for v in [-2^22, 2^22) the integer add 0x4B400000 + v yields the bits of the float
// 2^23 + 2^22 + v, and one exact FADD of -(2^23 + 2^22) leaves float(v).  The callers fold the constant into
// the bias they add anyway (`MAGIC` forms below) after proving the range: |acc| <= K * 2^14 from the
// descriptor, |bias| from the data (igemm: per bias slice; depthwise: per thread).
constexpr uint32_t I2F_MAGIC_BITS = 0x4B400000u;
constexpr float I2F_MAGIC = 12582912.f;
constexpr uint32_t I2F_RANGE = 1u << 22;
// A product that feeds an add must stay a separately rounded product -- and ptxas (12.9) contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 although the explicit .rn forbids that for the scalar
// forms; --fmad=false does not stop it, and fma(a, b, -0.0) is simplified back and fused as well
// (b200cb_selftest(2) caught the resulting 1-in-2e7 off-by-one outputs).  So the only mul -> add
// adjacency of the sequence, the dequantisation with z, stays on scalar mul.rn.f32 / add.rn.f32 (never
// contracted); everything after it has no product feeding an add: FMUL2 -> {FMNMX | FMUL2 | FFMA2
// multiplicand}, explicit FFMA2s, FADD2.RZ -> FADD2.RM.
// The dequantisation with z itself, x = RN(RN(a * s_bias) + RN(z * s_z)), is packed too: the sum is issued as
// fma(RN(a * s_bias), one, RN(z * s_z)) with `one` == 1.0f arriving as a kernel parameter -- an FMA by a run-time 1 rounds
// exactly like the add, and ptxas can neither fold a multiplier it does not know nor contract a product into an
// instruction that already multiplies.  2 FMUL2 + 1 FFMA2 + 1 FADD2 per output PAIR instead of 4 FMUL + 4 FADD (the
// z layers of the int8 ResNet-50 issued 405 instructions per 32 outputs, 128 of them these scalars; ncu instruction
// profile profiles/r02/ncu_inst_qs8_l1c3.txt).  b200cb_selftest(2) sweeps it against the scalar form.
__device__ __forceinline__ RequantPacked make_requant_packed(const RequantParams& q) {
    RequantPacked r;
    r.s_bias = f2_pack(q.s_bias, q.s_bias);
    r.s_z = f2_pack(q.s_z, q.s_z);
    r.one = f2_pack(q.one, q.one);
    r.neg_zmagic = f2_pack(-8388736.f, -8388736.f);
    r.inv = f2_pack(q.inv_dst, q.inv_dst);
    r.neg_s = f2_pack(-q.s_dst, -q.s_dst);
    r.half = f2_pack(0.5f, 0.5f);
    r.magic = f2_pack(8388608.f, 8388608.f);
    r.neg_i2f = f2_pack(-I2F_MAGIC, -I2F_MAGIC);
    r.neg_2p23 = f2_pack(-8388608.f, -8388608.f);
    r.c4096 = f2_pack(4096.f, 4096.f);
    return r;
}
// Any int32 -> float, correctly rounded (== I2F.RN), without the conversion pipe: v = hi * 4096 + lo with
// hi = v >> 12 (|hi| < 2^19) and lo = v & 4095 -- both convert exactly through magic-number adds, and one FMA
// fma(float(hi), 4096, float(lo)) rounds the exact value v once.  3 ALU + 1.5 packed FMA-pipe instructions per
// element; b200cb_selftest(2) sweeps all 2^32 accumulators against the I2F form.
// Measured, not deduced: in the rows-mode epilogue (int8 ResNet-50 l1.c2: nine short MMAs per 8192 outputs) this
// form is 17 % faster than I2F (0.087 -> 0.072 ms; ncu's `xu_realtime` triage metric read 92 % there while
// `sm__inst_executed_pipe_xu` read 3 %, so the mechanism -- conversion latency in the dependent chain rather than
// pipe occupancy -- is an inference); in every other layer the extra instructions cost more than the conversion did
// (qs8 step 2.75 -> 2.90 ms when used everywhere: those epilogues are ALU / issue bound -- ALU pipe 57-69 %, issue
// slots 71-77 % busy, profiles/r02/ncu_pipes_qs8_layer1.txt).  Only the rows-mode epilogue selects it.
__device__ __forceinline__ f32x2_t i2f_pair(int32_t v0, int32_t v1, const RequantPacked& k) {
    const f32x2_t hi = f2_add(f2_pack(__int_as_float((v0 >> 12) + static_cast<int32_t>(I2F_MAGIC_BITS)),
                                      __int_as_float((v1 >> 12) + static_cast<int32_t>(I2F_MAGIC_BITS))), k.neg_i2f);
    const f32x2_t lo = f2_add(f2_pack(__uint_as_float((static_cast<uint32_t>(v0) & 0xfffu) | 0x4B000000u),
                                      __uint_as_float((static_cast<uint32_t>(v1) & 0xfffu) | 0x4B000000u)), k.neg_2p23);
    return f2_fma(hi, k.c4096, lo);
}

// the part of requant_pair_* after the dequantisation: x = two dequantised values (RN-rounded floats), results in
// the low bytes of lo / hi.  Shared with the quantized elementwise ADD kernels (nnops.cu).
__device__ __forceinline__ void requant_tail_nonneg(f32x2_t x, const RequantPacked& k, uint32_t& lo, uint32_t& hi) {
    float x0, x1;
    f2_unpack(x, x0, x1);
    x = f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
    const f32x2_t q0 = f2_mul(x, k.inv);
    const f32x2_t e = f2_fma(q0, k.neg_s, x);   // x - q0 * s, exact
    const f32x2_t q1 = f2_fma(e, k.inv, q0);
#if B200CB_RQ_CVT
    float t0, t1;
    f2_unpack(f2_add_rz(q1, k.half), t0, t1);
    lo = cvt_floor_sat_s8(t0);
    hi = cvt_floor_sat_s8(t1);
#else
    float c0, c1;
    f2_unpack(q1, c0, c1);
    const f32x2_t u = f2_add_rm(f2_add_rz(f2_pack(fminf(c0, 127.f), fminf(c1, 127.f)), k.half), k.magic);
    float u0, u1;
    f2_unpack(u, u0, u1);
    lo = __float_as_uint(u0);
    hi = __float_as_uint(u1);
#endif
}
__device__ __forceinline__ void requant_tail_signed(f32x2_t x, const RequantPacked& k, uint32_t& lo, uint32_t& hi) {
    const f32x2_t q0 = f2_mul(x, k.inv);
    const f32x2_t e = f2_fma(q0, k.neg_s, x);
    const f32x2_t q1 = f2_fma(e, k.inv, q0);
    float c0, c1;
    f2_unpack(q1, c0, c1);
#if B200CB_RQ_CVT
    const f32x2_t hs = f2_pack(__uint_as_float(0x3f000000u | (__float_as_uint(c0) & 0x80000000u)),
                               __uint_as_float(0x3f000000u | (__float_as_uint(c1) & 0x80000000u)));
    float t0, t1;
    f2_unpack(f2_add_rz(q1, hs), t0, t1);
    lo = cvt_trunc_sat_s8(t0);
    hi = cvt_trunc_sat_s8(t1);
#else
    c0 = fminf(fmaxf(c0, -128.f), 127.f);
    c1 = fminf(fmaxf(c1, -128.f), 127.f);
    const f32x2_t u = f2_add_rm(f2_add_rz(f2_pack(fabsf(c0), fabsf(c1)), k.half), k.magic);
    float u0, u1;
    f2_unpack(u, u0, u1);
    const uint32_t s0 = static_cast<uint32_t>(static_cast<int32_t>(__float_as_uint(c0)) >> 31);
    const uint32_t s1 = static_cast<uint32_t>(static_cast<int32_t>(__float_as_uint(c1)) >> 31);
    lo = (__float_as_uint(u0) ^ s0) - s0;
    hi = (__float_as_uint(u1) ^ s1) - s1;
#endif
}

// two RELU outputs at once: (acc + bias) pairs -> the two result bytes in the low bytes of lo / hi.
// Same operation sequence as dequant + fmaxf + div_by_scale_noguard + round_sat_s8_nonneg.
// CVT selects the int32 -> float conversion: 0 = I2F (conversion pipe), 1 = sfb0 / sfb1 already carry
// I2F_MAGIC_BITS (acc + bias + magic, range proven by the caller), 2 = i2f_pair (full range, FMA / ALU pipes only)
template <bool HAS_Z, int CVT = 0>
__device__ __forceinline__ void requant_pair_nonneg(
        int32_t sfb0, int32_t sfb1, uint32_t zbits0, uint32_t zbits1, const RequantPacked& k, uint32_t& lo,
        uint32_t& hi) {
    f32x2_t x;
    f32x2_t a;
    if constexpr (CVT == 1)
        a = f2_add(f2_pack(__int_as_float(sfb0), __int_as_float(sfb1)), k.neg_i2f);   // exact
    else if constexpr (CVT == 2)
        a = i2f_pair(sfb0, sfb1, k);
    else
        a = f2_pack(__int2float_rn(sfb0), __int2float_rn(sfb1));
    if constexpr (HAS_Z) {
        const f32x2_t zf = f2_add(f2_pack(__uint_as_float(zbits0), __uint_as_float(zbits1)), k.neg_zmagic);  // exact
        x = f2_fma(f2_mul(a, k.s_bias), k.one, f2_mul(zf, k.s_z));
    } else {
        x = f2_mul(a, k.s_bias);
    }
    requant_tail_nonneg(x, k, lo, hi);
}

// the IDENTITY (any sign) form of the same: clamp to [-128, 127], round the magnitude, negate the byte
template <bool HAS_Z, int CVT = 0>
__device__ __forceinline__ void requant_pair_signed(
        int32_t sfb0, int32_t sfb1, uint32_t zbits0, uint32_t zbits1, const RequantPacked& k, uint32_t& lo,
        uint32_t& hi) {
    f32x2_t x;
    f32x2_t a;
    if constexpr (CVT == 1)
        a = f2_add(f2_pack(__int_as_float(sfb0), __int_as_float(sfb1)), k.neg_i2f);   // exact
    else if constexpr (CVT == 2)
        a = i2f_pair(sfb0, sfb1, k);
    else
        a = f2_pack(__int2float_rn(sfb0), __int2float_rn(sfb1));
    if constexpr (HAS_Z) {
        const f32x2_t zf = f2_add(f2_pack(__uint_as_float(zbits0), __uint_as_float(zbits1)), k.neg_zmagic);  // exact
        x = f2_fma(f2_mul(a, k.s_bias), k.one, f2_mul(zf, k.s_z));
    } else {
        x = f2_mul(a, k.s_bias);
    }
    requant_tail_signed(x, k, lo, hi);
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
// RELU on a packed pair of 16-bit outputs: max(x, 0), NaN propagated (the oracle's `x <= 0 ? 0 : x` keeps it)
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
    uint32_t r;
    asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0u));
    return r;
}
__device__ __forceinline__ uint32_t relu_f16x2(uint32_t v) {
    uint32_t r;
    asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0u));
    return r;
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

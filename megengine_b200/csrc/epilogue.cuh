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

__device__ __forceinline__ float act_f(float x, uint32_t nl) {
    if (nl == B200CB_NL_RELU)
        return x <= 0.f ? 0.f : x;
    if (nl == B200CB_NL_H_SWISH)
        return __fmul_rn(__fmul_rn(x, fminf(fmaxf(__fadd_rn(x, 3.f), 0.f), 6.f)), 1.f / 6.f);
    if (nl == B200CB_NL_SIGMOID)
        return 1.f / (expf(-x) + 1.f);
    return x;
}

// QuantizedS8 requantize, bit-exact restatement of the oracle: every step is one
// IEEE fp32 RN operation, no FMA contraction, true division, roundf = half away from zero.
__device__ __forceinline__ int32_t requant_s8(
        int32_t sfb, float s_bias, bool has_z, int32_t zq, float s_z, uint32_t nl, float s_dst) {
    float r = __fmul_rn(__int2float_rn(sfb), s_bias);
    if (has_z)
        r = __fadd_rn(r, __fmul_rn(__int2float_rn(zq), s_z));
    r = act_f(r, nl);
    float q = roundf(__fdiv_rn(r, s_dst));
    q = fminf(fmaxf(-128.f, q), 127.f);
    return __float2int_rn(q);
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

// Layout pre-pass for the NCHW4 int8 cells of ConvBias: the tcgen05 implicit GEMM
// wants K (channels) contiguous per pixel, NCHW4 keeps only 4 channels together, and a TMA box cannot
// be narrower than 16 bytes -- so src (N, C/4, H, W, 4) and filter (OC, IC/4, FH, FW, 4) are
// re-tiled once into NHWC / KRSC in the workspace, channels zero-padded to a multiple of 16.
// dst and z stay NCHW4: the epilogue of igemm_sm100.cu reads / writes them in place.
//
// One kernel serves both: a batched transpose of 32-bit words (4 channels),
//   in [b][c4][p]  ->  out [b][p][c4_pad],   zero for c4 >= C4.
// Pure HBM traffic: 32 x 32 word tiles through shared memory, both sides 128-byte coalesced.
// Reference semantics: megdnn RelayoutFormat (dnn/src/cuda/relayout_format/), which the graph
// optimizer inserts around NCHW4 operators (src/gopt/impl/tensor_reformat.cpp).
#include <string.h>
#include "common.h"

namespace b200cb {
namespace {

__global__ void __launch_bounds__(256) transpose_words_kernel(
        const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t C4, uint32_t C4pad,
        uint32_t P, uint64_t in_batch_words) {
    __shared__ uint32_t tile[32][33];
    const uint32_t b = blockIdx.z;
    const uint32_t p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t* src = in + static_cast<size_t>(b) * in_batch_words;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t c4 = c0 + ty + 8 * i, p = p0 + tx;
        tile[ty + 8 * i][tx] = (c4 < C4 && p < P) ? __ldg(src + static_cast<size_t>(c4) * P + p) : 0u;
    }
    __syncthreads();
    uint32_t* dst = out + static_cast<size_t>(b) * P * C4pad;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t p = p0 + ty + 8 * i, c4 = c0 + tx;
        if (p < P && c4 < C4pad)
            dst[static_cast<size_t>(p) * C4pad + c4] = tile[tx][ty + 8 * i];
    }
}

// Few channel groups (network stems: C4pad <= 8): one thread per pixel gathers its words from the C4
// planes (each plane read is coalesced across the warp) and writes them as one contiguous run.
template <int C4PAD>
__global__ void __launch_bounds__(256) gather_words_kernel(
        const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t C4, uint32_t P,
        uint64_t in_batch_words) {
    const uint32_t b = blockIdx.y;
    const uint32_t p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P)
        return;
    const uint32_t* src = in + static_cast<size_t>(b) * in_batch_words + p;
    uint32_t w[C4PAD];
#pragma unroll
    for (int c = 0; c < C4PAD; ++c)
        w[c] = c < (int)C4 ? __ldg(src + static_cast<size_t>(c) * P) : 0u;
    uint32_t* dst = out + (static_cast<size_t>(b) * P + p) * C4PAD;
#pragma unroll
    for (int c = 0; c < C4PAD; ++c)
        dst[c] = w[c];
}

// Batched transpose of 16-bit elements, in[b][r][s] -> out[b][s][r], rows of the output zero-padded up
// to out_ld (NCHW <-> NHWC of fp16 / bf16 tensors).  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) transpose_u16_kernel(
        const uint16_t* __restrict__ in, uint16_t* __restrict__ out, uint32_t R, uint32_t S, uint32_t out_ld,
        uint64_t in_batch, uint64_t out_batch) {
    __shared__ uint16_t tile[32][34];
    const uint32_t b = blockIdx.z;
    const uint32_t s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint16_t* src = in + static_cast<size_t>(b) * in_batch;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t r = r0 + ty + 8 * i, sidx = s0 + tx;
        tile[ty + 8 * i][tx] = (r < R && sidx < S) ? src[static_cast<size_t>(r) * S + sidx] : uint16_t(0);
    }
    __syncthreads();
    uint16_t* dst = out + static_cast<size_t>(b) * out_batch;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t sidx = s0 + ty + 8 * i, r = r0 + tx;
        if (sidx < S && r < out_ld)
            dst[static_cast<size_t>(sidx) * out_ld + r] = tile[tx][ty + 8 * i];
    }
}

// NHWC channel padding (gopt's PaddingChannelPass on the network input, src/gopt/impl/padding_channel.cpp):
// in [p][C] -> out [p][Cpad], zero for c >= C.  One thread per 16 output bytes; E = element type.
template <typename E>
__global__ void __launch_bounds__(256) pad_channels_kernel(
        const E* __restrict__ in, E* __restrict__ out, uint64_t total_out, uint32_t C, uint32_t Cpad) {
    constexpr uint32_t V = 16 / sizeof(E);
    const uint64_t nvec = (total_out + V - 1) / V;
    for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < nvec;
         t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t i0 = t * V;
        uint64_t pix = i0 / Cpad;
        uint32_t c = static_cast<uint32_t>(i0 - pix * Cpad);
        E v[V];
#pragma unroll
        for (uint32_t k = 0; k < V; ++k) {
            v[k] = (c < C && i0 + k < total_out) ? in[pix * C + c] : E(0);
            if (++c == Cpad)
                c = 0, ++pix;
        }
        if (i0 + V <= total_out) {
            *reinterpret_cast<uint4*>(out + i0) = *reinterpret_cast<const uint4*>(v);
        } else {
            for (uint32_t k = 0; i0 + k < total_out; ++k)
                out[i0 + k] = v[k];
        }
    }
}

// Batched transpose of W-byte words through shared memory: in[b][r][s] -> out[b][s][r'], r' < out_ld,
// zero words for r >= R (channel padding) -- both sides coalesced.  W = 1, 2, 4, 8 or 16 bytes: plain
// elements (NCHW <-> NHWC) or the 4-element groups of NCHW4 (NCHW4 <-> NHWC).
template <typename W>
__global__ void __launch_bounds__(256) transpose_any_kernel(
        const W* __restrict__ in, W* __restrict__ out, uint32_t R, uint32_t S, uint32_t out_ld) {
    __shared__ W tile[32][32 + (sizeof(W) >= 4 ? 1 : 4 / sizeof(W))];
    const uint32_t b = blockIdx.z;
    const uint32_t s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const W* src = in + static_cast<size_t>(b) * R * S;
    W zero;
    memset(&zero, 0, sizeof(W));
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t r = r0 + ty + 8 * i, sidx = s0 + tx;
        tile[ty + 8 * i][tx] = (r < R && sidx < S) ? src[static_cast<size_t>(r) * S + sidx] : zero;
    }
    __syncthreads();
    W* dst = out + static_cast<size_t>(b) * S * out_ld;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t sidx = s0 + ty + 8 * i, r = r0 + tx;
        if (sidx < S && r < out_ld)
            dst[static_cast<size_t>(sidx) * out_ld + r] = tile[tx][ty + 8 * i];
    }
}

// NCHW <-> NCHW4 and the small-channel NHWC -> NCHW4 case: one thread per (image, channel group, pixel);
// the plane side is coalesced across the warp, the grouped side is 4 contiguous elements per thread.
//   DIR 0: (n,c,p) -> (n,c4o,p,4)    DIR 1: (n,c4,p,4) -> (n,co,p)    DIR 2: (n,p,c) -> (n,c4o,p,4)
template <typename E, int DIR>
__global__ void __launch_bounds__(256) group4_kernel(
        const E* __restrict__ in, E* __restrict__ out, uint32_t N, uint32_t C, uint32_t P, uint32_t Cout) {
    const uint32_t groups = ((DIR == 1 ? C : Cout) + 3) / 4;
    const uint64_t total = static_cast<uint64_t>(N) * groups * P;
    for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t p = static_cast<uint32_t>(t % P);
        const uint32_t g = static_cast<uint32_t>((t / P) % groups);
        const uint64_t n = t / (static_cast<uint64_t>(P) * groups);
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            const uint32_t c = g * 4 + k;
            if (DIR == 0) {
                out[((n * groups + g) * P + p) * 4 + k] = c < C ? in[(n * C + c) * P + p] : E(0);
            } else if (DIR == 1) {
                if (c < Cout)
                    out[(n * Cout + c) * P + p] = in[((n * groups + g) * P + p) * 4 + k];
            } else {
                out[((n * groups + g) * P + p) * 4 + k] = c < C ? in[(n * P + p) * C + c] : E(0);
            }
        }
    }
}

template <typename W>
int launch_transpose_any(const void* in, void* out, uint32_t B, uint32_t R, uint32_t S, uint32_t out_ld, cudaStream_t st) {
    dim3 grid((S + 31) / 32, (out_ld + 31) / 32, B);
    if (grid.y > 65535 || grid.z > 65535)
        return B200CB_UNSUPPORTED_SHAPE;
    transpose_any_kernel<W><<<grid, 256, 0, st>>>(static_cast<const W*>(in), static_cast<W*>(out), R, S, out_ld);
    return B200CB_OK;
}
int transpose_words(uint32_t wbytes, const void* in, void* out, uint32_t B, uint32_t R, uint32_t S, uint32_t out_ld,
                    cudaStream_t st) {
    switch (wbytes) {
        case 1: return launch_transpose_any<uint8_t>(in, out, B, R, S, out_ld, st);
        case 2: return launch_transpose_any<uint16_t>(in, out, B, R, S, out_ld, st);
        case 4: return launch_transpose_any<uint32_t>(in, out, B, R, S, out_ld, st);
        case 8: return launch_transpose_any<uint2>(in, out, B, R, S, out_ld, st);
        case 16: return launch_transpose_any<uint4>(in, out, B, R, S, out_ld, st);
        default: return B200CB_UNSUPPORTED_DTYPE;
    }
}
template <int DIR>
int launch_group4(uint32_t es, const void* in, void* out, uint32_t N, uint32_t C, uint32_t P, uint32_t Cout, cudaStream_t st) {
    const uint64_t total = static_cast<uint64_t>(N) * (((DIR == 1 ? C : Cout) + 3) / 4) * P;
    uint64_t blocks = (total + 255) / 256;
    const uint64_t cap = static_cast<uint64_t>(device_sm_count()) * 32;
    if (blocks > cap)
        blocks = cap;
    const dim3 grid(static_cast<uint32_t>(blocks), 1, 1);
    if (es == 1)
        group4_kernel<uint8_t, DIR><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), N, C, P, Cout);
    else if (es == 2)
        group4_kernel<uint16_t, DIR><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), N, C, P, Cout);
    else if (es == 4)
        group4_kernel<uint32_t, DIR><<<grid, 256, 0, st>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), N, C, P, Cout);
    else
        return B200CB_UNSUPPORTED_DTYPE;
    return B200CB_OK;
}

// Float32 ConvBias operand split (capi.cpp, f32 adapter): x = x0 + x1 + r with x0 = bf16(x), x1 = bf16(x - x0),
// |r| <= 2^-17 |x|.  in [p][C] fp32 -> out [p][3 * Cp] bf16, channel blocks (x0, x0, x1) for the activations
// and (x0, x1, x0) for the filter, so that the bf16 dot product over 3 * Cp channels is
// sum a0 b0 + a0 b1 + a1 b0 = sum a b (1 + O(2^-16)), accumulated in fp32 by the tensor core.
__global__ void __launch_bounds__(256) split3_kernel(
        const float* __restrict__ in, uint16_t* __restrict__ out, uint64_t P, uint32_t C, uint32_t Cp, int filter_side) {
    const uint64_t total = P * Cp;
    for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t p = t / Cp;
        const uint32_t c = static_cast<uint32_t>(t - p * Cp);
        const float x = c < C ? in[p * C + c] : 0.f;
        const uint32_t xb = __float_as_uint(x);
        // round to nearest even on the top 16 bits (finite inputs; inf / nan keep their class)
        const uint32_t r0 = ((xb & 0x7f800000u) == 0x7f800000u) ? xb : xb + 0x7fffu + ((xb >> 16) & 1u);
        const uint16_t h0 = static_cast<uint16_t>(r0 >> 16);
        const float rem = x - __uint_as_float(static_cast<uint32_t>(h0) << 16);
        const uint32_t rb = __float_as_uint(rem);
        const uint32_t r1 = ((rb & 0x7f800000u) == 0x7f800000u) ? rb : rb + 0x7fffu + ((rb >> 16) & 1u);
        const uint16_t h1 = static_cast<uint16_t>(r1 >> 16);
        uint16_t* o = out + p * 3 * Cp + c;
        o[0] = h0;
        o[Cp] = filter_side ? h1 : h0;
        o[2 * Cp] = filter_side ? h0 : h1;
    }
}

}  // namespace

int relayout_split3(const void* in, void* out, uint64_t P, uint32_t C, uint32_t Cp, int filter_side, cudaStream_t stream) {
    if (P == 0 || C == 0 || Cp < C)
        return B200CB_UNSUPPORTED_SHAPE;
    uint64_t blocks = (P * Cp + 255) / 256;
    const uint64_t cap = static_cast<uint64_t>(device_sm_count()) * 32;
    if (blocks > cap)
        blocks = cap;
    split3_kernel<<<static_cast<uint32_t>(blocks), 256, 0, stream>>>(
            static_cast<const float*>(in), static_cast<uint16_t*>(out), P, C, Cp, filter_side);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("split3 launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

// b200cb_relayout: see include/b200cb.h.  Every mode is pure HBM traffic (read src once, write dst once).
int relayout_generic(
        uint32_t mode, uint32_t es, uint32_t n, uint32_t c, uint32_t h, uint32_t w, uint32_t c_out, const void* src,
        void* dst, cudaStream_t stream) {
    if (es != 1 && es != 2 && es != 4)
        return B200CB_UNSUPPORTED_DTYPE;
    const uint64_t p64 = static_cast<uint64_t>(h) * w;
    if (p64 > 0x7fffffffull)
        return B200CB_UNSUPPORTED_SHAPE;
    const uint32_t P = static_cast<uint32_t>(p64);
    int rc;
    switch (mode) {
        case B200CB_RELAYOUT_NCHW_NHWC:   // (n, c, P) -> (n, P, c_out), c_out >= c
            if (c_out < c)
                return B200CB_UNSUPPORTED_SHAPE;
            rc = transpose_words(es, src, dst, n, c, P, c_out, stream);
            break;
        case B200CB_RELAYOUT_NHWC_NCHW:   // (n, P, c) -> (n, c_out, P): padded or cropped channel planes
            if (c_out != c)
                return B200CB_UNSUPPORTED_SHAPE;  // (same channel count only: planes are written whole)
            rc = transpose_words(es, src, dst, n, P, c, P, stream);
            break;
        case B200CB_RELAYOUT_NCHW_NCHW4:
            if (c_out % 4 || c_out < c)
                return B200CB_UNSUPPORTED_SHAPE;
            rc = launch_group4<0>(es, src, dst, n, c, P, c_out, stream);
            break;
        case B200CB_RELAYOUT_NCHW4_NCHW:
            if (c % 4 || c_out > c)
                return B200CB_UNSUPPORTED_SHAPE;
            rc = launch_group4<1>(es, src, dst, n, c, P, c_out, stream);
            break;
        case B200CB_RELAYOUT_NCHW4_NHWC:  // words of 4 elements: (n, c/4, P) -> (n, P, c_out/4)
            if (c % 4 || c_out % 4 || c_out < c)
                return B200CB_UNSUPPORTED_SHAPE;
            rc = transpose_words(4 * es, src, dst, n, c / 4, P, c_out / 4, stream);
            break;
        case B200CB_RELAYOUT_NHWC_NCHW4:
            if (c_out % 4 || c_out < c)
                return B200CB_UNSUPPORTED_SHAPE;
            if (c % 4 == 0 && c_out == c)
                rc = transpose_words(4 * es, src, dst, n, P, c / 4, P, stream);
            else
                rc = launch_group4<2>(es, src, dst, n, c, P, c_out, stream);  // small / ragged channel counts
            break;
        case B200CB_RELAYOUT_NHWC_PAD_C:
            return relayout_pad_channels(src, dst, static_cast<uint64_t>(n) * P, c, c_out, es, stream);
        default:
            return B200CB_UNSUPPORTED_PARAM;
    }
    if (rc != B200CB_OK)
        return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("relayout launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

int relayout_pad_channels(
        const void* in, void* out, uint64_t pixels, uint32_t C, uint32_t Cpad, uint32_t es, cudaStream_t stream) {
    if (pixels == 0 || C == 0 || Cpad < C || (es != 1 && es != 2 && es != 4))
        return B200CB_UNSUPPORTED_SHAPE;
    const uint64_t total = pixels * Cpad;
    const uint64_t nvec = (total + 16 / es - 1) / (16 / es);
    uint64_t blocks = (nvec + 255) / 256;
    const uint64_t cap = static_cast<uint64_t>(device_sm_count()) * 32;
    if (blocks > cap)
        blocks = cap;
    const dim3 grid(static_cast<uint32_t>(blocks), 1, 1);
    if (es == 1)
        pad_channels_kernel<uint8_t><<<grid, 256, 0, stream>>>(
                static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), total, C, Cpad);
    else if (es == 2)
        pad_channels_kernel<uint16_t><<<grid, 256, 0, stream>>>(
                static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), total, C, Cpad);
    else
        pad_channels_kernel<uint32_t><<<grid, 256, 0, stream>>>(
                static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), total, C, Cpad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("pad_channels launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

int relayout_transpose16(
        const void* in, void* out, uint32_t B, uint32_t R, uint32_t S, uint32_t out_ld, uint64_t in_batch,
        uint64_t out_batch, cudaStream_t stream) {
    if (B == 0 || R == 0 || S == 0 || out_ld < R)
        return B200CB_UNSUPPORTED_SHAPE;
    dim3 grid((S + 31) / 32, (out_ld + 31) / 32, B);
    if (grid.y > 65535 || grid.z > 65535)
        return B200CB_UNSUPPORTED_SHAPE;
    transpose_u16_kernel<<<grid, 256, 0, stream>>>(
            static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), R, S, out_ld,
            in_batch ? in_batch : static_cast<uint64_t>(R) * S, out_batch ? out_batch : static_cast<uint64_t>(S) * out_ld);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("relayout launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

// (B, C/4, P, 4) int8 -> (B, P, Cpad) int8; `in_batch_elems` = elements between consecutive b (0 = dense)
int relayout_c4_to_channels_last(
        const void* in, void* out, uint32_t B, uint32_t C, uint32_t Cpad, uint32_t P,
        uint64_t in_batch_elems, cudaStream_t stream) {
    if (C % 4 || Cpad % 4 || Cpad < C || B == 0 || P == 0)
        return B200CB_UNSUPPORTED_SHAPE;
    const uint32_t C4 = C / 4, C4pad = Cpad / 4;
    const uint64_t bw = in_batch_elems ? in_batch_elems / 4 : static_cast<uint64_t>(C4) * P;
    if ((C4pad == 2 || C4pad == 4) && B <= 65535) {
        dim3 g((P + 255) / 256, B, 1);
        if (C4pad == 2)
            gather_words_kernel<2><<<g, 256, 0, stream>>>(
                    static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), C4, P, bw);
        else
            gather_words_kernel<4><<<g, 256, 0, stream>>>(
                    static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), C4, P, bw);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_last_error("relayout launch", cudaGetErrorString(e));
            return B200CB_ERR_CUDA;
        }
        return B200CB_OK;
    }
    dim3 grid((P + 31) / 32, (C4pad + 31) / 32, B);
    if (grid.z > 65535 || grid.y > 65535)
        return B200CB_UNSUPPORTED_SHAPE;
    transpose_words_kernel<<<grid, 256, 0, stream>>>(
            static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), C4, C4pad, P, bw);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("relayout launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

}  // namespace b200cb

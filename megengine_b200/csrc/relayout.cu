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

}  // namespace

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

// Internal declarations shared by the C-ABI layer and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "b200cb.h"

namespace b200cb {

constexpr uint32_t PLAN_MAGIC = 0xB200CB01u;

// kernel families (b200cb_plan::kernel)
enum : uint32_t { KERNEL_IGEMM = 1, KERNEL_DWCONV = 2 };

// element kinds of the implicit-GEMM kernel
enum : int { KIND_BF16 = 0, KIND_F16 = 1, KIND_S8 = 2 };

constexpr uint32_t MAX_DYN_SMEM = 232448;  // 227 KB opt-in limit per CTA on sm_100
constexpr uint32_t IGEMM_BLOCK_M = 128;
constexpr uint32_t IGEMM_KBYTES = 128;     // one k-block = one 128-byte swizzle row per pixel
constexpr uint32_t IGEMM_EPI_BUF_BYTES = 16384;
constexpr uint32_t IGEMM_MAX_STAGES = 16;

void set_last_error(const char* what, const char* detail);

// ---- driver entry points (resolved lazily through the runtime; no libcuda link dependency)
CUresult encode_tiled(
        CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* gaddr,
        const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
        const uint32_t* elem_strides, CUtensorMapSwizzle swz, CUtensorMapL2promotion l2);
CUresult encode_im2col(
        CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* gaddr,
        const uint64_t* dims, const uint64_t* strides_bytes, const int* lower, const int* upper,
        uint32_t channels_per_pixel, uint32_t pixels_per_column, const uint32_t* elem_strides,
        CUtensorMapSwizzle swz, CUtensorMapL2promotion l2, uint64_t tensor_bytes);

// ---- per-family host entry points
int igemm_query(const b200cb_desc& d, uint32_t block_n, b200cb_plan* plan);
int igemm_launch(
        const b200cb_plan& plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// weight preprocess of the W-folded stem: packed size (0 = none needed) and the packing launch
size_t igemm_packed_filter_bytes(const b200cb_plan& plan);
int igemm_pack_filter(const b200cb_plan& plan, const void* filter, void* packed, cudaStream_t stream);
int igemm_selftest(int which, cudaStream_t stream, uint64_t* mismatches);

int dwconv_query(const b200cb_desc& d, b200cb_plan* plan);
int dwconv_launch(
        const b200cb_plan& plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, cudaStream_t stream);

// NCHW4 pre-pass (relayout.cu): (B, C/4, P, 4) int8 -> (B, P, Cpad) int8, zero-padded channels
int relayout_c4_to_channels_last(
        const void* in, void* out, uint32_t B, uint32_t C, uint32_t Cpad, uint32_t P,
        uint64_t in_batch_elems, cudaStream_t stream);

// in[b][r][s] -> out[b][s][r] for 16-bit elements, output rows zero-padded to out_ld (relayout.cu)
int relayout_transpose16(
        const void* in, void* out, uint32_t B, uint32_t R, uint32_t S, uint32_t out_ld, uint64_t in_batch,
        uint64_t out_batch, cudaStream_t stream);

// in [pixels][C] -> out [pixels][Cpad], zero-padded channels; es = element bytes (relayout.cu)
int relayout_pad_channels(
        const void* in, void* out, uint64_t pixels, uint32_t C, uint32_t Cpad, uint32_t es, cudaStream_t stream);

// b200cb_relayout (relayout.cu): every mode of include/b200cb.h
int relayout_generic(
        uint32_t mode, uint32_t es, uint32_t n, uint32_t c, uint32_t h, uint32_t w, uint32_t c_out, const void* src,
        void* dst, cudaStream_t stream);

// fp32 [P][C] -> bf16 [P][3*Cp] operand split of the Float32 ConvBias path (relayout.cu)
int relayout_split3(const void* in, void* out, uint64_t P, uint32_t C, uint32_t Cp, int filter_side, cudaStream_t stream);

int device_sm_count();

inline uint32_t dtype_size(uint32_t dt) {
    switch (dt) {
        case B200CB_DT_F32:
        case B200CB_DT_QS32:
            return 4;
        case B200CB_DT_F16:
        case B200CB_DT_BF16:
            return 2;
        case B200CB_DT_QS8:
            return 1;
        default:
            return 0;
    }
}

}  // namespace b200cb

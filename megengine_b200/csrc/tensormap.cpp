// TMA descriptor encoding through driver entry points resolved at run time, so that
// libb200cb.so loads on a machine without libcuda (the C-ABI symbol test runs without a GPU).
#include <mutex>
#include "common.h"

namespace b200cb {

namespace {
using EncodeTiledFn = CUresult (*)(
        CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(
        CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
        const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::once_flag g_once;
EncodeTiledFn g_tiled = nullptr;
EncodeIm2colFn g_im2col = nullptr;
int g_driver_version = 0;

void resolve() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) ==
                cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) ==
                cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    cudaDriverGetVersion(&g_driver_version);
}
}  // namespace

CUresult encode_tiled(
        CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* gaddr,
        const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
        const uint32_t* elem_strides, CUtensorMapSwizzle swz, CUtensorMapL2promotion l2) {
    std::call_once(g_once, resolve);
    if (!g_tiled) {
        set_last_error("cuTensorMapEncodeTiled", "driver entry point not found");
        return CUDA_ERROR_NOT_FOUND;
    }
    cuuint64_t gdims[5], gstr[5];
    cuuint32_t gbox[5], gel[5];
    for (uint32_t i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        gel[i] = elem_strides ? elem_strides[i] : 1;
        if (i + 1 < rank)
            gstr[i] = strides_bytes[i];
    }
    return g_tiled(
            map, dt, rank, const_cast<void*>(gaddr), gdims, gstr, gbox, gel,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

CUresult encode_im2col(
        CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* gaddr,
        const uint64_t* dims, const uint64_t* strides_bytes, const int* lower, const int* upper,
        uint32_t channels_per_pixel, uint32_t pixels_per_column, const uint32_t* elem_strides,
        CUtensorMapSwizzle swz, CUtensorMapL2promotion l2, uint64_t tensor_bytes) {
    std::call_once(g_once, resolve);
    if (!g_im2col) {
        set_last_error("cuTensorMapEncodeIm2col", "driver entry point not found");
        return CUDA_ERROR_NOT_FOUND;
    }
    cuuint64_t gdims[5], gstr[5];
    cuuint32_t gel[5];
    for (uint32_t i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gel[i] = elem_strides[i];
        if (i + 1 < rank)
            gstr[i] = strides_bytes[i];
    }
    CUresult r = g_im2col(
            map, dt, rank, const_cast<void*>(gaddr), gdims, gstr, lower, upper, channels_per_pixel,
            pixels_per_column, gel, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // Drivers up to CUDA 13.1 set a descriptor flag that breaks im2col loads on tensors smaller
    // than 128 KiB; the public CUTLASS im2col path clears bit 21 of the second descriptor word
    // in that case and so do we.
    if (r == CUDA_SUCCESS && g_driver_version <= 13010 && tensor_bytes < 131072) {
        reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    }
    return r;
}

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace b200cb

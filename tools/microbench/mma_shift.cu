// Microbenchmark: does a ROW-SHIFTED A descriptor cost tensor-pipe time?  The rows-mode kernel (igemm_sm100.cu, a_mode 3)
// reads tap (fh, fw) of a 3x3 stride-1 layer as the resident SWIZZLE_128B input slab shifted by fh * PW + fw rows, i.e.
// with a descriptor start address that is a multiple of 128 B but not of the 1024-byte swizzle atom.  One CTA per SM
// issues L back-to-back tcgen05.mma (kind::f16, M = 128, K = 16) whose A descriptor starts `shift` rows into the slab,
// optionally cycling through the 9 tap shifts of a 3x3 window like the real issue loop; clk / MMA = elapsed / L.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I megengine_b200/csrc -I include tools/microbench/mma_shift.cu -o tools/microbench/mma_shift
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200cb::ptx;

__global__ void __launch_bounds__(128, 1) bench(uint32_t N, uint32_t shift, uint32_t pw, int taps, int L, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bar = base, tslot = base + 16;
    const uint32_t sA = base + 1024, sB = sA + 512 * 128;  // A slab 512 rows x 128 B, B 256 rows x 128 B (SW128)
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) {
        tmem_alloc(tslot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 16);
    if (threadIdx.x == 0) {
        const uint64_t bdesc = umma_smem_desc(sB, 1024, 2);
        const uint32_t idesc = umma_idesc(1, 1, 128, N);
        const long long t0 = clock64();
        if (taps == 2) {
            // the product's issue pattern: per (tap, channel block) four MMAs whose descriptors are base + 2 * k, the
            // base advanced per tap by adds only (no division, no constant-bank loads)
            const uint32_t hi = static_cast<uint32_t>(bdesc >> 32);
            const uint32_t a0 = static_cast<uint32_t>(umma_smem_desc(sA, 1024, 2)), b0 = static_cast<uint32_t>(bdesc);
            for (int i = 0; i < L; i += 4) {
                const uint32_t t = (i >> 2) & 7;
                const uint32_t a_lo = a0 + t * (pw * 8 + 8), b_lo = b0 + t * 512;
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    mma_f16(tmem + ((i >> 5) & 1) * N, (static_cast<uint64_t>(hi) << 32) | (a_lo + 2 * k),
                            (static_cast<uint64_t>(hi) << 32) | (b_lo + 2 * k), idesc, i > 64);
            }
        } else
        for (int i = 0; i < L; ++i) {
            uint32_t rows = shift;
            if (taps) {  // cycle through the 3x3 window's shifts: (i / 4) % 9 -> fh * pw + fw, 4 MMAs (K = 64) per tap
                const uint32_t t = (i >> 2) % 9;
                rows = shift + (t / 3) * pw + t % 3;
            }
            const uint64_t adesc = umma_smem_desc(sA + rows * 128 + (i & 3) * 32, 1024, 2);
            mma_f16(tmem + (i & 1) * N, adesc, bdesc, idesc, i > 1);
        }
        mma_commit(bar);
        mbar_wait(bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    const int L = 3600;
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const uint32_t Ns[] = {64, 128, 256};
    const uint32_t shifts[] = {0, 1, 2, 3, 4, 7, 8, 58, 59, 60, 116};
    auto run = [&](uint32_t N, uint32_t shift, uint32_t pw, int taps) {
        long long h[148];
        for (int rep = 0; rep < 2; ++rep) {
            bench<<<148, 128, 200 * 1024>>>(N, shift, pw, taps, L, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("N %u shift %u: %s\n", N, shift, cudaGetErrorString(e));
                return;
            }
        }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 148; ++i)
            s += h[i];
        const double clk = s / 148 / L;
        printf("N %3u  %s shift %3u rows (pw %2u): %6.1f clk / MMA -> %5.1f %% of 8192 flop/clk/SM\n", N,
               taps == 2 ? "unrolled4," : taps ? "3x3 taps," : "fixed,   ", shift, pw, clk, 100.0 * (2.0 * 128 * N * 16 / clk) / 8192.0);
    };
    for (uint32_t N : Ns) {
        for (uint32_t s : shifts)
            run(N, s, 0, 0);
        run(N, 0, 58, 1);   // ResNet-50 l1.c2: OW = 56, padded width 58
        run(N, 0, 64, 1);   // padded width a multiple of 8: fh shifts stay atom-aligned, fw = 1, 2 do not
        run(N, 0, 30, 1);   // 28 x 28 layers
        run(N, 0, 58, 2);   // the product's issue pattern (adds only, 4 MMAs per tap unrolled)
    }
    return 0;
}

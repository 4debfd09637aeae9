// Microbenchmark: how long does one tcgen05.mma (kind::f16, bf16 -> fp32, K = 16) occupy the tensor pipe
// for a given (M, N), operands from shared memory (SS) or A from tensor memory (TS)?  One CTA per SM issues
// L back-to-back MMAs on fixed (garbage) operands and waits for the commit; clk / MMA = elapsed / L.
// Guides tile-shape choices: build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I megengine_b200/csrc -I include tools/microbench/mma_shapes.cu -o /tmp/mma_shapes
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200cb::ptx;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
            "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
}

template <int NACC, int TS>
__global__ void __launch_bounds__(128, 1) bench(uint32_t M, uint32_t N, int L, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bar = base, tslot = base + 16;
    const uint32_t sA = base + 1024, sB = sA + 128 * 128;  // A 128 rows x 128 B, B 256 rows x 128 B (SW128)
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
        const uint64_t adesc = umma_smem_desc(sA, 1024, 2), bdesc = umma_smem_desc(sB, 1024, 2);
        const uint32_t idesc = umma_idesc(1, 1, M, N);
        const long long t0 = clock64();
        for (int i = 0; i < L; i += NACC) {
            // NACC independent accumulators, round robin: is the floor a dependent-accumulate latency?
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                if (TS)
                    mma_ts(tmem + a * N, tmem + 384, bdesc, idesc, i != 0);  // A from TMEM columns 384.. (garbage)
                else
                    mma_f16(tmem + a * N, adesc, bdesc, idesc, i != 0);
            }
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
    const int L = 4000;
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    
    const uint32_t Ms[] = {128, 64}, Ns[] = {256, 128, 64, 32};
    auto run = [&](auto NA, auto TSV) {
        constexpr int nacc = decltype(NA)::value, ts = decltype(TSV)::value;
        auto kern = bench<nacc, ts>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        for (uint32_t M : Ms)
            for (uint32_t N : Ns) {
                if (nacc * N > (ts ? 384u : 512u))
                    continue;
                long long h[148];
                for (int rep = 0; rep < 2; ++rep) {
                    kern<<<148, 128, 64 * 1024>>>(M, N, L, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("%s M %u N %u: %s\n", ts ? "TS" : "SS", M, N, cudaGetErrorString(e));
                        return;
                    }
                }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double s = 0;
                for (int i = 0; i < 148; ++i)
                    s += h[i];
                const double clk = s / 148 / L;
                printf("%s nacc %d M %3u N %3u : %6.1f clk / MMA  -> %5.1f %% of 8192 flop/clk/SM\n", ts ? "TS" : "SS", nacc,
                       M, N, clk, 100.0 * (2.0 * M * N * 16 / clk) / 8192.0);
            }
    };
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I4 = std::integral_constant<int, 4>;
    using I0 = std::integral_constant<int, 0>;
    run(I1{}, I0{}); run(I2{}, I0{}); run(I4{}, I0{});
    run(I1{}, I1{}); run(I2{}, I1{}); run(I4{}, I1{});
    return 0;
}

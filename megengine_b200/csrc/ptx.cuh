// Inline-PTX wrappers for the sm_100a primitives used by the ConvBias kernels:
// mbarrier, TMA (cp.async.bulk.tensor, tiled + im2col), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200cb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
    return r;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
            "{\n"
            ".reg .pred P;\n"
            "elect.sync _|P, 0xffffffff;\n"
            "selp.u32 %0, 1, 0, P;\n"
            "}\n"
            : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch
// wait: blocks until the preceding kernel of the stream has completed and its writes are visible;
// launch_dependents: lets the next kernel's CTAs be scheduled as this grid's CTAs retire.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// B200CB_WAIT_HINT_NS (experiment, tools builds): suspend-time hint of the potentially blocking try_wait --
// a waiting warp then re-issues the poll less often and leaves the issue slots to the working warps
// suspend-time hint of mbarrier.try_wait (ns): waiting warps sleep in hardware until the barrier flips.  At the
// software power cap (sustained runs, SM clock 1.6 GHz) 20 us measured 0.8 % faster than no hint, at full clocks
// the same (profiles/r02/wait_hint_sustained.txt)
#ifndef B200CB_WAIT_HINT_NS
#define B200CB_WAIT_HINT_NS 20000
#endif
// B200CB_WAIT_BACKOFF_NS (experiment, off): after 4 failed polls a waiter sleeps this long between further polls.
// ncu on the issue-bound int8 layers: a quarter of all issued instructions are the SYNCS / BRA / NANOSLEEP of waiting
// warps (a hinted try_wait is re-issued every ~150 clk whatever the hint says) -- but taking them away changed
// nothing: 300 / 1000 ns measured qs8 2.72 -> 2.76 ms, bf16 within noise (profiles/r02/wait_backoff.txt).
#ifndef B200CB_WAIT_BACKOFF_NS
#define B200CB_WAIT_BACKOFF_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if B200CB_WAIT_BACKOFF_NS > 0
    asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            ".reg .u32 cnt;\n"
            "mov.u32 cnt, 0;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
            "@P1 bra DONE;\n"
            "add.u32 cnt, cnt, 1;\n"
            "setp.lt.u32 P1, cnt, 4;\n"
            "@P1 bra LAB_WAIT;\n"
            "nanosleep.u32 %3;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(bar),
            "r"(parity), "r"(static_cast<uint32_t>(B200CB_WAIT_HINT_NS)), "r"(static_cast<uint32_t>(B200CB_WAIT_BACKOFF_NS))
            : "memory");
#elif B200CB_WAIT_HINT_NS > 0
    asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(bar),
            "r"(parity), "r"(static_cast<uint32_t>(B200CB_WAIT_HINT_NS))
            : "memory");
#else
    asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
#endif
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
            : "memory");
}
__device__ __forceinline__ void tma_load_3d(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ void tma_load_4d(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
            : "memory");
}
// plain bulk copy (no tensor map): `bytes` contiguous bytes, 16-byte aligned on both sides
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
            : "memory");
}
// im2col mode, NHWC tensor seen as (C, W, H, N): coordinates are the channel start and the
// *base pixel* (top-left tap of the window, may be negative); {off_w, off_h} select the tap.
__device__ __forceinline__ void tma_load_im2col_4d(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c, int32_t w, int32_t h, int32_t n,
        uint16_t off_w, uint16_t off_h) {
    asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
            "h"(off_w), "h"(off_h)
            : "memory");
}
// multicast variant: the box lands at the same smem offset of every CTA in `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_3d_mc(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2, uint16_t mask) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            ".multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
            : "memory");
}
// CTA-pair (cta_group::2) loads: the data lands in the executing CTA's smem, the complete_tx goes to
// `bar`, a shared::cluster address that may name the PEER CTA's mbarrier (the MMA leader's).
__device__ __forceinline__ void tma_load_2d_2sm(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
            : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(
        uint32_t dst, const void* tmap, uint32_t bar, int32_t c, int32_t w, int32_t h, int32_t n,
        uint16_t off_w, uint16_t off_h) {
    asm volatile(
            "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
            "h"(off_w), "h"(off_h)
            : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int32_t c0, int32_t c1) {
    asm volatile(
            "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(tmap)),
            "r"(src), "r"(c0), "r"(c1)
            : "memory");
}
__device__ __forceinline__ void tma_store_3d(
        const void* tmap, uint32_t src, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
            "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(tmap)),
            "r"(src), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// CTA-pair allocation: executed by one warp of EACH CTA of the pair
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind f16 covers fp16 and bf16 inputs, fp32 accumulate
__device__ __forceinline__ void mma_f16(
        uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// int8 x int8 -> int32 accumulate
__device__ __forceinline__ void mma_i8(
        uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// CTA-pair MMAs (issued by the leader CTA only): M = 256 rows, 128 from each CTA's A tile, B rows
// [0, N/2) from the leader's smem and [N/2, N) from the peer's; each CTA's TMEM gets its 128 rows
__device__ __forceinline__ void mma_f16_2sm(
        uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void mma_i8_2sm(
        uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// pair commit: arrives on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_2sm_mc(uint32_t bar, uint16_t mask) {
    asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                    bar),
            "h"(mask)
            : "memory");
}
// accumulate-always forms: the enable-input-d predicate is a compile-time constant (folds to PT), so the
// issuing thread spends no instruction on it
__device__ __forceinline__ void mma_f16_acc(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc)
            : "memory");
}
__device__ __forceinline__ void mma_i8_acc(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc)
            : "memory");
}
__device__ __forceinline__ void mma_f16_2sm_acc(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc)
            : "memory");
}
__device__ __forceinline__ void mma_i8_2sm_acc(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc)
            : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile(
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
            : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (thread-block cluster)
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile(
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                    bar),
            "h"(mask)
            : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp receives columns [c, c+32) of TMEM lane
// (lane_base + i). A warp may only touch the lane quadrant 32*(warp_id % 4).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
              "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr)
            : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// K-major operand tile in shared memory: rows of `swizzle_bytes` bytes, 8-row groups
// `8 * swizzle_bytes` apart (SBO), TMA swizzle mode == UMMA layout type.
//   bits [0,14)  start address >> 4      bits [16,30) LBO >> 4 (ignored for swizzled K-major)
//   bits [32,46) SBO >> 4                bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout: 0 none, 2 SW128, 4 SW64, 6 SW32
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout & 7) << 61;
    return d;
}
// K-major operand without swizzle ("interleaved" core matrices): a core matrix is 8 rows x 16 bytes
// stored as 128 contiguous bytes; LBO = byte distance between the two core matrices of one K=32-byte
// step, SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_smem_desc_lbo(
        uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout & 7) << 61;
    return d;
}
// instruction descriptor (dense, K-major A and B):
//   [4,6) D format (1 = F32, 2 = S32)   [7,10) A format   [10,13) B format
//   (kind::f16: 0 = F16, 1 = BF16; kind::i8: 0 = U8, 1 = S8)
//   [17,23) N >> 3    [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t dfmt, uint32_t abfmt, uint32_t m, uint32_t n) {
    return (dfmt << 4) | (abfmt << 7) | (abfmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// byte offset -> swizzled byte offset for Swizzle<B,4,3> with B = log2(swizzle_bytes / 16):
// XOR address bits [4, 4+B) with bits [7, 7+B).
template <int SWZ_BYTES>
__device__ __forceinline__ uint32_t swizzle_off(uint32_t off) {
    constexpr uint32_t mask = (SWZ_BYTES / 16 - 1);
    return off ^ (((off >> 7) & mask) << 4);
}

}  // namespace ptx
}  // namespace b200cb

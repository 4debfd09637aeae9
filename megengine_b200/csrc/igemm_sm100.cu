// Dense NHWC ConvBias as an implicit GEMM on Blackwell tensor cores (sm_100a).
//
//   GEMM view: M = N*OH*OW output pixels, N = OC, K = FH*FW*IC.
//   A (activations) tiles: TMA im2col mode straight from the NHWC tensor -- 128 pixels x KB bytes
//     of channels per (filter tap, channel block); padding / stride / image wrap-around are done
//     by the TMA unit (zero fill), so no im2col buffer ever exists in HBM.  Pointwise stride-1
//     convs use a plain tiled map over the (IC, M) matrix.
//   B (filter, KRSC) tiles: tiled TMA, BLOCK_N output channels x KB bytes.
//   K-block width KB is 128, 64 or 32 bytes (swizzle mode 128B / 64B / 32B) so that layers with few
//     input channels do not multiply zero padding.  Stride-2 convs with <= 16 bytes of channels
//     (the network stem, IC padded 3 -> 8) are "W-folded": pixel pairs are viewed as 2*IC channels,
//     which turns the FWx stride-2 window into a ceil((FW+1)/2)x stride-1 window over a tensor of
//     half the width; the filter is re-packed accordingly (workspace or weight-preprocess).
//   MMA: tcgen05.mma cta_group::1, M=128, N=BLOCK_N, K=32 bytes per instruction, fp32 (bf16/fp16)
//     or int32 (int8) accumulators in TMEM, double buffered so the epilogue of tile i overlaps
//     the main loop of tile i+1.
//   Epilogue: 4 (16-bit) or 8 (int8) warps, each fully independent: tcgen05.ld of its 32 TMEM lanes
//     -> registers -> (+bias, +z, activation, requantize) -> its own swizzled smem slab -> its own
//     TMA store.  z tiles are prefetched by TMA into smem by a dedicated warp.
//
// Replaces (does not call) the reference algos listed in SURVEY.md 8(a13):
//   dnn/src/cuda/conv_bias/implicit_gemm_int8_nhwc_imma.cpp, cudnn_conv_bias_activation.cpp,
//   bfloat16.cpp.  Epilogue arithmetic follows the naive oracle
//   dnn/src/naive/conv_bias/opr_impl.cpp:21-98.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include "common.h"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace b200cb {

struct IgemmParams {
    CUtensorMap tmA;  // activations: im2col (C,W,H,N) or tiled (C, M); stem: even pixels (C, W/2, H, N)
    CUtensorMap tmA2; // stem only: odd pixels
    CUtensorMap tmB;  // filter (IC, FH*FW, OC)
    CUtensorMap tmD;  // dst (OC, M), box = EPI_N x 32 rows (one epilogue warp)
    CUtensorMap tmZ;  // z   (OC, M), box = EPI_N x 128 rows
    const void* bias;
    void* dst;        // rows mode stores without TMA
    uint32_t M, OC;
    uint32_t OHW, OW;
    uint32_t FW, num_taps;
    int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    uint32_t num_kc;   // channel blocks per tap
    uint32_t kb_bytes; // bytes of K per pipeline stage row: 128 / 64 / 32
    uint32_t kb_elems;
    uint32_t m_tiles, n_tiles;
    uint32_t stages;
    uint32_t a_mode;   // 0 tiled, 1 im2col, 2 stem (resident filter, one input-row slab per tile)
    // stem mode: tile = 128 consecutive output pixels of ONE output row
    uint32_t OH, nseg;          // output rows per image, 128-pixel segments per output row
    uint32_t stem_nj;           // pixel pairs per staged input row
    uint32_t stem_fh, stem_fw2; // filter rows, tap pairs per row
    int32_t stem_pad_lo2;       // pixel pairs of left padding
    // stem4: a pixel PAIR is one 16-byte unit (IC * es == 8, e.g. bf16 IC = 4): a single plane, K step =
    // two consecutive pairs (LBO = 16 B), filter rows of 64 bytes (SWIZZLE_64B)
    uint32_t stem4;
    // stem "row ring" (stem4, one segment per output row): input rows arrive as whole contiguous boxes, two
    // rows (one even / odd row pair) per box, into a ring of STEM_RING_PAIRS pair slots; consecutive output
    // rows of an image share all but one pair, and every CTA walks a contiguous range of output rows
    uint32_t stem_ring;      // 0 = off
    uint32_t stem_rowb;      // bytes of one staged input row (stem_nj pixel pairs x 16)
    uint32_t stem_np;        // row pairs one output row needs: ceil((FH + (pad_h & 1)) / 2)
    // rows mode (a_mode 3): stride-1 conv over the padded-linear pixel index m' = oh * PW + ow',
    // PW = OW + FW - 1; a tile = 128 consecutive m' of one image, its input rows resident in smem
    uint32_t PW, tiles_per_img, rows_rh;
    uint32_t cluster;  // CTAs per thread-block cluster sharing each filter tile by TMA multicast (1, 2, 4)
    uint32_t pair;     // 1: the 2 CTAs of a cluster form a cta_group::2 MMA pair (256 x BN tile, B split)
    uint32_t flip;     // mode == CONVOLUTION
    uint32_t has_bias, has_z, nonline;  // has_z: z arrives through the TMA ring (NHWC)
    // NCHW4 destination (N, OC/4, OH, OW, 4): dst and z are accessed in place by the epilogue
    uint32_t nz;  // z ring depth
    // 1 (a_mode 0, single CTA): the CTA's filter tile (all k-blocks of ONE N tile) is loaded once and stays in shared
    // memory; the grid is a multiple of n_tiles, so every tile of a CTA has the same N index and the pipeline
    // stages carry the activations only (half the bytes in flight per tile on the residual 1x1 layers)
    uint32_t b_res;
    uint32_t z_jit;  // 1: the z chunks of a tile are requested only once the previous tile's accumulator is complete
    uint32_t nchw4, z_direct;
    // 1: the epilogue copies its staged rows out with coalesced st.global instead of a TMA store.  A TMA
    // store queues behind every load the producer has in flight, so waiting for it to release a slab or
    // a z chunk costs a whole queue drain (thousands of clocks when the layer is HBM bound); plain
    // stores free the staging memory as soon as the warp has read it back.
    uint32_t direct_store;
    // OUT32 kernels: 0 = QuantizedS32 dst (int32 accumulator + bias), 1 = Float32 dst: float(acc) * acc_scale
    // (QuantizedS8 inputs) or the fp32 accumulator (bf16 inputs) + fp32 bias [+ fp32 z], activation
    uint32_t out_f32;
    float acc_scale;
    uint32_t acc_bound;  // QuantizedS8: K * 2^14 >= |accumulator| (saturated), for the conversion-free int -> float path
    const void* z;
    RequantParams rq;
    long long* dbg;   // -DB200CB_DEBUG_TIMING builds only: 148 x 3 roles x 17 counters
};

namespace {

using namespace ptx;

// Optional in-kernel wait-time accounting (build with -DB200CB_DEBUG_TIMING, see tools/): each warp
// role accumulates the cycles it spends blocked on each mbarrier class; CTA 0..gridDim write 8 counters.
#ifdef B200CB_DEBUG_TIMING
#define TIMED_WAIT(slot, bar, parity)                  \
    do {                                               \
        const long long t0__ = clock64();              \
        mbar_wait(bar, parity);                        \
        dbg_acc[slot] += clock64() - t0__;             \
    } while (0)
#define DBG_MARK(v) const long long v = clock64()
#define DBG_ADD(slot, since) dbg_acc[slot] += clock64() - (since)
#else
#define TIMED_WAIT(slot, bar, parity) mbar_wait(bar, parity)
#define DBG_MARK(v) do {} while (0)
#define DBG_ADD(slot, since) do {} while (0)
#endif

// UMMA shared-memory descriptor split into its 32-bit halves so that the single MMA-issuing thread
// advances operands with one integer add instead of rebuilding 64-bit descriptors:
//   lo = (addr >> 4) | (LBO >> 4) << 16        hi = (SBO >> 4) | version 1 << 14 | layout << 29
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
template <int KIND>
__device__ __forceinline__ void mma_issue_2sm(
        uint32_t taddr, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
        uint32_t accumulate) {
    const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | a_lo;
    const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | b_lo;
    if constexpr (KIND == KIND_S8)
        mma_i8_2sm(taddr, adesc, bdesc, idesc, accumulate);
    else
        mma_f16_2sm(taddr, adesc, bdesc, idesc, accumulate);
}
template <int KIND>
__device__ __forceinline__ void mma_issue(
        uint32_t taddr, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
        uint32_t accumulate) {
    const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | a_lo;
    const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | b_lo;
    if constexpr (KIND == KIND_S8)
        mma_i8(taddr, adesc, bdesc, idesc, accumulate);
    else
        mma_f16(taddr, adesc, bdesc, idesc, accumulate);
}

template <int KIND, int OUT32>
struct Traits {
    static constexpr int IN_ES = (KIND == KIND_S8) ? 1 : 2;
    static constexpr int OUT_ES = OUT32 ? 4 : IN_ES;
    static constexpr int EPI_N = OUT32 ? 32 : 64;       // columns per epilogue chunk
    static constexpr int ROW_BYTES = EPI_N * OUT_ES;    // 128 (16-bit, 32-bit) or 64 (int8)
    // Epilogue warp groups (4 warps each, one per TMEM lane quadrant).  Layers with a short k-loop are
    // bound by the epilogue, and the epilogue by latency, not by issue slots (profiles/r02: with 8 warps
    // the MMA warp waits 70-85 % of the time for an accumulator while the epilogue warps run at ~0.6
    // instructions / clk / scheduler): 16 epilogue warps, every group on its own chunk.
    static constexpr int GROUPS = OUT32 ? 1 : 4;
    static constexpr int THREADS = 128 + 128 * GROUPS;
    static constexpr int SLAB_BYTES = 32 * ROW_BYTES;   // one warp's staging slab (32 rows)
    // slabs per warp: double buffered where 64 KiB allow it (int8, int32)
    __host__ __device__ static constexpr int nslab(int) { return (4 * GROUPS * 2 * SLAB_BYTES <= 65536) ? 2 : 1; }
    __host__ __device__ static constexpr int slabs_bytes(int bn) { return 4 * GROUPS * nslab(bn) * SLAB_BYTES; }
    static constexpr int ZBUF_BYTES = 128 * ROW_BYTES;  // one z chunk
    // accumulator buffers in TMEM (512 columns): 4 for tiles up to 128 columns, 2 for 256
    __host__ __device__ static constexpr int nacc(int bn) { return (OUT32 || bn > 128) ? 2 : 4; }
};

// shared memory map (offsets from the 1024-aligned base)
constexpr uint32_t OFF_FULL = 0;       // [16] x 8
constexpr uint32_t OFF_EMPTY = 128;    // [16] x 8
constexpr uint32_t OFF_TFULL = 256;    // [4]
constexpr uint32_t OFF_TEMPTY = 288;   // [4]
constexpr uint32_t OFF_ZFULL = 320;    // [8]
constexpr uint32_t OFF_ZEMPTY = 384;   // [8]
constexpr uint32_t OFF_TMEM = 448;
constexpr uint32_t OFF_BRES = 456;     // stem: resident filter barrier
// z ring depth: 6 (4 epilogue groups busy + 2 chunks in flight) up to 8 -- short-k layers are bound by
// the z / dst streams, so their spare shared memory holds z chunks (a tile or more ahead) rather than
// more A/B stages
constexpr uint32_t NZ_MIN = 4, NZ_MAX = 8;  // GROUPS sub-rings of depth 1 or 2 (see the z prefetch warp)
constexpr uint32_t STEM_NJ_K = 136;    // pixel pairs per staged stem input row (== STEM_NJ on the host)
constexpr uint32_t STEM_RING_PAIRS = 16, STEM_RING_AHEAD = 4;  // row-pair slots; tiles the producer runs ahead
constexpr uint32_t OFF_BIAS = 1024;    // 1 KiB: per epilogue group the 64 bias values of its chunk (OUT32: the tile's BN)
// host mirror of Traits::slabs_bytes
__host__ __device__ constexpr uint32_t slabs_bytes(uint32_t groups, uint32_t /*bn*/, uint32_t row_bytes) {
    return 4 * groups * ((4 * groups * 2 * 32 * row_bytes <= 65536) ? 2u : 1u) * 32 * row_bytes;
}
__host__ __device__ constexpr uint32_t off_sd(uint32_t /*groups*/, uint32_t /*bn*/) {
    return OFF_BIAS + 1024u;
}
// With z: the epilogue area is a ring of NZ z chunks (128 rows x ROW_BYTES, TMA-prefetched) and the
// epilogue works IN PLACE -- each warp overwrites its 32 rows of the z chunk with the outputs and
// TMA-stores them from there; without z it is the per-warp slabs.  Then the pipeline stages.

// QuantizedS8 fast epilogue for one thread's 32 accumulator columns (IDENTITY / RELU, fast_div == 2):
// bias add, dequantize, optional z, activation, correctly rounded division, round-half-away, clamp,
// pack -- all on the FMA / ALU pipes (see epilogue.cuh), written as two swizzled 16-byte pieces.
// MAGIC: the bias slice in shared memory carries I2F_MAGIC_BITS and |acc + bias| < 2^22 was proven for the
// chunk (epilogue.cuh): no conversion-pipe instruction is left in the sequence.
template <bool HAS_Z, bool NONNEG, int MAGIC>   // MAGIC: the CVT selector of requant_pair_* (epilogue.cuh)
__device__ __forceinline__ void requant_row32(
        const uint32_t (&r)[32], const uint32_t* bcol, const uint8_t* zrow, uint8_t* drow, uint32_t h,
        uint32_t lane, const RequantParams& rq) {
    const RequantPacked kp = make_requant_packed(rq);
#pragma unroll
    for (uint32_t j = 0; j < 2; ++j) {
        // 64-byte rows, SWIZZLE_64B: 16-byte piece index ^= (row >> 1) & 3
        const uint32_t off = (h * 2 + j) ^ ((lane >> 1) & 3);
        uint32_t zx[4] = {0, 0, 0, 0};
        if constexpr (HAS_Z) {
            const uint4 zv = *reinterpret_cast<const uint4*>(zrow + off * 16);
            zx[0] = zv.x ^ 0x80808080u, zx[1] = zv.y ^ 0x80808080u;
            zx[2] = zv.z ^ 0x80808080u, zx[3] = zv.w ^ 0x80808080u;
        }
        uint32_t ow[4];
#pragma unroll
        for (uint32_t w = 0; w < 4; ++w) {
            const uint4 bv = *reinterpret_cast<const uint4*>(bcol + j * 16 + w * 4);
            const uint32_t bb[4] = {bv.x, bv.y, bv.z, bv.w};
            uint32_t ub[4];
            if constexpr (NONNEG) {
                // RELU: two elements per instruction (FMUL2 / FFMA2 / FADD2)
#pragma unroll
                for (uint32_t b = 0; b < 4; b += 2) {
                    const int32_t s0 = static_cast<int32_t>(r[j * 16 + w * 4 + b] + bb[b]);
                    const int32_t s1 = static_cast<int32_t>(r[j * 16 + w * 4 + b + 1] + bb[b + 1]);
                    requant_pair_nonneg<HAS_Z, MAGIC>(
                            s0, s1, __byte_perm(zx[w], 0x4B000000u, 0x7540u | b),
                            __byte_perm(zx[w], 0x4B000000u, 0x7540u | (b + 1)), kp, ub[b], ub[b + 1]);
                }
            } else {
#pragma unroll
                for (uint32_t b = 0; b < 4; b += 2) {
                    const int32_t s0 = static_cast<int32_t>(r[j * 16 + w * 4 + b] + bb[b]);
                    const int32_t s1 = static_cast<int32_t>(r[j * 16 + w * 4 + b + 1] + bb[b + 1]);
                    requant_pair_signed<HAS_Z, MAGIC>(
                            s0, s1, __byte_perm(zx[w], 0x4B000000u, 0x7540u | b),
                            __byte_perm(zx[w], 0x4B000000u, 0x7540u | (b + 1)), kp, ub[b], ub[b + 1]);
                }
            }
            ow[w] = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
        }
        *reinterpret_cast<uint4*>(drow + off * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// NCHW4 epilogue for one thread's 32 accumulator columns: the thread owns one output pixel, its 32
// channels are 8 int8x4 words that live in 8 consecutive (N, OC/4) planes `plane_words` apart, so a
// warp's 32 pixels make every store (and z load) a coalesced 128-byte line.  `nwords` = planes that
// exist (OC tail / rows beyond M: 0).  Same arithmetic as requant_row32; any activation.
// MODE: 0 = generic sequence (any activation / scale), 1 = fast IDENTITY, 2 = fast RELU
template <bool HAS_Z, int MODE>
__device__ __forceinline__ void requant_row32_nchw4(
        const uint32_t (&r)[32], const uint32_t* bcol, const uint32_t* zw, uint32_t* dw,
        size_t plane_words, uint32_t nwords, uint32_t nl, const RequantParams& rq) {
    const RequantPacked kp = make_requant_packed(rq);
    uint32_t zv[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
        zv[j] = (HAS_Z && j < nwords) ? __ldg(zw + j * plane_words) : 0u;
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
        const uint4 bv = *reinterpret_cast<const uint4*>(bcol + j * 4);
        const uint32_t bb[4] = {bv.x, bv.y, bv.z, bv.w};
        const uint32_t zx = zv[j] ^ 0x80808080u;
        uint32_t ub[4];
        if constexpr (MODE != 0) {
#pragma unroll
            for (uint32_t b = 0; b < 4; b += 2) {
                const int32_t s0 = static_cast<int32_t>(r[j * 4 + b] + bb[b]);
                const int32_t s1 = static_cast<int32_t>(r[j * 4 + b + 1] + bb[b + 1]);
                const uint32_t z0 = __byte_perm(zx, 0x4B000000u, 0x7540u | b);
                const uint32_t z1 = __byte_perm(zx, 0x4B000000u, 0x7540u | (b + 1));
                if constexpr (MODE == 2)
                    requant_pair_nonneg<HAS_Z>(s0, s1, z0, z1, kp, ub[b], ub[b + 1]);
                else
                    requant_pair_signed<HAS_Z>(s0, s1, z0, z1, kp, ub[b], ub[b + 1]);
            }
        } else
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b) {
            const int32_t sfb = static_cast<int32_t>(r[j * 4 + b] + bb[b]);
            if constexpr (MODE != 0) {
                float x = __fmul_rn(__int2float_rn(sfb), rq.s_bias);
                if constexpr (HAS_Z) {
                    const float zf =
                            __uint_as_float(__byte_perm(zx, 0x4B000000u, 0x7540u | b)) - 8388736.f;
                    x = __fadd_rn(x, __fmul_rn(zf, rq.s_z));
                }
                if constexpr (MODE == 2) {
                    x = fmaxf(x, 0.f);
                    ub[b] = round_sat_s8_nonneg(div_by_scale_noguard(x, rq));
                } else {
                    ub[b] = round_sat_s8_signed(div_by_scale_noguard(x, rq));
                }
            } else {
                const int32_t zq = static_cast<int32_t>(static_cast<int8_t>((zv[j] >> (8 * b)) & 0xff));
                ub[b] = static_cast<uint32_t>(requant_s8(sfb, HAS_Z ? zq : 0, nl, rq));
            }
        }
        if (j < nwords)
            dw[j * plane_words] = pack_low_bytes(ub[0], ub[1], ub[2], ub[3]);
    }
}

// PAIR = 1 instantiations contain cta_group::2 instructions and must be launched as clusters of 2
// (the driver rejects them otherwise), hence a template parameter rather than a runtime flag.
template <int KIND, int BN, int OUT32, int PAIR>
__global__ void __launch_bounds__(Traits<KIND, OUT32>::THREADS, 1)
        igemm_kernel(const __grid_constant__ IgemmParams p) {
    using T = Traits<KIND, OUT32>;
    constexpr uint32_t NACC = T::nacc(BN);  // accumulator buffers: tile sequence number % NACC
    constexpr uint32_t TMEM_COLS = NACC * BN;  // 128 / 256 / 512: powers of two >= 32
    constexpr uint32_t NUM_CHUNKS = BN / T::EPI_N;
    // chunk c of the CTA's ts-th tile belongs to epilogue group (ts * NUM_CHUNKS + c) % GROUPS: with 4
    // groups a 256-column tile is drained by all of them at once, 128-column tiles alternate between
    // two pairs of groups and 64-column tiles rotate through the four
    constexpr uint32_t GROUPS_PER_TILE = NUM_CHUNKS < T::GROUPS ? NUM_CHUNKS : T::GROUPS;
    constexpr uint32_t IDESC =
            (KIND == KIND_S8) ? umma_idesc(2, 1, 128, BN)
                              : umma_idesc(1, KIND == KIND_BF16 ? 1 : 0, 128, BN);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    uint8_t* smem_gen = smem_raw;
    if (smem_base & 1023u)
        __trap();  // swizzled TMA / UMMA tiles need a 1024-byte aligned window
    constexpr uint32_t OFF_SD = off_sd(T::GROUPS, BN);
    constexpr uint32_t OFF_SZ = OFF_SD;  // the z ring replaces the slabs

    const uint32_t bar_full = smem_base + OFF_FULL;
    const uint32_t bar_empty = smem_base + OFF_EMPTY;
    const uint32_t bar_tfull = smem_base + OFF_TFULL;
    const uint32_t bar_tempty = smem_base + OFF_TEMPTY;
    const uint32_t bar_zfull = smem_base + OFF_ZFULL;
    const uint32_t bar_zempty = smem_base + OFF_ZEMPTY;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_TMEM);

    const bool stem = p.a_mode == 2;
    const bool rows = p.a_mode == 3;
    const uint32_t plane_bytes = rows ? p.rows_rh * p.PW * p.kb_bytes   // rows: one channel block (128 / 64 B) of the slab
                                      : p.stem_fh * p.stem_nj * 16; // stem: one pixel parity, all filter rows
    const uint32_t a_bytes = stem ? (p.stem4 ? plane_bytes : 2 * plane_bytes)
                             : rows ? p.num_kc * plane_bytes
                                    : IGEMM_BLOCK_M * p.kb_bytes;
    const uint32_t stem_rb = p.stem4 ? 64u : 128u;  // bytes of one packed filter row
    // CTA pair: each CTA stages only its half of the filter tile's rows
    constexpr bool pair = PAIR != 0;
    const uint32_t b_rows = pair ? BN / 2 : BN;
    const bool bres_tiled = p.b_res != 0;  // generic 1x1 path with the N tile's filter resident
    const uint32_t stage_bytes = (stem || rows || bres_tiled) ? a_bytes : a_bytes + b_rows * p.kb_bytes;
    const uint32_t bres_bytes = stem ? p.stem_fh * BN * stem_rb
                                     : (rows || bres_tiled) ? p.num_taps * p.num_kc * BN * p.kb_bytes : 0u;
    const uint32_t sZ = smem_base + OFF_SZ;
    constexpr uint32_t NSLAB = T::nslab(BN);
    const uint32_t nz = p.nz;
    const uint32_t sBres = sZ + (p.has_z ? nz * T::ZBUF_BYTES : T::slabs_bytes(BN));  // resident filter
    const uint32_t sStage = sBres + bres_bytes;
    const uint32_t bar_bres = smem_base + OFF_BRES;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
#ifdef B200CB_DEBUG_TIMING
    long long dbg_acc[16] = {0};
    const long long dbg_t0 = clock64();
#endif
    const uint32_t num_kb = p.num_taps * p.num_kc;
    // Tile schedule: a cluster of C CTAs walks "cluster tiles" = C consecutive M tiles x one N tile, so
    // that all its CTAs need the same filter tile at the same time (each loads 1/C of it and
    // multicasts).  C == 1 (always in stem / rows mode) degenerates to tile = blockIdx.x + i * gridDim.x.
    const uint32_t C = p.cluster;
    const uint32_t crank = C > 1 ? cluster_ctarank() : 0;
    const uint32_t cl_id = blockIdx.x / C, cl_num = gridDim.x / C;
    const uint32_t total_ct = ((p.m_tiles + C - 1) / C) * p.n_tiles;
    const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1);
    // tiles of this CTA (cluster): ct = t_first, t_first + t_step, ... < t_end.  Round robin over the grid,
    // except for the stem's row ring, where every CTA owns a contiguous range of output rows
    const bool ring = p.stem_ring != 0;
    const uint32_t t_first = ring ? static_cast<uint32_t>(static_cast<uint64_t>(blockIdx.x) * total_ct / gridDim.x) : cl_id;
    const uint32_t t_end = ring ? static_cast<uint32_t>(static_cast<uint64_t>(blockIdx.x + 1) * total_ct / gridDim.x) : total_ct;
    const uint32_t t_step = ring ? 1u : cl_num;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&p.tmA);
        prefetch_tensormap(&p.tmB);
        if (!p.nchw4)
            prefetch_tensormap(&p.tmD);
        if (p.has_z)
            prefetch_tensormap(&p.tmZ);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t i = 0; i < p.stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            // multicast clusters: one MMA commit per CTA; pair: one (multicast) commit of the leader
            mbar_init(bar_empty + 8 * i, pair ? 1 : C);
        }
        for (uint32_t i = 0; i < NACC; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            // one arrive per epilogue warp that reads the accumulator (pair: of both CTAs, on the leader's
            // barrier)
            mbar_init(bar_tempty + 8 * i, (pair ? 2 : 1) * 4 * GROUPS_PER_TILE);
        }
        for (uint32_t i = 0; i < nz; ++i) {
            mbar_init(bar_zfull + 8 * i, 1);
            mbar_init(bar_zempty + 8 * i, 4);               // the 4 warps of the consuming group
        }
        mbar_init(bar_bres, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if (pair) {
            tmem_alloc_2sm(smem_base + OFF_TMEM, TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(smem_base + OFF_TMEM, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (C > 1)
        cluster_sync_all();  // peers multicast into this CTA's stages and arrive on its barriers
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
    // prefetch) overlapped the tail of the previous kernel in the stream; no global memory is read or
    // written before its completion is observed here.  The next kernel may start its own prologue as
    // soon as this grid's CTAs retire.
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0 && ring) {
        // ============================ TMA producer, stem row ring ============================
        // tile = output row (img, oh) of the CTA's contiguous range.  Its window is the last stem_np row
        // pairs loaded: a tile that continues the previous one's image loads ONE new pair (2 input rows,
        // one box), the first tile of the range / of an image loads all stem_np.  Pair seq q lives in ring
        // slot q % STEM_RING_PAIRS; tile t is loaded once tile t - STEM_RING_AHEAD has been consumed, which
        // keeps every slot it overwrites (seq <= q - 16) clear of the windows still in use (a tile adds at
        // most stem_np = 4 pairs: 16 >= 4 * STEM_RING_AHEAD).
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_bres, bres_bytes);
            for (uint32_t fh = 0; fh < p.stem_fh; ++fh)
                tma_load_3d(sBres + fh * BN * stem_rb, &p.tmB, bar_bres, 0, static_cast<int32_t>(fh), 0);
        }
        uint32_t stage = 0, phase = 0, pc = 0;
        for (uint32_t ct = t_first; ct < t_end; ++ct) {
            const uint32_t img = ct / p.OH, oh = ct % p.OH;
            const uint32_t nnew = (ct == t_first || oh == 0) ? p.stem_np : 1u;
            // even row of the window's first pair: 2 * oh - pad_h rounded down to even
            const int32_t row0 = static_cast<int32_t>(oh) * 2 - p.pad_h - (p.pad_h & 1);
            TIMED_WAIT(0, bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
                const uint32_t full = bar_full + 8 * stage;
                mbar_arrive_expect_tx(full, nnew * 2 * p.stem_rowb);
                for (uint32_t k = 0; k < nnew; ++k) {
                    const uint32_t slot = (pc + k) % STEM_RING_PAIRS;
                    const int32_t ih = row0 + 2 * static_cast<int32_t>(p.stem_np - nnew + k);
                    tma_load_3d(sStage + slot * 2 * p.stem_rowb, &p.tmA, full, -2 * p.stem_pad_lo2, ih,
                                static_cast<int32_t>(img));
                }
            }
            __syncwarp();
            pc += nnew;
            if (++stage == STEM_RING_AHEAD) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1 && ring) {
        // ============================ MMA issuer, stem row ring ============================
        // as the classic stem issuer, but filter row fh reads ring row 2 * slot + (r & 1) with
        // r = fh + (pad_h & 1) the row's offset in the window and slot = (window's first pair + r / 2) % 16
        if (elect_one()) {
            auto mma_acc = [&](uint32_t taddr, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi) {
                const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | a_lo;
                const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | b_lo;
                if constexpr (KIND == KIND_S8)
                    mma_i8_acc(taddr, adesc, bdesc, IDESC);
                else
                    mma_f16_acc(taddr, adesc, bdesc, IDESC);
            };
            const uint32_t ksteps = (p.stem_fw2 + 1) / 2;   // stem4: one K step (32 bytes) per two tap pairs
            const uint32_t a_hi = desc_hi(128, 0), b_hi = desc_hi(512, 4);
            const uint32_t a0 = desc_lo(sStage, 16u), b0 = desc_lo(sBres, 16);
            const uint32_t row16 = p.stem_rowb >> 4, brow = BN * stem_rb / 16u;
            const uint32_t rodd = static_cast<uint32_t>(p.pad_h) & 1u;
            constexpr uint32_t RING_ROWS = 2 * STEM_RING_PAIRS;  // ring row of window row r: (2 * w0 + r) % 32
            auto run = [&](auto KS_) {
                constexpr uint32_t KS = decltype(KS_)::value;
                uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, pc = 0;
                TIMED_WAIT(6, bar_bres, 0);
                for (uint32_t ct = t_first; ct < t_end; ++ct) {
                    const uint32_t oh = ct % p.OH;
                    pc += (ct == t_first || oh == 0) ? p.stem_np : 1u;
                    uint32_t rr = (2 * (pc - p.stem_np) + rodd) % RING_ROWS;  // ring row of filter row 0
                    TIMED_WAIT(1, bar_tempty + 8 * acc, acc_phase ^ 1);
                    TIMED_WAIT(2, bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + acc * BN;
                    uint32_t a_lo = a0 + rr * row16, b_lo = b0;
                    mma_issue<KIND>(taddr, a_lo, a_hi, b_lo, b_hi, IDESC, 0);  // overwrites the accumulator
#pragma unroll
                    for (uint32_t t2 = 1; t2 < KS; ++t2)
                        mma_acc(taddr, a_lo + t2 * 2, a_hi, b_lo + 2 * t2, b_hi);
                    for (uint32_t fh = 1; fh < p.stem_fh; ++fh) {
                        rr = (rr + 1) % RING_ROWS;
                        a_lo = a0 + rr * row16;
                        b_lo += brow;
#pragma unroll
                        for (uint32_t t2 = 0; t2 < KS; ++t2)
                            mma_acc(taddr, a_lo + t2 * 2, a_hi, b_lo + 2 * t2, b_hi);
                    }
                    mma_commit(bar_empty + 8 * stage);
                    mma_commit(bar_tfull + 8 * acc);
                    if (++stage == STEM_RING_AHEAD) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (++acc == NACC) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
            };
            if (ksteps == 1)
                run(std::integral_constant<uint32_t, 1>{});
            else
                run(std::integral_constant<uint32_t, 2>{});
        }
        __syncwarp();
    } else if (warp == 0 && stem) {
        // ============================ TMA producer, stem mode ============================
        // the whole (packed) filter once, then per tile the FH input rows it needs, de-interleaved
        // into even / odd pixel planes: [parity][fh][pixel pair][16 bytes], no swizzle
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_bres, bres_bytes);
            for (uint32_t fh = 0; fh < p.stem_fh; ++fh)
                tma_load_3d(sBres + fh * BN * stem_rb, &p.tmB, bar_bres, 0, static_cast<int32_t>(fh), 0);
        }
        uint32_t stage = 0, phase = 0;
        for (uint32_t ct = t_first; ct < t_end; ct += t_step) {
            const uint32_t tile = ct;  // C == 1 in this mode
            const uint32_t seg = tile % p.nseg, row3 = tile / p.nseg;  // row3 = img * OH + oh
            const uint32_t img = row3 / p.OH, oh = row3 % p.OH;
            const int32_t j0 = static_cast<int32_t>(seg * IGEMM_BLOCK_M) - p.stem_pad_lo2;
            const int32_t ih0 = static_cast<int32_t>(oh) * p.stride_h - p.pad_h;
            TIMED_WAIT(0, bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
                const uint32_t full = bar_full + 8 * stage;
                const uint32_t sa = sStage + stage * stage_bytes;
                mbar_arrive_expect_tx(full, stage_bytes);
                tma_load_4d(sa, &p.tmA, full, 0, j0, ih0, static_cast<int32_t>(img));
                if (!p.stem4)
                    tma_load_4d(sa + plane_bytes, &p.tmA2, full, 0, j0, ih0, static_cast<int32_t>(img));
            }
            __syncwarp();
            if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1 && stem) {
        // ============================ MMA issuer, stem mode ============================
        // One thread, every MMA of a tile unrolled per filter row with immediate descriptor offsets and a
        // constant accumulate predicate: a 128 x 64 MMA occupies the tensor pipe only 48 clk, so the
        // issue loop itself bounds this mode (tools/microbench/mma_shapes.cu).
        // A: rows = output pixels (16 bytes apart), K = [even pixel 8ch | odd pixel 8ch]: no swizzle,
        // LBO = parity-plane distance, SBO = 128 B; one entry (16 B) per tap pair.
        // stem4: K = [pair j | pair j + 1], the two halves 16 bytes apart, one K step per two pairs.
        auto mma_acc = [&](uint32_t taddr, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi) {
            const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | a_lo;
            const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | b_lo;
            if constexpr (KIND == KIND_S8)
                mma_i8_acc(taddr, adesc, bdesc, IDESC);
            else
                mma_f16_acc(taddr, adesc, bdesc, IDESC);
        };
        auto run = [&](auto KS_, auto AS_) {
            constexpr uint32_t KS = decltype(KS_)::value;   // K steps (MMAs) per filter row
            constexpr uint32_t AS = decltype(AS_)::value;   // A entries (16 B) per K step
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const uint32_t a_hi = desc_hi(128, 0);
            const uint32_t b_hi = AS == 2 ? desc_hi(512, 4) : desc_hi(1024, 2);
            const uint32_t a0 = desc_lo(sStage, AS == 2 ? 16u : plane_bytes), b0 = desc_lo(sBres, 16);
            const uint32_t stage16 = stage_bytes >> 4, brow = BN * stem_rb / 16u;
            uint32_t soff = 0;
            TIMED_WAIT(6, bar_bres, 0);
            for (uint32_t ct = t_first; ct < t_end; ct += t_step) {
                TIMED_WAIT(1, bar_tempty + 8 * acc, acc_phase ^ 1);
                TIMED_WAIT(2, bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BN;
                uint32_t a_lo = a0 + soff, b_lo = b0;
                mma_issue<KIND>(taddr, a_lo, a_hi, b_lo, b_hi, IDESC, 0);  // overwrites the accumulator
#pragma unroll
                for (uint32_t t2 = 1; t2 < KS; ++t2)
                    mma_acc(taddr, a_lo + t2 * AS, a_hi, b_lo + 2 * t2, b_hi);
                for (uint32_t fh = 1; fh < p.stem_fh; ++fh) {
                    a_lo += STEM_NJ_K;   // next filter row: one staged input row further
                    b_lo += brow;        // next K row block of the resident filter
#pragma unroll
                    for (uint32_t t2 = 0; t2 < KS; ++t2)
                        mma_acc(taddr, a_lo + t2 * AS, a_hi, b_lo + 2 * t2, b_hi);
                }
                mma_commit(bar_empty + 8 * stage);
                mma_commit(bar_tfull + 8 * acc);
                soff += stage16;
                if (++stage == p.stages) {
                    stage = 0;
                    soff = 0;
                    phase ^= 1;
                }
                if (++acc == NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        };
        using U1 = std::integral_constant<uint32_t, 1>;
        using U2 = std::integral_constant<uint32_t, 2>;
        using U3 = std::integral_constant<uint32_t, 3>;
        using U4 = std::integral_constant<uint32_t, 4>;
        if (elect_one()) {
            const uint32_t ksteps = p.stem4 ? (p.stem_fw2 + 1) / 2 : p.stem_fw2;
            if (p.stem4) {
                if (ksteps == 1) run(U1{}, U2{});
                else run(U2{}, U2{});
            } else {
                if (ksteps == 1) run(U1{}, U1{});
                else if (ksteps == 2) run(U2{}, U1{});
                else if (ksteps == 3) run(U3{}, U1{});
                else run(U4{}, U1{});
            }
        }
        __syncwarp();
    } else if (warp == 0 && rows) {
        // ============================ TMA producer, rows mode ============================
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_bres, bres_bytes);
            for (uint32_t tap = 0; tap < p.num_taps; ++tap)
                for (uint32_t kc = 0; kc < p.num_kc; ++kc)
                    tma_load_3d(
                            sBres + (tap * p.num_kc + kc) * BN * p.kb_bytes, &p.tmB, bar_bres,
                            static_cast<int32_t>(kc * p.kb_elems), static_cast<int32_t>(tap), 0);
        }
        uint32_t stage = 0, phase = 0;
        for (uint32_t ct = t_first; ct < t_end; ct += t_step) {
            const uint32_t tile = ct;  // C == 1 in this mode
            const uint32_t img = tile / p.tiles_per_img, t = tile % p.tiles_per_img;
            const uint32_t r0 = (t * IGEMM_BLOCK_M) / p.PW;  // first (padded) input row of the slab
            TIMED_WAIT(0, bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
                const uint32_t full = bar_full + 8 * stage;
                const uint32_t sa = sStage + stage * stage_bytes;
                mbar_arrive_expect_tx(full, stage_bytes);
                for (uint32_t kc = 0; kc < p.num_kc; ++kc)
                    tma_load_4d(
                            sa + kc * plane_bytes, &p.tmA, full, static_cast<int32_t>(kc * p.kb_elems),
                            -p.pad_w, static_cast<int32_t>(r0) - p.pad_h, static_cast<int32_t>(img));
            }
            __syncwarp();
            if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1 && rows) {
        // ============================ MMA issuer, rows mode ============================
        // tap (fh, fw) of output m' reads slab entry m' - r0*PW + fh*PW + fw: a pure row shift of the
        // A descriptor (RB-byte rows, RB = 128 / 64 with SWIZZLE_128B / 64B keyed on absolute smem address
        // bits).  The slab row a tap starts on is usually not a multiple of 8: the swizzle is keyed on
        // absolute address bits (as TMA wrote it), so a plain row-shifted start address is correct
        // (base-offset field 0; verified on B200 against the oracle).  One thread issues; the RB / 32 MMAs
        // of a (tap, channel block) are unrolled with immediate offsets and a constant accumulate predicate.
        if (elect_one()) {
            auto run = [&](auto NM_) {
            constexpr uint32_t NM = decltype(NM_)::value;   // MMAs (32 bytes of K) per (tap, channel block)
            constexpr uint32_t RB = NM * 32;                // bytes per slab / filter row
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const uint32_t ab_hi = desc_hi(8 * RB, NM == 4 ? 2u : 4u);
            const uint32_t plane16 = plane_bytes >> 4, stage16 = stage_bytes >> 4, brow = BN * RB / 16u;
            const uint32_t a0 = desc_lo(sStage, 16), b0 = desc_lo(sBres, 16);
            const uint32_t FH = p.num_taps / p.FW, roww = p.PW * (RB / 16);
            const bool fast33 = FH == 3 && p.FW == 3 && p.num_kc == 1;
            uint32_t soff = 0;
            auto mma4 = [&](uint32_t taddr, uint32_t a_lo, uint32_t b_lo) {
#pragma unroll
                for (uint32_t k = 0; k < NM; ++k) {
                    const uint64_t adesc = (static_cast<uint64_t>(ab_hi) << 32) | (a_lo + 2 * k);
                    const uint64_t bdesc = (static_cast<uint64_t>(ab_hi) << 32) | (b_lo + 2 * k);
                    if constexpr (KIND == KIND_S8)
                        mma_i8_acc(taddr, adesc, bdesc, IDESC);
                    else
                        mma_f16_acc(taddr, adesc, bdesc, IDESC);
                }
            };
            TIMED_WAIT(6, bar_bres, 0);
            for (uint32_t ct = t_first; ct < t_end; ct += t_step) {
                const uint32_t t = ct % p.tiles_per_img;  // C == 1 in this mode: ct is the tile
                const uint32_t off0 = (t * IGEMM_BLOCK_M) % p.PW;
                TIMED_WAIT(1, bar_tempty + 8 * acc, acc_phase ^ 1);
                TIMED_WAIT(2, bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BN;
                uint32_t row_lo = a0 + soff + off0 * (RB / 16);  // tap (fh, 0); RB-byte rows = RB / 16 units
                uint32_t b_lo = b0;
                // the very first MMA of the tile overwrites the accumulator
                mma_issue<KIND>(taddr, row_lo, ab_hi, b_lo, ab_hi, IDESC, 0);
#pragma unroll
                for (uint32_t k = 1; k < NM; ++k)
                    mma_issue<KIND>(taddr, row_lo + 2 * k, ab_hi, b_lo + 2 * k, ab_hi, IDESC, 1);
                if (fast33) {
                    // 3x3 window, one channel block (ResNet l1 3x3): straight-line issue.  The generic loops below
                    // cost ~34 uniform-datapath instructions per tap (three constant-bank loads, three branches, the
                    // counters), ~300 clk for the 4 x 48 clk of tensor work of a 128 x 64 tap: the issuing thread, not
                    // the tensor pipe, bounded the layer at ~80 clk per MMA (ncu instruction profile + tools/microbench/
                    // mma_shift.cu, profiles/r02/microbench_mma_shift.txt)
#pragma unroll
                    for (uint32_t t = 1; t < 9; ++t)
                        mma4(taddr, row_lo + (t / 3) * roww + (t % 3) * (RB / 16), b0 + t * brow);
                } else {
                bool skip = true;  // (fh, fw, kc) == (0, 0, 0) is done
                for (uint32_t fh = 0; fh < FH; ++fh) {
                    uint32_t tap_lo = row_lo;
                    for (uint32_t fw = 0; fw < p.FW; ++fw) {
                        uint32_t a_lo = tap_lo;
                        for (uint32_t kc = 0; kc < p.num_kc; ++kc) {
                            if (!skip)
                                mma4(taddr, a_lo, b_lo);
                            skip = false;
                            a_lo += plane16;
                            b_lo += brow;
                        }
                        tap_lo += RB / 16;  // next pixel
                    }
                    row_lo += roww;  // next input row
                }
                }
                mma_commit(bar_empty + 8 * stage);
                mma_commit(bar_tfull + 8 * acc);
                soff += stage16;
                if (++stage == p.stages) {
                    stage = 0;
                    soff = 0;
                    phase ^= 1;
                }
                if (++acc == NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            };
            if (p.kb_bytes == 128)
                run(std::integral_constant<uint32_t, 4>{});
            else
                run(std::integral_constant<uint32_t, 2>{});
        }
        __syncwarp();
    } else if (warp == 0) {
        // ============================ TMA producer (A, B) ============================
        uint32_t stage = 0, phase = 0;
        if (bres_tiled && t_first < t_end && elect_one()) {
            // resident filter tile: every tile of this CTA has N index t_first % n_tiles (grid % n_tiles == 0)
            const int32_t n0 = static_cast<int32_t>((t_first % p.n_tiles) * BN);
            mbar_arrive_expect_tx(bar_bres, bres_bytes);
            for (uint32_t kc = 0; kc < p.num_kc; ++kc)
                tma_load_3d(sBres + kc * BN * p.kb_bytes, &p.tmB, bar_bres, static_cast<int32_t>(kc * p.kb_elems), 0, n0);
        }
        __syncwarp();
        for (uint32_t ct = t_first; ct < t_end; ct += t_step) {
            // a phantom tile (M tiles not a multiple of C) re-loads the last valid tile: every CTA of the
            // cluster must run the same k-loop because the filter slices are exchanged in lock step
            const uint32_t m_raw = (ct / p.n_tiles) * C + crank, n_blk = ct % p.n_tiles;
            const uint32_t m_blk = m_raw < p.m_tiles ? m_raw : p.m_tiles - 1;
            const int32_t m0 = static_cast<int32_t>(m_blk * IGEMM_BLOCK_M);
            const int32_t n0 = static_cast<int32_t>(n_blk * BN);
            const uint32_t img = static_cast<uint32_t>(m0) / p.OHW;
            const uint32_t rem = static_cast<uint32_t>(m0) % p.OHW;
            const int32_t base_h = static_cast<int32_t>(rem / p.OW) * p.stride_h - p.pad_h;
            const int32_t base_w = static_cast<int32_t>(rem % p.OW) * p.stride_w - p.pad_w;
            uint32_t fh = 0, fw = 0;
            for (uint32_t tap = 0; tap < p.num_taps; ++tap) {
                // mode CONVOLUTION reads the spatially flipped filter tap for the same src pixel
                const uint32_t btap = p.flip ? (p.num_taps - 1 - tap) : tap;
                for (uint32_t kc = 0; kc < p.num_kc; ++kc) {
                    TIMED_WAIT(0, bar_empty + 8 * stage, phase ^ 1);
                    DBG_MARK(tissue__);   // slot 12: time inside the TMA issue block (back-pressure of the TMA unit)
                    if (elect_one()) {
                        const uint32_t full = bar_full + 8 * stage;
                        const uint32_t sa = sStage + stage * stage_bytes;
                        const uint32_t sb = sa + a_bytes;
                        const int32_t c0 = static_cast<int32_t>(kc * p.kb_elems);
                        if (pair) {
                            // both CTAs' loads complete on the LEADER's barrier, which expects both stages
                            const uint32_t lfull = mapa(full, 0);
                            if (crank == 0)
                                mbar_arrive_expect_tx(full, 2 * stage_bytes);
                            if (p.a_mode == 0) {
                                tma_load_2d_2sm(sa, &p.tmA, lfull, c0, m0);
                            } else {
                                tma_load_im2col_4d_2sm(
                                        sa, &p.tmA, lfull, c0, base_w, base_h, static_cast<int32_t>(img),
                                        static_cast<uint16_t>(fw * p.dil_w),
                                        static_cast<uint16_t>(fh * p.dil_h));
                            }
                            tma_load_3d_2sm(
                                    sb, &p.tmB, lfull, c0, static_cast<int32_t>(btap),
                                    n0 + static_cast<int32_t>(crank * b_rows));
                        }
                        if (!pair) {
                        mbar_arrive_expect_tx(full, stage_bytes);
                        if (p.a_mode == 0) {
                            tma_load_2d(sa, &p.tmA, full, c0, m0);
                        } else {
                            tma_load_im2col_4d(
                                    sa, &p.tmA, full, c0, base_w, base_h, static_cast<int32_t>(img),
                                    static_cast<uint16_t>(fw * p.dil_w),
                                    static_cast<uint16_t>(fh * p.dil_h));
                        }
                        if (bres_tiled) {
                            // (filter resident: the stage is the activation tile alone)
                        } else if (C == 1) {
                            tma_load_3d(sb, &p.tmB, full, c0, static_cast<int32_t>(btap), n0);
                        } else {
                            // this CTA's 1/C of the filter tile's rows, delivered to every CTA of the cluster
                            const uint32_t rows_c = BN / C;
                            tma_load_3d_mc(
                                    sb + crank * rows_c * p.kb_bytes, &p.tmB, full, c0,
                                    static_cast<int32_t>(btap), n0 + static_cast<int32_t>(crank * rows_c), cmask);
                        }
                        }
                    }
                    __syncwarp();
                    DBG_ADD(12, tissue__);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++fw == p.FW) {
                    fw = 0;
                    ++fh;
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        // ONE thread runs the whole loop.  tcgen05.mma is asynchronous, but the issuing thread must stay
        // ahead of the tensor pipe: a 128 x 64 MMA occupies it 48 clk, a 128 x 128 one 64 clk
        // (tools/microbench/mma_shapes.cu), so every instruction between two MMAs counts.  Hence: no
        // per-k-block elect / warp sync, descriptors advanced by one add, the MMAs of a k-block unrolled
        // with immediate offsets and a constant accumulate predicate, commit variants chosen outside.
        // UMMA layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B; SBO = 8 rows
        const uint32_t layout = p.kb_bytes == 128 ? 2u : (p.kb_bytes == 64 ? 4u : 6u);
        const uint32_t ab_hi = desc_hi(8 * p.kb_bytes, layout);
        // pair: M = 256 in the instruction descriptor, only the leader CTA issues
        constexpr uint32_t IDESC_PAIR = (IDESC & ~(0x1Fu << 24)) | ((256u >> 4) << 24);
        constexpr uint32_t IDESC_USE = pair ? IDESC_PAIR : IDESC;
        auto issue_first = [&](uint32_t taddr, uint32_t a_lo, uint32_t b_lo, uint32_t accum) {
            if constexpr (pair)
                mma_issue_2sm<KIND>(taddr, a_lo, ab_hi, b_lo, ab_hi, IDESC_USE, accum);
            else
                mma_issue<KIND>(taddr, a_lo, ab_hi, b_lo, ab_hi, IDESC_USE, accum);
        };
        auto issue_acc = [&](uint32_t taddr, uint32_t a_lo, uint32_t b_lo) {
            const uint64_t adesc = (static_cast<uint64_t>(ab_hi) << 32) | a_lo;
            const uint64_t bdesc = (static_cast<uint64_t>(ab_hi) << 32) | b_lo;
            if constexpr (pair) {
                if constexpr (KIND == KIND_S8)
                    mma_i8_2sm_acc(taddr, adesc, bdesc, IDESC_USE);
                else
                    mma_f16_2sm_acc(taddr, adesc, bdesc, IDESC_USE);
            } else {
                if constexpr (KIND == KIND_S8)
                    mma_i8_acc(taddr, adesc, bdesc, IDESC_USE);
                else
                    mma_f16_acc(taddr, adesc, bdesc, IDESC_USE);
            }
        };
        auto run = [&](auto NM_) {
            constexpr uint32_t NM = decltype(NM_)::value;  // MMAs (32 bytes of K each) per k-block
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const uint32_t stage16 = stage_bytes >> 4;     // descriptor units are 16 bytes
            const uint32_t a0 = desc_lo(sStage, 16);
            const uint32_t b0 = bres_tiled ? desc_lo(sBres, 16) : desc_lo(sStage + a_bytes, 16);
            const uint32_t bres16 = (BN * p.kb_bytes) >> 4;  // resident filter: one k-block of the tile
            uint32_t soff = 0;
            if (bres_tiled && t_first < t_end)
                TIMED_WAIT(6, bar_bres, 0);
            for (uint32_t ct = t_first; ct < t_end; ct += t_step) {
                TIMED_WAIT(1, bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BN;
                uint32_t accum = 0;  // the tile's first MMA overwrites the accumulator
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    TIMED_WAIT(2, bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_lo = a0 + soff, b_lo = b0 + (bres_tiled ? kb * bres16 : soff);
                    issue_first(taddr, a_lo, b_lo, accum);
                    accum = 1;
#pragma unroll
                    for (uint32_t k = 1; k < NM; ++k)
                        issue_acc(taddr, a_lo + 2 * k, b_lo + 2 * k);
                    const uint32_t be = bar_empty + 8 * stage;
                    if constexpr (pair) {
                        mma_commit_2sm_mc(be, 3);  // frees the stage in both CTAs
                        if (kb == num_kb - 1)
                            mma_commit_2sm_mc(bar_tfull + 8 * acc, 3);  // both CTAs' epilogues
                    } else {
                        if (C == 1)
                            mma_commit(be);
                        else
                            mma_commit_mc(be, cmask);  // frees the stage in every CTA
                        if (kb == num_kb - 1)
                            mma_commit(bar_tfull + 8 * acc);
                    }
                    soff += stage16;
                    if (++stage == p.stages) {
                        stage = 0;
                        soff = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        };
        // (elect.sync, not `lane == 0`: the compiler then knows that exactly one lane is active and emits
        // plain UTCHMMA / UTCBAR instead of a per-instruction ELECT + R2UR.BROADCAST loop)
        if (!(pair && crank != 0) && elect_one()) {
            if (p.kb_bytes == 128)
                run(std::integral_constant<uint32_t, 4>{});
            else if (p.kb_bytes == 64)
                run(std::integral_constant<uint32_t, 2>{});
            else
                run(std::integral_constant<uint32_t, 1>{});
        }
        __syncwarp();
    } else if (warp == 2) {
        // ============================ z prefetch (TMA) ============================
        // Every epilogue group owns a private sub-ring of ZD = nz / GROUPS chunks: slot = group * ZD + (the
        // group's chunk count % ZD).  A slot is then only ever consumed by ONE group, in order -- with a
        // single ring shared by all groups a slot's consumer alternates, a fast group can wait for a slot
        // a whole barrier phase ahead of it, and mbarrier parity waits alias one phase back (this hung /
        // faulted l1.c3 with a 6-deep ring; compute-sanitizer racecheck flags the cross-group slot reuse).
        if (p.has_z) {
            const uint32_t ZD = nz / T::GROUPS;
            uint32_t zcnt = 0, zph = 0;  // per group: 4-bit slot counters / 1-bit phases, packed
            uint32_t first = 0, ts = 0;
            for (uint32_t ct = t_first; ct < t_end; ct += t_step, first = (first + NUM_CHUNKS) % T::GROUPS, ++ts) {
                // Just-in-time z (z_jit): a z slot is occupied from its TMA issue until the epilogue has stored the
                // chunk; requested a ring depth ahead it sits idle for two tile periods, and on the layers whose
                // k-loop is long enough to want every kilobyte for A/B stages (ResNet l3.c3 / l4.c3) that idle
                // residency is what caps the tile rate.  Here the request waits for the PREVIOUS tile's accumulator
                // (tfull): one tile of lead.  No parity aliasing: tfull[acc] flips again only after epilogue(ts - 1)
                // handed the accumulator back, which needs the z chunk issued just before this wait to have landed.
                if (p.z_jit && ts > 0) {
                    const uint32_t prev = ts - 1;
                    mbar_wait(bar_tfull + 8 * (prev % NACC), (prev / NACC) & 1u);
                }
                const uint32_t m_blk = (ct / p.n_tiles) * C + crank, n_blk = ct % p.n_tiles;
                if (m_blk >= p.m_tiles)
                    continue;  // phantom tile
                const int32_t m0 = static_cast<int32_t>(m_blk * IGEMM_BLOCK_M);
                for (uint32_t c = 0; c < NUM_CHUNKS; ++c) {
                    const uint32_t col0 = n_blk * BN + c * T::EPI_N;
                    if (col0 >= p.OC)
                        break;
                    const uint32_t g = (first + c) % T::GROUPS;
                    const uint32_t zi = (zcnt >> (4 * g)) & 0xfu;
                    const uint32_t zb = g * ZD + zi;
                    TIMED_WAIT(5, bar_zempty + 8 * zb, ((zph >> g) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_zfull + 8 * zb, IGEMM_BLOCK_M * T::ROW_BYTES);
                        if (stem)
                            tma_load_3d(
                                    sZ + zb * T::ZBUF_BYTES, &p.tmZ, bar_zfull + 8 * zb,
                                    static_cast<int32_t>(col0),
                                    static_cast<int32_t>((m_blk % p.nseg) * IGEMM_BLOCK_M),
                                    static_cast<int32_t>(m_blk / p.nseg));
                        else
                            tma_load_2d(
                                    sZ + zb * T::ZBUF_BYTES, &p.tmZ, bar_zfull + 8 * zb,
                                    static_cast<int32_t>(col0), m0);
                    }
                    __syncwarp();
                    if (zi + 1 == ZD) {
                        zcnt &= ~(0xfu << (4 * g));
                        zph ^= 1u << g;
                    } else {
                        zcnt += 1u << (4 * g);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ============================ epilogue ============================
        // warp e = warp - 4: TMEM lane quadrant (e & 3) == rows [32*(e&3), +32) of the tile;
        // group (e >> 2) handles the chunks c with (tile sequence * NUM_CHUNKS + c) % GROUPS == group.
        const uint32_t e = warp - 4;
        const uint32_t quad = warp & 3;
        const uint32_t group = e >> 2;
        const uint32_t row = quad * 32 + lane;            // row of the tile == TMEM lane
        const uint32_t slab_u32 = smem_base + OFF_SD + e * NSLAB * T::SLAB_BYTES;
        uint8_t* slab_gen = smem_gen + OFF_SD + e * NSLAB * T::SLAB_BYTES;
        const uint8_t* sZ_gen = smem_gen + OFF_SZ;
        // RELU: x <= 0 ? 0 : x;  IDENTITY: the same select against -inf (a no-op) -- branch-free
        // TMEM hand-back: to the MMA warp of this CTA, or (pair) of the leader CTA
        const bool remote_tempty = pair && crank != 0;
        auto arrive_tempty = [&](uint32_t bar) {
            if (remote_tempty)
                mbar_arrive_cluster(mapa(bar, 0));
            else
                mbar_arrive(bar);
        };
        const float act_lo = p.nonline == B200CB_NL_RELU ? 0.f : -INFINITY;  // (int8 generic path)
        const bool generic_act = p.nonline == B200CB_NL_H_SWISH || p.nonline == B200CB_NL_SIGMOID;
        const bool fast8 = KIND == KIND_S8 && !generic_act && p.rq.fast_div == 2;
        uint32_t acc = 0, acc_phase = 0, db = 0, cur_key = 0xffffffffu;
        bool magic = false;  // (int8) the current bias slice carries I2F_MAGIC_BITS
        // this group's z sub-ring (see the z prefetch warp): position and phase advance once per chunk the
        // group consumes
        const uint32_t ZD = nz / T::GROUPS;
        uint32_t zi = 0, zphase = 0;
        auto next_acc = [&]() {
            if (++acc == NACC) {
                acc = 0;
                acc_phase ^= 1;
            }
        };
        // (m, n) tile indices advance incrementally: a division per tile is a ~25-instruction MUFU chain on
        // the critical path of short tiles (10 % of the low-K pointwise layers' samples)
        uint32_t tq = t_first / p.n_tiles, tr = t_first % p.n_tiles;
        const uint32_t dq = t_step / p.n_tiles, dr = t_step % p.n_tiles, ntl = p.n_tiles;
        auto next_tile = [&]() {
            tq += dq, tr += dr;
            if (tr >= ntl)
                tr -= ntl, ++tq;
        };
        // bias values of the group's current chunk (OUT32: of the whole tile), shared by the group's 4 warps
        uint32_t* sbias = reinterpret_cast<uint32_t*>(smem_gen + OFF_BIAS) + (T::GROUPS > 1 ? group * T::EPI_N : 0u);
        constexpr uint32_t BIAS_COLS = T::GROUPS > 1 ? T::EPI_N : BN;
        // group of the tile's chunk 0: (tile sequence number * NUM_CHUNKS) % GROUPS
        uint32_t first = 0;
        for (uint32_t ct = t_first; ct < t_end;
             ct += t_step, next_tile(), next_acc(), first = (first + NUM_CHUNKS) % T::GROUPS) {
            const uint32_t m_blk = tq * C + crank, n_blk = tr;
            const uint32_t n0 = n_blk * BN;
            // the chunks of this tile the group owns: c_own, c_own + GROUPS, ... (with 4 groups: at most one)
            const uint32_t c_own = (group + T::GROUPS - first) % T::GROUPS;
            if (c_own >= NUM_CHUNKS)
                continue;  // a tile of the other groups
            if (m_blk >= p.m_tiles) {
                // phantom tile: consume the accumulator and hand it straight back
                TIMED_WAIT(3, bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    arrive_tempty(bar_tempty + 8 * acc);
                continue;
            }
            const int32_t m0 = static_cast<int32_t>(m_blk * IGEMM_BLOCK_M);
            // chunks of this tile that exist (OC tail)
            uint32_t nchunks = 0;
#pragma unroll
            for (uint32_t c = 0; c < NUM_CHUNKS; ++c)
                if (n0 + c * T::EPI_N < p.OC)
                    nchunks = c + 1;
            // bias slice -> the group's smem (fp32 for float kinds, int32 for int8), only when it changes
            // (never again for layers whose N tiles map to a fixed chunk per CTA); the 4 warps of a group
            // walk the same tiles, a named barrier per group orders the refill against the readers
            const uint32_t key = T::GROUPS > 1 ? n_blk * NUM_CHUNKS + c_own : n_blk;
            if (key != cur_key) {
                const uint32_t b0 = n0 + (T::GROUPS > 1 ? c_own * T::EPI_N : 0u);
                if constexpr (KIND == KIND_S8 && !OUT32) {
                    // conversion-free int -> float (epilogue.cuh): legal for this chunk when K * 2^14 + max |bias|
                    // of its 64 columns stays below 2^22; every warp of the group derives the same flag
                    uint32_t mx = 0;
                    if (p.has_bias) {
                        const int32_t* bp = reinterpret_cast<const int32_t*>(p.bias);
                        for (uint32_t i = lane; i < BIAS_COLS; i += 32)
                            if (b0 + i < p.OC) {
                                const int32_t v = bp[b0 + i];
                                const uint32_t a = v < 0 ? 0u - static_cast<uint32_t>(v) : static_cast<uint32_t>(v);
                                mx = a > mx ? a : mx;
                            }
                    }
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    magic = fast8 && !p.nchw4 && static_cast<uint64_t>(p.acc_bound) + mx < I2F_RANGE;
                }
                named_bar_sync(1 + group, 128);
                for (uint32_t i = quad * 32 + lane; i < BIAS_COLS; i += 128) {
                    uint32_t v = 0;
                    if (p.has_bias && b0 + i < p.OC) {
                        if (OUT32 && p.out_f32) {
                            v = reinterpret_cast<const uint32_t*>(p.bias)[b0 + i];  // Float32 bias
                        } else if constexpr (KIND == KIND_S8) {
                            v = static_cast<uint32_t>(reinterpret_cast<const int32_t*>(p.bias)[b0 + i]);
                        } else if constexpr (KIND == KIND_BF16) {
                            v = static_cast<uint32_t>(
                                        reinterpret_cast<const unsigned short*>(p.bias)[b0 + i])
                                << 16;
                        } else {
                            v = __float_as_uint(
                                    __half2float(reinterpret_cast<const __half*>(p.bias)[b0 + i]));
                        }
                    }
                    sbias[i] = v + (magic ? I2F_MAGIC_BITS : 0u);
                }
                named_bar_sync(1 + group, 128);
                cur_key = key;
            }
            // last chunk this warp's group owns (decides when TMEM can be handed back)
            int32_t last_c = -1;
            for (uint32_t c = c_own; c < nchunks; c += T::GROUPS)
                last_c = static_cast<int32_t>(c);
            TIMED_WAIT(3, bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            if (last_c < 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    arrive_tempty(bar_tempty + 8 * acc);
            }
#pragma unroll 1
            for (uint32_t c = c_own; c < nchunks; c += T::GROUPS) {
                const uint32_t zb = group * ZD + zi;  // (has_z only)
                const uint32_t col0 = n0 + c * T::EPI_N;
                // slab db must have been drained by the TMA store issued NSLAB chunks ago
#ifdef B200CB_DEBUG_TIMING
                const long long tsw__ = clock64();
#endif
                // (elect.sync is deterministic for a given member mask: the same lane commits and waits
                // on the bulk groups, and the compiler emits plain UTMASTG instead of a per-lane R2UR loop)
                if (!p.has_z && !p.direct_store) {
                    if (elect_one())
                        tma_store_wait_read<NSLAB - 1>();
                    __syncwarp();
                }
#ifdef B200CB_DEBUG_TIMING
                dbg_acc[7] += clock64() - tsw__;
#endif
                if (p.has_z)
                    TIMED_WAIT(4, bar_zfull + 8 * zb, zphase);
                DBG_MARK(tbody__);
                // with z the outputs overwrite the warp's 32 rows of the z chunk (each thread reads and
                // writes only its own row's pieces), without z they go to the warp's private slab
                const uint8_t* zrow = sZ_gen + zb * T::ZBUF_BYTES + row * T::ROW_BYTES;
                uint8_t* drow = p.has_z ? const_cast<uint8_t*>(zrow)
                                        : slab_gen + db * T::SLAB_BYTES + lane * T::ROW_BYTES;
                const uint32_t st_src = p.has_z ? sZ + zb * T::ZBUF_BYTES + quad * T::SLAB_BYTES
                                                : slab_u32 + db * T::SLAB_BYTES;
                // 32 columns at a time; the accumulator is handed back to the MMA warp as soon as the group's
                // last columns are in registers
                constexpr uint32_t HALVES = T::EPI_N / 32;
                constexpr uint32_t HB = 1;  // halves loaded per batch (640 threads: <= 102 registers each)
                uint32_t rr[HB][32];
#pragma unroll
                for (uint32_t h = 0; h < HALVES; ++h) {
                    if (h % HB == 0) {
                        DBG_MARK(tld__);
#pragma unroll
                        for (uint32_t hh = 0; hh < HB; ++hh)
                            tmem_ld_32x32(
                                    tmem_base + ((quad * 32) << 16) + acc * BN + c * T::EPI_N + (h + hh) * 32, rr[hh]);
                        tmem_ld_wait();
                        DBG_ADD(8, tld__);
                        if (static_cast<int32_t>(c) == last_c && h + HB == HALVES) {
                            // accumulator fully drained into registers: hand TMEM back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0)
                                arrive_tempty(bar_tempty + 8 * acc);
                        }
                    }
                    if (col0 + h * 32 >= p.OC)
                        continue;  // 32 columns entirely past OC (OC = 16 / 96 / 144 ...): the store clips them
                    const uint32_t (&r)[32] = rr[h % HB];
                    const uint32_t* bcol = sbias + (T::GROUPS > 1 ? 0u : c * T::EPI_N) + h * 32;
                    constexpr uint32_t COLS_PER_16B = 16 / T::OUT_ES;
                    constexpr uint32_t PIECES = 32 / COLS_PER_16B;  // 16-byte pieces per 32 columns
                    if constexpr (OUT32) {
                      if (p.out_f32) {
                        // Float32 destination: x = acc (int8: dequantised with s_src * s_flt) + bias [+ z],
                        // activation, stored as fp32 (the fp32 ConvBias path and QuantizedS8 -> Float32)
#pragma unroll
                        for (uint32_t j = 0; j < PIECES; ++j) {
                            const uint32_t off = (h * PIECES + j) ^ (lane & 7);
                            const uint4 bv = *reinterpret_cast<const uint4*>(bcol + j * 4);
                            const uint32_t bb[4] = {bv.x, bv.y, bv.z, bv.w};
                            uint32_t zz[4] = {0, 0, 0, 0};
                            if (p.has_z) {
                                const uint4 zv = *reinterpret_cast<const uint4*>(zrow + off * 16);
                                zz[0] = zv.x, zz[1] = zv.y, zz[2] = zv.z, zz[3] = zv.w;
                            }
                            uint32_t ov[4];
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                float x;
                                if constexpr (KIND == KIND_S8)
                                    x = __fmul_rn(__int2float_rn(static_cast<int32_t>(r[j * 4 + k])), p.acc_scale);
                                else
                                    x = __uint_as_float(r[j * 4 + k]);
                                x = __fadd_rn(x, __uint_as_float(bb[k]));
                                if (p.has_z)
                                    x = __fadd_rn(x, __uint_as_float(zz[k]));
                                ov[k] = __float_as_uint(act_f(x, p.nonline));
                            }
                            *reinterpret_cast<uint4*>(drow + off * 16) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                        }
                      } else {
                        // QuantizedS32 destination: dst = acc + bias (IDENTITY, no z)
#pragma unroll
                        for (uint32_t j = 0; j < PIECES; ++j) {
                            const uint4 bv = *reinterpret_cast<const uint4*>(bcol + j * 4);
                            uint4 ov;
                            ov.x = r[j * 4 + 0] + bv.x;
                            ov.y = r[j * 4 + 1] + bv.y;
                            ov.z = r[j * 4 + 2] + bv.z;
                            ov.w = r[j * 4 + 3] + bv.w;
                            const uint32_t off = (h * PIECES + j) ^ (lane & 7);
                            *reinterpret_cast<uint4*>(drow + off * 16) = ov;
                        }
                      }
                    } else if constexpr (KIND == KIND_S8) {
                      if (p.nchw4) {
                        // this thread's output pixel (stem tiles are 128-pixel segments of one output row)
                        uint32_t img, pix;
                        bool px_ok;
                        if (stem) {
                            const uint32_t row3 = m_blk / p.nseg, ow = (m_blk % p.nseg) * IGEMM_BLOCK_M + row;
                            img = row3 / p.OH;
                            pix = (row3 % p.OH) * p.OW + ow;
                            px_ok = ow < p.OW;
                        } else {
                            const uint32_t m = static_cast<uint32_t>(m0) + row;
                            img = m / p.OHW;
                            pix = m - img * p.OHW;
                            px_ok = m < p.M;
                        }
                        const uint32_t col = col0 + h * 32;
                        uint32_t nwords = 0;
                        if (px_ok && col < p.OC)
                            nwords = (p.OC - col) / 4 < 8 ? (p.OC - col) / 4 : 8;
                        const size_t w0 = (static_cast<size_t>(img) * (p.OC / 4) + col / 4) * p.OHW + pix;
                        uint32_t* dw = static_cast<uint32_t*>(p.dst) + w0;
                        const uint32_t* zw = static_cast<const uint32_t*>(p.z) + w0;
                        const uint32_t mode = !fast8 ? 0u : (p.nonline == B200CB_NL_RELU ? 2u : 1u);
#define B200CB_NCHW4_CASE(Z_, M_)                                                                          \
    requant_row32_nchw4<Z_, M_>(r, bcol, zw, dw, p.OHW, nwords, p.nonline, p.rq)
                        if (p.z_direct) {
                            if (mode == 2) B200CB_NCHW4_CASE(true, 2);
                            else if (mode == 1) B200CB_NCHW4_CASE(true, 1);
                            else B200CB_NCHW4_CASE(true, 0);
                        } else {
                            if (mode == 2) B200CB_NCHW4_CASE(false, 2);
                            else if (mode == 1) B200CB_NCHW4_CASE(false, 1);
                            else B200CB_NCHW4_CASE(false, 0);
                        }
#undef B200CB_NCHW4_CASE
                      } else if (fast8) {
#define B200CB_RQ_CASE(Z_, R_)                                                      \
    do {                                                                            \
        if (magic)                                                                  \
            requant_row32<Z_, R_, 1>(r, bcol, zrow, drow, h, lane, p.rq);           \
        else if (!Z_ && rows)   /* rows mode: conversion-pipe bound (epilogue.cuh) */\
            requant_row32<false, R_, 2>(r, bcol, zrow, drow, h, lane, p.rq);        \
        else                                                                        \
            requant_row32<Z_, R_, 0>(r, bcol, zrow, drow, h, lane, p.rq);           \
    } while (0)
                        if (p.has_z) {
                            if (p.nonline == B200CB_NL_RELU)
                                B200CB_RQ_CASE(true, true);
                            else
                                B200CB_RQ_CASE(true, false);
                        } else {
                            if (p.nonline == B200CB_NL_RELU)
                                B200CB_RQ_CASE(false, true);
                            else
                                B200CB_RQ_CASE(false, false);
                        }
#undef B200CB_RQ_CASE
                      } else {
                        float x[32];
#pragma unroll
                        for (uint32_t j = 0; j < PIECES; ++j) {
                            // 64-byte rows, SWIZZLE_64B: 16-byte piece index ^= (row >> 1) & 3
                            const uint32_t off = (h * PIECES + j) ^ ((lane >> 1) & 3);
                            uint4 zv = make_uint4(0, 0, 0, 0);
                            if (p.has_z)
                                zv = *reinterpret_cast<const uint4*>(zrow + off * 16);
                            const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                            for (uint32_t w = 0; w < 4; ++w) {
                                const uint4 bv = *reinterpret_cast<const uint4*>(bcol + j * 16 + w * 4);
                                const uint32_t bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                                for (uint32_t b = 0; b < 4; ++b) {
                                    const int32_t sfb = static_cast<int32_t>(r[j * 16 + w * 4 + b] + bb[b]);
                                    const int32_t zq = static_cast<int32_t>(
                                            static_cast<int8_t>((zw[w] >> (8 * b)) & 0xff));
                                    // no z: zq == 0 and s_z == 0, r + (+0) leaves r unchanged
                                    x[j * 16 + w * 4 + b] = dequant_acc(sfb, zq, p.rq);
                                }
                            }
                        }
                        if (generic_act) {
#pragma unroll
                            for (uint32_t i = 0; i < 32; ++i)
                                x[i] = act_generic(x[i], p.nonline);
                        } else {
#pragma unroll
                            for (uint32_t i = 0; i < 32; ++i)
                                x[i] = x[i] <= act_lo ? act_lo : x[i];
                        }
#pragma unroll
                        for (uint32_t j = 0; j < PIECES; ++j) {
                            const uint32_t off = (h * PIECES + j) ^ ((lane >> 1) & 3);
                            uint32_t ow[4];
#pragma unroll
                            for (uint32_t w = 0; w < 4; ++w) {
                                uint32_t packed = 0;
#pragma unroll
                                for (uint32_t b = 0; b < 4; ++b)
                                    packed |= (static_cast<uint32_t>(quant_s8(x[j * 16 + w * 4 + b], p.rq)) & 0xff)
                                              << (8 * b);
                                ow[w] = packed;
                            }
                            *reinterpret_cast<uint4*>(drow + off * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                        }
                      }
                    } else {
                        // 16-bit kinds.  With 16 epilogue warps the epilogue is bound by instruction issue, so
                        // the arithmetic is packed: fp32x2 adds (bias, z) on register pairs, one conversion per
                        // output pair, RELU as a packed 16-bit max on the converted pair (max and the rounding
                        // commute: both are monotonic and fix 0; NaN is propagated as the scalar select did)
                        f32x2_t v[16];
#pragma unroll
                        for (uint32_t j = 0; j < 8; ++j) {
                            const uint4 bv = *reinterpret_cast<const uint4*>(bcol + j * 4);
                            v[2 * j] = f2_add(
                                    f2_pack(__uint_as_float(r[j * 4 + 0]), __uint_as_float(r[j * 4 + 1])),
                                    f2_pack(__uint_as_float(bv.x), __uint_as_float(bv.y)));
                            v[2 * j + 1] = f2_add(
                                    f2_pack(__uint_as_float(r[j * 4 + 2]), __uint_as_float(r[j * 4 + 3])),
                                    f2_pack(__uint_as_float(bv.z), __uint_as_float(bv.w)));
                        }
                        // convert + activation + staging store, instantiated once per z / no-z path: a common tail after
                        // the `has_z` branch made ptxas shuffle all 32 value registers into one allocation at the merge
                        // point (28 MOVs per 32 columns, ncu instruction profile of the low-K pointwise layers)
                        auto finish16 = [&]() {
                        uint32_t o[16];
                        if (generic_act) {
#pragma unroll
                            for (uint32_t i = 0; i < 16; ++i) {
                                float lo, hi;
                                f2_unpack(v[i], lo, hi);
                                lo = act_generic(lo, p.nonline), hi = act_generic(hi, p.nonline);
                                o[i] = (KIND == KIND_BF16) ? pack_bf16x2(lo, hi) : pack_f16x2(lo, hi);
                            }
                        } else {
#pragma unroll
                            for (uint32_t i = 0; i < 16; ++i) {
                                float lo, hi;
                                f2_unpack(v[i], lo, hi);
                                o[i] = (KIND == KIND_BF16) ? pack_bf16x2(lo, hi) : pack_f16x2(lo, hi);
                            }
                            if (p.nonline == B200CB_NL_RELU) {
#pragma unroll
                                for (uint32_t i = 0; i < 16; ++i)
                                    o[i] = (KIND == KIND_BF16) ? relu_bf16x2(o[i]) : relu_f16x2(o[i]);
                            }
                        }
#pragma unroll
                        for (uint32_t j = 0; j < PIECES; ++j) {
                            const uint32_t off = (h * PIECES + j) ^ (lane & 7);
                            *reinterpret_cast<uint4*>(drow + off * 16) =
                                    make_uint4(o[j * 4], o[j * 4 + 1], o[j * 4 + 2], o[j * 4 + 3]);
                        }
                        };
                        if (p.has_z) {
#pragma unroll
                            for (uint32_t j = 0; j < PIECES; ++j) {
                                const uint32_t off = (h * PIECES + j) ^ (lane & 7);
                                const uint4 zv = *reinterpret_cast<const uint4*>(zrow + off * 16);
                                const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
                                for (uint32_t w = 0; w < 4; ++w) {
                                    if constexpr (KIND == KIND_BF16)
                                        v[j * 4 + w] = f2_add(v[j * 4 + w], f2_pack(bf16lo(zw[w]), bf16hi(zw[w])));
                                    else
                                        v[j * 4 + w] = f2_add(v[j * 4 + w], f2_pack(f16lo(zw[w]), f16hi(zw[w])));
                                }
                            }
                            finish16();
                        } else {
                            finish16();
                        }
                    }
                }
                DBG_ADD(9, tbody__);
                if (p.nchw4)
                    continue;  // written in place above, no slab / TMA store
                DBG_MARK(tst__);
                if (p.direct_store && !rows) {
                    // the warp's 32 staged rows -> global memory, 8 (4) lanes per 128- (64-) byte row, every
                    // store instruction covers whole rows; rows past M / OW and columns past OC are clipped
                    __syncwarp();
                    constexpr uint32_t PIECES_ROW = T::ROW_BYTES / 16;
                    constexpr uint32_t ROWS_PER_IT = 32 / PIECES_ROW;
                    const uint32_t piece = lane % PIECES_ROW;
                    const uint8_t* rd = p.has_z ? sZ_gen + zb * T::ZBUF_BYTES + quad * T::SLAB_BYTES
                                                : slab_gen + db * T::SLAB_BYTES;
                    const uint32_t col = col0 + piece * (16 / T::OUT_ES);
                    uint32_t lim, mbase;  // rows [0, lim) of the warp's 32 exist; mbase = their first pixel
                    if (stem) {
                        const uint32_t ow0 = (m_blk % p.nseg) * IGEMM_BLOCK_M + quad * 32;
                        lim = ow0 < p.OW ? p.OW - ow0 : 0u;
                        mbase = (m_blk / p.nseg) * p.OW + ow0;
                    } else {
                        mbase = static_cast<uint32_t>(m0) + quad * 32;
                        lim = mbase < p.M ? p.M - mbase : 0u;
                    }
                    uint8_t* out = static_cast<uint8_t*>(p.dst) +
                                   (static_cast<size_t>(mbase) * p.OC + col) * T::OUT_ES;
                    const size_t pitch = static_cast<size_t>(p.OC) * T::OUT_ES;
#pragma unroll
                    for (uint32_t it = 0; it < 32 / ROWS_PER_IT; ++it) {
                        const uint32_t r = it * ROWS_PER_IT + lane / PIECES_ROW;
                        const uint32_t sw = T::ROW_BYTES == 128 ? (piece ^ (r & 7)) : (piece ^ ((r >> 1) & 3));
                        const uint4 v = *reinterpret_cast<const uint4*>(rd + r * T::ROW_BYTES + sw * 16);
                        if (r < lim && col < p.OC)
                            *reinterpret_cast<uint4*>(out + r * pitch) = v;
                    }
                    __syncwarp();
                    if (p.has_z && lane == 0)
                        mbar_arrive(bar_zempty + 8 * zb);  // every lane has read the chunk's rows back
                    db ^= (NSLAB - 1);
                    if (++zi == ZD) {
                        zi = 0;
                        zphase ^= 1;
                    }
                    DBG_ADD(10, tst__);
                    continue;
                }
                if (rows) {
                    // rows mode: the slab's 32 consecutive m' span several output rows and skip the
                    // PW - OW padding columns, which a TMA box cannot express (negative store
                    // coordinates are illegal): copy the slab out with coalesced 16-byte stores,
                    // 8 lanes per 128-byte pixel row.
                    __syncwarp();
                    const uint32_t img = m_blk / p.tiles_per_img, t = m_blk % p.tiles_per_img;
                    const uint32_t mp = t * IGEMM_BLOCK_M + quad * 32;
                    constexpr uint32_t PIECES_ROW = T::ROW_BYTES / 16;   // 8 (128-byte rows) or 4
                    constexpr uint32_t ROWS_PER_IT = 32 / PIECES_ROW;
                    const uint32_t piece = lane % PIECES_ROW;
                    const uint8_t* slab_rd = slab_gen + db * T::SLAB_BYTES;
                    uint8_t* dst_img = static_cast<uint8_t*>(p.dst) +
                                       static_cast<size_t>(img) * p.OH * p.OW * p.OC * T::OUT_ES;
#pragma unroll
                    for (uint32_t it = 0; it < 32 / ROWS_PER_IT; ++it) {
                        const uint32_t r = it * ROWS_PER_IT + lane / PIECES_ROW;
                        const uint32_t mrow = mp + r;
                        const uint32_t oh = mrow / p.PW, owp = mrow % p.PW;
                        const uint32_t sw = T::ROW_BYTES == 128 ? (piece ^ (r & 7)) : (piece ^ ((r >> 1) & 3));
                        const uint4 v = *reinterpret_cast<const uint4*>(slab_rd + r * T::ROW_BYTES + sw * 16);
                        const uint32_t col = col0 + piece * (16 / T::OUT_ES);
                        if (owp < p.OW && oh < p.OH && col < p.OC)
                            *reinterpret_cast<uint4*>(
                                    dst_img + (static_cast<size_t>(oh * p.OW + owp) * p.OC + col) * T::OUT_ES) = v;
                    }
                    __syncwarp();
                    db ^= (NSLAB - 1);
                    continue;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    if (stem)
                        tma_store_3d(
                                &p.tmD, st_src, static_cast<int32_t>(col0),
                                static_cast<int32_t>((m_blk % p.nseg) * IGEMM_BLOCK_M + quad * 32),
                                static_cast<int32_t>(m_blk / p.nseg));
                    else
                        tma_store_2d(
                                &p.tmD, st_src, static_cast<int32_t>(col0),
                                m0 + static_cast<int32_t>(quad * 32));
                    tma_store_commit();
                    if (p.has_z) {
                        // the z chunk is free again once the store has read this warp's rows
                        DBG_MARK(tzr__);
                        tma_store_wait_read<0>();
                        DBG_ADD(11, tzr__);
                        mbar_arrive(bar_zempty + 8 * zb);
                    }
                }
                db ^= (NSLAB - 1);
                if (++zi == ZD) {
                    zi = 0;
                    zphase ^= 1;
                }
                DBG_ADD(10, tst__);
            }
        }
        if (elect_one())
            tma_store_wait<0>();
    }

#ifdef B200CB_DEBUG_TIMING
    if (p.dbg && lane == 0 && (warp == 0 || warp == 1 || warp == 4)) {
        long long* o = p.dbg + (static_cast<size_t>(blockIdx.x) * 3 + (warp == 4 ? 2 : warp)) * 17;
        for (int i = 0; i < 16; ++i)
            o[i] = dbg_acc[i];
        o[16] = clock64() - dbg_t0;
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (C > 1)
        cluster_sync_all();  // no CTA may exit while a peer can still signal its barriers
    if (warp == 2) {
        tc_fence_after();
        if (pair)
            tmem_dealloc_2sm(tmem_base, TMEM_COLS);
        else
            tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// Re-packs a KRSC filter for the W-folded stem (see file header): (OC, FH, FW, IC) ->
// (OC, FH, FW2, 2*IC) where folded tap t2, channel half hf holds original tap fw = 2*t2 + hf - shift
// (zero when outside [0, FW)).  One thread per output element.
template <typename E>
__global__ void fold_filter_kernel(
        const E* __restrict__ src, E* __restrict__ dst, uint32_t OC, uint32_t FH, uint32_t FW,
        uint32_t IC, uint32_t FW2, uint32_t shift) {
    const uint32_t total = OC * FH * FW2 * 2 * IC;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t c2 = i % (2 * IC);
        const uint32_t t2 = (i / (2 * IC)) % FW2;
        const uint32_t fh = (i / (2 * IC * FW2)) % FH;
        const uint32_t oc = i / (2 * IC * FW2 * FH);
        const int32_t fw = static_cast<int32_t>(2 * t2 + c2 / IC) - static_cast<int32_t>(shift);
        E v = E(0);
        if (fw >= 0 && fw < static_cast<int32_t>(FW))
            v = src[((static_cast<size_t>(oc) * FH + fh) * FW + fw) * IC + c2 % IC];
        dst[i] = v;
    }
}

// Self-test 0: sweeps every fp32 dividend against a set of scales and counts the cases where the
// fast division of the requantize epilogue differs from IEEE division (a) as a quotient, when the
// quotient is below the saturation guard, or (b) after roundf + clamp to int8.
__global__ void selftest_div_kernel(const float* scales, int nscales, unsigned long long* bad) {
    unsigned long long local = 0;
    for (int si = 0; si < nscales; ++si) {
        RequantParams q;
        q.s_bias = 1.f, q.s_z = 0.f, q.s_dst = scales[si], q.inv_dst = 1.0f / scales[si], q.fast_div = 1;
        q.one = scales[4];  // 1.0f, read from memory
        RequantParams ref = q;
        ref.fast_div = 0;
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
             i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
            const float r = __uint_as_float(static_cast<uint32_t>(i));
            if (!isfinite(r))
                continue;
            const float a = div_by_scale(r, q), b = div_by_scale(r, ref);
            // (a) raw quotient: the FMA residual r - q0*s needs 48 more bits below r, so dividends
            // under ~2^-83 can underflow it (Markstein's theorem assumes no underflow); their
            // quotients are far below 0.5 and are covered by check (b)
            if (fabsf(b) < 524288.f && fabsf(r) > 1e-25f && __float_as_uint(a) != __float_as_uint(b))
                ++local;
            // (b) the value the epilogue actually stores
            if (quant_s8(r, q) != quant_s8(r, ref))
                ++local;
        }
    }
    if (local)
        atomicAdd(bad, local);
}

// Self-test 1: the ALU-only quantize of the fast int8 epilogue (unguarded division + magic-number
// rounding, both sign variants) against the generic sequence with IEEE division, over every fp32
// dividend whose quotient stays below 2^60 (the range make_requant admits for fast_div == 2).
__global__ void selftest_fastq_kernel(const float* scales, int nscales, unsigned long long* bad) {
    unsigned long long local = 0;
    for (int si = 0; si < nscales; ++si) {
        RequantParams q;
        q.s_bias = 1.f, q.s_z = 0.f, q.s_dst = scales[si], q.inv_dst = 1.0f / scales[si], q.fast_div = 2;
        q.one = scales[4];
        RequantParams ref = q;
        ref.fast_div = 0;
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
             i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
            const float r = __uint_as_float(static_cast<uint32_t>(i));
            if (!(fabsf(__fmul_rn(r, q.inv_dst)) < 1152921504606846976.f))
                continue;
            const uint32_t want = static_cast<uint32_t>(quant_s8(r, ref)) & 0xffu;
            if ((round_sat_s8_signed(div_by_scale_noguard(r, q)) & 0xffu) != want)
                ++local;
            if (r >= 0.f && (round_sat_s8_nonneg(div_by_scale_noguard(r, q)) & 0xffu) != want)
                ++local;
            // the single-conversion forms (cvt.{rzi,rmi}.sat.s8 after a round-toward-zero add of +-0.5)
            if ((round_sat_s8_signed_cvt(div_by_scale_noguard(r, q)) & 0xffu) != want)
                ++local;
            if (r >= 0.f && (round_sat_s8_nonneg_cvt(div_by_scale_noguard(r, q)) & 0xffu) != want)
                ++local;
        }
    }
    if (local)
        atomicAdd(bad, local);
}

// `grid` == 0: one CTA per SM (clusters: as many whole clusters as the device keeps co-resident, asked
// from the occupancy API once per kernel variant / cluster size -- GPC boundaries strand a few SMs for
// clusters of 4), capped by the number of tiles.
// Self-test 2: the packed fp32x2 requantize (both lanes, RELU and signed forms, with z) against the
// scalar fast sequence that self-test 1 pins to IEEE division, over all 2^32 accumulator values.
// (This test is what caught ptxas contracting mul.rn.f32x2 + add.rn.f32x2: see the Makefile.)
__global__ void selftest_packed_kernel(const float* scales, int nscales, unsigned long long* bad) {
    unsigned long long local = 0;
    for (int si = 0; si < nscales; ++si) {
        RequantParams q;
        q.s_bias = 0.0013f * (si + 1), q.s_z = 0.07f, q.s_dst = scales[si], q.inv_dst = 1.0f / scales[si];
        q.fast_div = 2;
        q.one = scales[4];  // h_scales[4] == 1.0f: a run-time value here as in the product kernels (a kernel parameter there)
        const RequantPacked kp = make_requant_packed(q);
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
             i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
            const int32_t a0 = static_cast<int32_t>(static_cast<uint32_t>(i));
            const int32_t a1 = static_cast<int32_t>(static_cast<uint32_t>(i * 2654435761ull));
            const uint32_t zw = static_cast<uint32_t>(i * 40503ull) ^ 0x80808080u;  // int8x4 word ^ 0x80
            const uint32_t zb0 = __byte_perm(zw, 0x4B000000u, 0x7540u), zb1 = __byte_perm(zw, 0x4B000000u, 0x7541u);
            auto scalar = [&](int32_t sfb, uint32_t zbits, bool relu) {
                float x = __fmul_rn(__int2float_rn(sfb), q.s_bias);
                x = __fadd_rn(x, __fmul_rn(__fadd_rn(__uint_as_float(zbits), -8388736.f), q.s_z));
                if (relu)
                    return round_sat_s8_nonneg(div_by_scale_noguard(fmaxf(x, 0.f), q)) & 0xffu;
                return round_sat_s8_signed(div_by_scale_noguard(x, q)) & 0xffu;
            };
            uint32_t lo, hi;
            requant_pair_nonneg<true>(a0, a1, zb0, zb1, kp, lo, hi);
            if ((lo & 0xffu) != scalar(a0, zb0, true) || (hi & 0xffu) != scalar(a1, zb1, true))
                ++local;
            requant_pair_signed<true>(a0, a1, zb0, zb1, kp, lo, hi);
            if ((lo & 0xffu) != scalar(a0, zb0, false) || (hi & 0xffu) != scalar(a1, zb1, false))
                ++local;
            // the conversion-free int32 -> float forms: full range (i2f_pair) against I2F.RN bit for bit, and the
            // range-proven magic add on the accumulators inside its range
            float c0, c1;
            f2_unpack(i2f_pair(a0, a1, kp), c0, c1);
            if (__float_as_uint(c0) != __float_as_uint(__int2float_rn(a0)) ||
                __float_as_uint(c1) != __float_as_uint(__int2float_rn(a1)))
                ++local;
            requant_pair_nonneg<true, 2>(a0, a1, zb0, zb1, kp, lo, hi);
            if ((lo & 0xffu) != scalar(a0, zb0, true) || (hi & 0xffu) != scalar(a1, zb1, true))
                ++local;
            const int32_t m0 = (a0 >> 10), m1 = (a1 >> 10);   // |m| < 2^21: inside the magic range
            requant_pair_signed<true, 1>(m0 + static_cast<int32_t>(I2F_MAGIC_BITS), m1 + static_cast<int32_t>(I2F_MAGIC_BITS),
                                         zb0, zb1, kp, lo, hi);
            if ((lo & 0xffu) != scalar(m0, zb0, false) || (hi & 0xffu) != scalar(m1, zb1, false))
                ++local;
        }
    }
    if (local)
        atomicAdd(bad, local);
}

template <int KIND, int BN, int OUT32, int PAIR>
cudaError_t launch_variant(
        const IgemmParams& p, uint32_t grid, uint64_t total_tiles, uint32_t smem, cudaStream_t stream) {
    auto kern = igemm_kernel<KIND, BN, OUT32, PAIR>;
    static bool attr_set[64] = {false};
    static int max_clusters[64][5] = {{0}};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(MAX_DYN_SMEM));
        if (e != cudaSuccess)
            return e;
        attr_set[dev] = true;
    }
    const uint32_t C = p.cluster;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(Traits<KIND, OUT32>::THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
#ifdef B200CB_TOOLS
    static const bool use_pdl = getenv("B200CB_NO_PDL") == nullptr;
#else
    constexpr bool use_pdl = true;
#endif
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = C;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = C > 1 ? 2 : 1;
    if (grid == 0) {
        grid = static_cast<uint32_t>(device_sm_count());
        if (C > 1) {
            if (max_clusters[dev][C] == 0) {
                cfg.gridDim = dim3(grid / C * C, 1, 1);
                int n = 0;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
                if (e != cudaSuccess)
                    return e;
                max_clusters[dev][C] = n > 0 ? n : 1;
            }
            const uint32_t fit = static_cast<uint32_t>(max_clusters[dev][C]) * C;
            grid = (grid < fit ? grid : fit) / C * C;
        }
    }
    const uint64_t cluster_tiles = (total_tiles + C - 1) / C;
    if (grid / C > cluster_tiles)
        grid = static_cast<uint32_t>(cluster_tiles) * C;
    else if (p.b_res)  // every tile of a CTA must share one N index: tile step == grid == k * n_tiles
        grid = grid < p.n_tiles ? p.n_tiles : grid / p.n_tiles * p.n_tiles;
    cfg.gridDim = dim3(grid, 1, 1);
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// (device, plan, tensor addresses) -> finished kernel parameters; process-wide, mutex-guarded, bounded
struct ParamKey {
    b200cb_plan plan;
    const void *src, *flt, *bias, *z;
    void* dst;
    int dev;
};
class ParamCache {
public:
    bool lookup(const ParamKey& k, IgemmParams* out) {
        const uint64_t h = hash(k);
        std::lock_guard<std::mutex> lk(m_mtx);
        auto it = m_map.find(h);
        if (it == m_map.end() || memcmp(&it->second.key, &k, sizeof(k)) != 0)
            return false;
        *out = it->second.params;
        return true;
    }
    void store(const ParamKey& k, const IgemmParams& p) {
        const uint64_t h = hash(k);
        std::lock_guard<std::mutex> lk(m_mtx);
        if (m_map.size() >= 4096)
            m_map.clear();
        Entry& e = m_map[h];
        e.key = k;
        e.params = p;
    }

private:
    struct Entry {
        ParamKey key;
        IgemmParams params;
    };
    static uint64_t hash(const ParamKey& k) {
        // FNV-1a over the descriptor part of the plan and the addresses
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void* p, size_t n) {
            const unsigned char* b = static_cast<const unsigned char*>(p);
            for (size_t i = 0; i < n; ++i)
                h = (h ^ b[i]) * 1099511628211ull;
        };
        mix(&k.plan.desc, sizeof(k.plan.desc));
        mix(&k.plan.algo, sizeof(k.plan.algo));
        mix(k.plan.reserved, sizeof(k.plan.reserved));
        mix(&k.src, sizeof(void*) * 5 + sizeof(int));
        return h;
    }
    std::mutex m_mtx;
    std::unordered_map<uint64_t, Entry> m_map;
};
ParamCache& param_cache() {
    static ParamCache c;
    return c;
}

// geometry actually fed to the kernel (differs from the descriptor when the conv is W-folded)
struct Geometry {
    uint32_t ic, iw, fw, stride_w;
    int32_t pad_w_lo, pad_w_hi;
    uint32_t fold_shift;  // zero taps prepended to the filter row
    uint32_t fw_packed;   // tap pairs stored per filter row in the packed filter
    bool folded;
    bool stem;            // a_mode 2: resident filter + one input-row slab per tile
    bool stem4;           // ... with a pixel pair per 16-byte unit (IC * es == 8)
    bool ring;            // ... staged as whole rows into the row ring (IgemmParams::stem_ring)
    uint32_t ring_nj;     // pixel pairs per staged row
    uint32_t ring_np;     // row pairs per output row
};
constexpr uint32_t STEM_NJ = STEM_NJ_K;  // 128 output pixels + up to 4 tap pairs - 1, rounded up to 8

// The NCHW4 cells run on an NHWC / KRSC copy of src / filter in the workspace (relayout.cu), with the
// channels zero-padded to a TMA-legal multiple of 16: this is the descriptor of that equivalent conv.
// Network stems (IC <= 8, stride 2, even width) pad to 8-byte pixels instead and take the stem path.
b200cb_desc nchw4_equivalent(const b200cb_desc& d) {
    b200cb_desc e = d;
    e.format = B200CB_FORMAT_NHWC;
    e.src_batch_stride = 0;
    const bool stem_like = d.ic <= 8 && d.stride_w == 2 && d.stride_h == 2 && d.dilate_w == 1 &&
                           d.dilate_h == 1 && d.iw % 2 == 0 && d.mode == B200CB_MODE_CROSS_CORRELATION;
    e.ic = stem_like ? 8 : (d.ic + 15) / 16 * 16;
    return e;
}
size_t align256(size_t v) {
    return (v + 255) & ~static_cast<size_t>(255);
}

Geometry make_geometry(const b200cb_desc& d, uint32_t in_es, bool allow_fold = true) {
    Geometry g;
    g.ic = d.ic, g.iw = d.iw, g.fw = d.fw, g.stride_w = d.stride_w;
    g.pad_w_lo = (int32_t)d.pad_w;
    g.pad_w_hi = (int32_t)d.pad_w;
    g.fold_shift = 0;
    g.fw_packed = d.fw;
    g.folded = false;
    g.stem = false;
    g.stem4 = false;
    g.ring = false;
    g.ring_nj = g.ring_np = 0;
    if (allow_fold && d.stride_w == 2 && d.dilate_w == 1 && d.iw % 2 == 0 && d.ic * in_es <= 16 &&
        d.mode == B200CB_MODE_CROSS_CORRELATION && d.src_batch_stride == 0) {
        // window of output ow covers pixels [2*ow - pad, 2*ow - pad + FW); extend it to an even start
        // and an even length: pixel pairs [ow - pad_lo2, ow - pad_lo2 + FW2)
        const uint32_t shift = d.pad_w & 1;
        const uint32_t fw_ext = d.fw + shift;
        const uint32_t fw2 = (fw_ext + 1) / 2;
        const int32_t pad_lo2 = (int32_t)((d.pad_w + shift) / 2);
        const int32_t iw2 = (int32_t)(d.iw / 2);
        const int32_t pad_hi2 = (int32_t)d.ow + (int32_t)fw2 - 1 - iw2 - pad_lo2;
        g.ic = d.ic * 2, g.iw = d.iw / 2, g.fw = fw2, g.stride_w = 1;
        g.pad_w_lo = pad_lo2, g.pad_w_hi = pad_hi2;
        g.fold_shift = shift;
        g.fw_packed = fw2;
        g.folded = true;
        // stem mode: every tap pair of a filter row in one 128-byte K row (4 pairs), the whole
        // filter resident in shared memory (<= 64 output channels, <= 8 filter rows)
        if ((d.ic * in_es == 16 || d.ic * in_es == 8) && fw2 <= 4 && d.fh <= 8 && d.oc <= 64 &&
            d.stride_h == 2 && d.dilate_h == 1 && pad_lo2 <= 64) {
            g.stem = true;
            g.stem4 = d.ic * in_es == 8;
            g.fw_packed = 4;
            // row ring: one 128-pixel segment covers the output row, a staged row (OW + FW2 - 1 pixel pairs
            // of 16 bytes) fits one TMA box of <= 256 8-byte elements, <= 4 row pairs per output row
            // (rows of 8 pixel pairs = 128 bytes: a TMA box must land on a 128-byte aligned address)
            const uint32_t nj = (d.ow + fw2 - 1 + 7) / 8 * 8, np = (d.fh + (d.pad_h & 1) + 1) / 2;
            // (16-bit kinds: conv1 0.214 -> 0.159 ms; the int8 stem is epilogue bound and measured 15 % slower
            // on contiguous row ranges, profiles/r02)
            bool ring_ok = g.stem4 && in_es == 2 && d.ow <= IGEMM_BLOCK_M && nj <= 128 && np <= 4 && d.has_z == 0;
#ifdef B200CB_TOOLS
            if (const char* e = getenv("B200CB_RING"))
                ring_ok = ring_ok && atoi(e) != 0;
#endif
            if (ring_ok) {
                g.ring = true;
                g.ring_nj = nj;
                g.ring_np = np;
            }
        }
    }
    return g;
}

// rows mode (a_mode 3) geometry: padded width, slab rows, tiles per image
struct RowsGeom {
    uint32_t pw, rh, tiles_per_img, num_kc, a_stage, bres;
    bool ok;
};
RowsGeom rows_geometry(const b200cb_desc& d, uint32_t in_es, uint32_t block_n) {
    RowsGeom r{};
    r.ok = false;
    if (d.stride_h != 1 || d.stride_w != 1 || d.dilate_h != 1 || d.dilate_w != 1)
        return r;
    if (d.fh * d.fw == 1 || d.mode != B200CB_MODE_CROSS_CORRELATION || d.has_z)
        return r;
    // channel rows of 128 bytes (SWIZZLE_128B), or of 64 bytes (SWIZZLE_64B: QuantizedS8 IC = 64, ResNet l1 3x3)
    const uint32_t rb = (d.ic * in_es) % 128 == 0 ? 128u : 64u;
    if ((rb == 64 && d.ic * in_es != 64) || d.oc > block_n)
        return r;
    r.pw = d.iw + 2 * d.pad_w;
    if (r.pw != d.ow + d.fw - 1 || r.pw > 256)
        return r;
    r.rh = (r.pw - 1 + (d.fh - 1) * r.pw + d.fw - 1 + 127) / r.pw + 1;
    if (r.rh > 256)
        return r;
    r.num_kc = d.ic * in_es / rb;
    r.tiles_per_img = (d.oh * r.pw + IGEMM_BLOCK_M - 1) / IGEMM_BLOCK_M;
    r.a_stage = r.num_kc * r.rh * r.pw * rb;
    r.bres = d.fh * d.fw * r.num_kc * block_n * rb;
    r.ok = true;
    return r;
}

// CTA-pair heuristic: a single CTA's 128 x BN MMA is bound by shared-memory bandwidth (operand reads
// plus TMA writes: 192 B/clk at BN = 256, 256 B/clk at BN = 128, against 128 B/clk); the pair halves
// the filter-side traffic.  It pays once the k-loop is long enough for the main loop, not the epilogue
// or HBM, to bound the tile.
uint32_t pick_pair(const b200cb_desc& d, uint32_t block_n, uint32_t m_tiles) {
    const uint32_t in_es = d.src_dtype == B200CB_DT_QS8 ? 1 : 2;
    const uint64_t k_bytes = (uint64_t)d.fh * d.fw * d.ic * in_es;
    (void)block_n;
    // (few M tiles are no reason not to pair: 98 tiles of ResNet-50 layer4 run 10-20 % faster as 49 pairs,
    // tools/sweep_resnet.py r02)
    if (m_tiles < 8)
        return 0;
    // measured crossover (tools/sweep_resnet.py): long k-loops always; 1 KiB k-loops only for wide
    // 16-bit tiles without the z stream; shorter ones are epilogue / HBM bound and lose to the pairing
    if (k_bytes >= 2048)
        return 1;
    return (in_es == 2 && k_bytes >= 1024 && !d.has_z && d.oc >= 256) ? 1u : 0u;
}

// Direct-store epilogue (IgemmParams::direct_store): layers whose k-loop is short are bound by HBM and
// by the epilogue; there the TMA store's round trip through the (always full) TMA load queue is what
// the epilogue warps wait for.
bool pick_direct_store(const b200cb_desc& d, bool nchw4, bool rows) {
    if (nchw4 || rows)
        return false;
    // measured (profiles/r02): with the 16-warp epilogue the TMA store wins on every ResNet-50 / MobileNet
    // layer (bf16 step 3.47 ms vs 3.51 ms with direct stores everywhere), so the product never picks the
    // direct path; it stays reachable through the tools build (B200CB_DIRECT=1) and its parity run.
    (void)d;
    return false;
}

// Resident filter tile (IgemmParams::b_res) for 1x1 stride-1 layers run by single CTAs: the CTA's N tile of the
// filter (all k-blocks) is loaded once, the stages carry activations only.  What it buys is shared-memory
// residency: on the residual 1x1 layers a slot is occupied from its TMA issue to the end of its MMA, and halving
// the bytes a tile keeps in flight raises the tile rate the 227 KB can sustain.
bool pick_bres(const b200cb_desc& d, uint32_t block_n, uint32_t k_bytes, uint32_t kb_bytes, uint64_t m_tiles) {
    const uint32_t n_tiles = (d.oc + block_n - 1) / block_n;
    const uint32_t num_kc = (k_bytes + kb_bytes - 1) / kb_bytes;
    const uint32_t bres = num_kc * block_n * kb_bytes;
    (void)m_tiles;
    // the grid is cut to a multiple of n_tiles; the tile must leave room for the z ring and >= 2 stages
    if (n_tiles > 16 || bres > 64 * 1024)
        return false;
    // measured (tools/exp_bres.py, tools/bench_layers.py --set pw; profiles/r02/exp_resident_filter.txt):
    //  * low-K layers (K <= 64 bytes: MobileNetV2 / ShuffleNetV2 pointwise 16->96, 24->144, 32->192, 32->16) are
    //    bound by the number of short TMA rows per tile, of which the filter tile is 1/2 .. 2/3: bf16 0.238 -> 0.130,
    //    0.154 -> 0.060, 0.024 -> 0.019, 0.149 -> 0.096 ms (9-layer table 40 % -> 65 % of HBM; int8 13.5 % -> 24.6 %);
    //  * QuantizedS8 layers with longer K gain 0-10 % (l4.c3 0.029 -> 0.026) and never lose more than 1 %;
    //  * 16-bit layers with K >= 128 bytes are within +-2 % except l2.c3 (-8 %) and K = 512 (-3 %): streamed
    //  * the pixel-pair folded narrow layers (capi.cpp: 32 -> 16 runs as 64 -> 32, K = 128 bytes): 0.0580 -> 0.0557 ms
    return d.src_dtype == B200CB_DT_QS8 || k_bytes <= 64 || (k_bytes <= 128 && d.oc <= 32);
}

uint32_t pick_kb_bytes(uint32_t ic_bytes) {
    return ic_bytes > 64 ? 128u : (ic_bytes > 32 ? 64u : 32u);
}

uint32_t fixed_smem(uint32_t groups, uint32_t bn, bool has_z, uint32_t row_bytes, uint32_t nz = NZ_MIN) {
    return off_sd(groups, bn) + (has_z ? nz * 128 * row_bytes : slabs_bytes(groups, bn, row_bytes));
}

}  // namespace

RequantParams make_requant(const b200cb_desc& d) {
    RequantParams q;
    // absent bias: the accumulator's own scale s_src * s_flt (float product, dnn/src/common/utils.cpp:218-236)
    q.s_bias = d.bias_dtype == B200CB_DT_QS32 ? d.bias_scale : d.src_scale * d.filter_scale;
    q.s_z = d.has_z ? d.z_scale : 0.f;
    q.s_dst = d.dst_scale;
    q.one = 1.0f;
    const float y = 1.0f / d.dst_scale;
    q.inv_dst = y;
    q.fast_div = (isnormal(d.dst_scale) && isnormal(y) && d.dst_scale > 0.f) ? 1u : 0u;
    // fast_div == 2: every dividend the epilogue can produce (|acc + bias| <= 2^31, |z| <= 128, IDENTITY /
    // RELU do not amplify) has a finite quotient far from overflow, so the guard on |q0| can be dropped
    // and the ALU-only rounding of epilogue.cuh applies (b200cb_selftest(1) sweeps it).
    const double rmax = 2147483648.0 * fabs((double)q.s_bias) + 128.0 * fabs((double)q.s_z);
    if (q.fast_div && isfinite(q.s_bias) && isfinite(q.s_z) && rmax * (double)y < 1152921504606846976.0)
        q.fast_div = 2u;
    return q;
}

int igemm_selftest(int which, cudaStream_t stream, uint64_t* mismatches) {
    if (which < 0 || which > 2 || !mismatches)
        return B200CB_UNSUPPORTED_PARAM;
    const float h_scales[8] = {19.99f, 0.3f, 0.41f, 0.9f, 1.0f, 0.11f, 3.0e-3f, 1.7e4f};
    float* d_scales = nullptr;
    unsigned long long* d_bad = nullptr;
    cudaError_t e = cudaMalloc(&d_scales, sizeof(h_scales));
    if (e == cudaSuccess)
        e = cudaMalloc(&d_bad, sizeof(unsigned long long));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_scales, h_scales, sizeof(h_scales), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), stream);
    if (e == cudaSuccess) {
        if (which == 0)
            selftest_div_kernel<<<148 * 16, 256, 0, stream>>>(d_scales, 8, d_bad);
        else if (which == 1)
            selftest_fastq_kernel<<<148 * 16, 256, 0, stream>>>(d_scales, 8, d_bad);
        else
            selftest_packed_kernel<<<148 * 16, 256, 0, stream>>>(d_scales, 8, d_bad);
        e = cudaGetLastError();
    }
    unsigned long long bad = 0;
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
    cudaFree(d_scales);
    cudaFree(d_bad);
    if (e != cudaSuccess) {
        set_last_error("selftest", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    *mismatches = bad;
    return B200CB_OK;
}

// --------------------------------------------------------------------------- host side
// The dense convolution one kernel launch computes for descriptor `d` (already NHWC):
//  * GROUP (non-depthwise): one launch per group over channel-sliced views of src / dst / z -- what the
//    reference's AlgoGroupConvGeneral does with sub-operators (dnn/src/cuda/conv_bias/group_conv.cpp:1-190):
//    ic / oc become the per-group counts, *ld_src / *ld_dst keep the parent tensors' channel counts (the
//    pixel pitch of the views);
//  * BiasMode::BIAS (dst-shaped bias, dnn/src/common/conv_bias.cpp:38-49): bias + [no z] is the same
//    arithmetic as no bias + z = bias (both are added in fp32 before the activation), so the full bias
//    travels through the z ring (*bias_as_z).
// Returns B200CB_OK or the reason the cell is not supported.
int igemm_normalise(const b200cb_desc& d, b200cb_desc* out, uint32_t* ld_src, uint32_t* ld_dst, bool* bias_as_z) {
    *out = d;
    *ld_src = d.ic, *ld_dst = d.oc;
    *bias_as_z = false;
    if (d.group == 0 || d.ic % d.group || d.oc % d.group)
        return B200CB_UNSUPPORTED_SHAPE;
    if (d.group > 1) {
        if (d.sparse != B200CB_SPARSE_GROUP || d.format != B200CB_FORMAT_NHWC)
            return B200CB_UNSUPPORTED_PARAM;
        if (d.src_batch_stride || d.dst_batch_stride || d.z_batch_stride)
            return B200CB_UNSUPPORTED_SHAPE;  // grouped views of batch-strided tensors: not implemented
        out->ic = d.ic / d.group, out->oc = d.oc / d.group;
        out->group = 1, out->sparse = B200CB_SPARSE_DENSE;
    } else if (d.sparse != B200CB_SPARSE_DENSE) {
        return B200CB_UNSUPPORTED_PARAM;
    }
    if (d.bias_mode == B200CB_BIAS_FULL) {
        const bool fl = d.src_dtype == B200CB_DT_BF16 || d.src_dtype == B200CB_DT_F16;
        if (!fl || d.has_z || d.bias_dtype != d.dst_dtype || d.format != B200CB_FORMAT_NHWC)
            return B200CB_UNSUPPORTED_PARAM;  // (with z as well: two dst-shaped streams, not implemented)
        out->bias_mode = B200CB_BIAS_NONE, out->bias_dtype = B200CB_DT_NONE;
        out->has_z = 1, out->z_dtype = d.bias_dtype, out->z_batch_stride = 0;
        *bias_as_z = true;
    }
    return B200CB_OK;
}

int igemm_query(const b200cb_desc& d0, uint32_t block_n, b200cb_plan* plan) {
    const bool nchw4 = d0.format == B200CB_FORMAT_NCHW4;
    if (d0.format != B200CB_FORMAT_NHWC && !nchw4)
        return B200CB_UNSUPPORTED_FORMAT;
    if (nchw4) {
        // QuantizedS8 only (the reference's NCHW4 algos are int8: implicit_gemm_int8_nchw4_*.cpp)
        if (d0.src_dtype != B200CB_DT_QS8 || d0.dst_dtype != B200CB_DT_QS8)
            return B200CB_UNSUPPORTED_DTYPE;
        if (d0.ic % 4 != 0 || d0.oc % 4 != 0)
            return B200CB_UNSUPPORTED_ALIGNMENT;
        const uint64_t simg = (uint64_t)d0.ih * d0.iw * d0.ic;
        if (d0.src_batch_stride != 0 && (d0.src_batch_stride < simg || d0.src_batch_stride % 4 != 0))
            return B200CB_UNSUPPORTED_ALIGNMENT;
    }
    b200cb_desc d;
    uint32_t ld_src, ld_dst;
    bool bias_as_z;
    {
        const int rc = igemm_normalise(nchw4 ? nchw4_equivalent(d0) : d0, &d, &ld_src, &ld_dst, &bias_as_z);
        if (rc != B200CB_OK)
            return rc;
    }
    const bool grouped = d0.group > 1;
    int kind;
    if (d.src_dtype == B200CB_DT_BF16)
        kind = KIND_BF16;
    else if (d.src_dtype == B200CB_DT_F16)
        kind = KIND_F16;
    else if (d.src_dtype == B200CB_DT_QS8)
        kind = KIND_S8;
    else
        return B200CB_UNSUPPORTED_DTYPE;
    if (d.filter_dtype != d.src_dtype)
        return B200CB_UNSUPPORTED_DTYPE;
    bool out32 = false, out_f32 = false;
    if (d.dst_dtype == B200CB_DT_F32 && (kind == KIND_S8 || kind == KIND_BF16)) {
        // Float32 destination: QuantizedS8 -> Float32 (naive DISPATCH(QuantizedS8, Float32),
        // dnn/src/naive/conv_bias/opr_impl.cpp:248-275: Float32 bias and z) and the bf16 -> fp32 accumulate
        // kernel the Float32 ConvBias path runs on its split operands (capi.cpp)
        if ((d.bias_mode != B200CB_BIAS_NONE && d.bias_dtype != B200CB_DT_F32) || (d.has_z && d.z_dtype != B200CB_DT_F32))
            return B200CB_UNSUPPORTED_DTYPE;
        if (nchw4)
            return B200CB_UNSUPPORTED_DTYPE;
        out32 = out_f32 = true;
    } else if (kind == KIND_S8) {
        if (d.dst_dtype == B200CB_DT_QS32) {
            // raw int32 accumulator + bias; the oracle writes sfb straight into dst
            // (naive/conv_bias/opr_impl.cpp:223-228) only for IDENTITY without z
            if (d.has_z || d.nonline_mode != B200CB_NL_IDENTITY)
                return B200CB_UNSUPPORTED_PARAM;
            out32 = true;
        } else if (d.dst_dtype != B200CB_DT_QS8) {
            return B200CB_UNSUPPORTED_DTYPE;
        }
        if (d.bias_mode != B200CB_BIAS_NONE && d.bias_dtype != B200CB_DT_QS32)
            return B200CB_UNSUPPORTED_DTYPE;
        if (d.has_z && d.z_dtype != B200CB_DT_QS8)
            return B200CB_UNSUPPORTED_DTYPE;
        if (d.nonline_mode == B200CB_NL_SIGMOID)
            return B200CB_UNSUPPORTED_PARAM;  // reference forbids it too (common/conv_bias.cpp:313-316)
    } else {
        if (d.dst_dtype != d.src_dtype)
            return B200CB_UNSUPPORTED_DTYPE;
        if (d.bias_mode != B200CB_BIAS_NONE && d.bias_dtype != d.src_dtype)
            return B200CB_UNSUPPORTED_DTYPE;
        if (d.has_z && d.z_dtype != d.dst_dtype)
            return B200CB_UNSUPPORTED_DTYPE;
        // fp16 / bf16 kernels accumulate in fp32 whatever compute_mode says.  For Float16 with ComputeMode::DEFAULT
        // the naive operator accumulates in fp16 (naive/conv_bias/opr_impl.cpp:266); the reference's own CUDA
        // algorithms accumulate in fp32 there as well and its tests grant them 2e-2 against naive
        // (dnn/test/cuda/conv_bias.cpp:242-247,663-664) -- the same cell, the same tolerance
        // (tests/test_parity_gpu.py::test_fp16_default_compute_mode).
    }
    if (block_n != 64 && block_n != 128 && block_n != 256)
        return B200CB_UNSUPPORTED_ALGO;
    const uint32_t in_es = kind == KIND_S8 ? 1 : 2;
    const uint32_t out_es = out32 ? 4 : in_es;
    // TMA needs 16-byte aligned global strides
    // (8-byte pixels are legal when the W-fold turns a pixel pair into one 16-byte unit: checked below)
    if (((d.ic * in_es) % 16 != 0 && d.ic * in_es != 8) || (!nchw4 && (d.oc * out_es) % 16 != 0))
        return B200CB_UNSUPPORTED_ALIGNMENT;
    if (grouped && (d.ic * in_es) % 16 != 0)
        return B200CB_UNSUPPORTED_ALIGNMENT;  // every group's channel slice must start on a 16-byte boundary
    if (d.n == 0 || d.oh == 0 || d.ow == 0 || d.fh == 0 || d.fw == 0)
        return B200CB_UNSUPPORTED_SHAPE;
    if (d.stride_h == 0 || d.stride_w == 0 || d.stride_h > 8 || d.stride_w > 8)
        return B200CB_UNSUPPORTED_PARAM;
    if (d.dilate_h == 0 || d.dilate_w == 0)
        return B200CB_UNSUPPORTED_PARAM;
    // im2col TMA: corner offsets are 8-bit, tap offsets 16-bit
    if (d.pad_h > 127 || d.pad_w > 127)
        return B200CB_UNSUPPORTED_PARAM;
    if ((int)((d.fh - 1) * d.dilate_h) - (int)d.pad_h > 128 ||
        (int)((d.fw - 1) * d.dilate_w) - (int)d.pad_w > 128)
        return B200CB_UNSUPPORTED_PARAM;
    // dst / z must be dense over the batch (the output is one (OC, M) matrix)
    const uint64_t dst_img = (uint64_t)d.oh * d.ow * d.oc;
    if ((d.dst_batch_stride != 0 && d.dst_batch_stride != dst_img) ||
        (d.has_z && d.z_batch_stride != 0 && d.z_batch_stride != dst_img))
        return B200CB_UNSUPPORTED_SHAPE;
    const uint64_t src_img = (uint64_t)d.ih * d.iw * d.ic;
    if (d.src_batch_stride != 0 && (d.src_batch_stride < src_img || (d.src_batch_stride * in_es) % 16 != 0))
        return B200CB_UNSUPPORTED_ALIGNMENT;
    const uint64_t M = (uint64_t)d.n * d.oh * d.ow;
    if (M > 0x7fffffffull)
        return B200CB_UNSUPPORTED_SHAPE;
    // smallest sensible tile: do not offer a tile more than 2x wider than OC
    if (block_n > 64 && block_n / 2 >= d.oc)
        return B200CB_UNSUPPORTED_SHAPE;

    b200cb_desc dd = d;
    if (d.src_batch_stride == src_img)
        dd.src_batch_stride = 0;
    // NCHW4 runs the generic modes only (no W-fold / stem / rows: their filter packing and output
    // paths are NHWC-specific); its z is read in place by the epilogue, not through the TMA ring
    const Geometry g = make_geometry(dd, in_es, !grouped && !out_f32 && (!nchw4 || dd.ic == 8));
    if ((g.ic * in_es) % 16 != 0)
        return B200CB_UNSUPPORTED_ALIGNMENT;
    if (g.folded && (g.pad_w_hi - (int32_t)(g.fw - 1) < -128 || g.pad_w_lo > 127))
        return B200CB_UNSUPPORTED_PARAM;
    const bool stem = g.stem && block_n == 64;
    const uint32_t kb_bytes = pick_kb_bytes(g.ic * in_es);
    const uint32_t fixed0 =
            fixed_smem(out32 ? 1 : 4, block_n, d.has_z && !nchw4, (out32 || in_es == 2) ? 128 : 64);
    // rows mode: the filter and >= 2 input slabs must fit next to the epilogue buffers
    const RowsGeom rg = rows_geometry(dd, in_es, block_n);
    const bool rows = !nchw4 && !grouped && !out32 && !stem && !g.folded && rg.ok &&
                      fixed0 + rg.bres + 2 * rg.a_stage <= MAX_DYN_SMEM;
    // generic modes: CTA pair (cta_group::2, 256 x BN tile, filter rows split between the two CTAs) or
    // a thread-block cluster sharing each filter tile by TMA multicast
    uint32_t cluster = 1, pair = 0;
    if (!stem && !rows) {
        const uint32_t m_tiles = (uint32_t)((M + IGEMM_BLOCK_M - 1) / IGEMM_BLOCK_M);
        pair = pick_pair(d, block_n, m_tiles);
#ifdef B200CB_TOOLS  // experiment knobs of the tools build (make tools); the product query is pure
        if (const char* e = getenv("B200CB_PAIR"))
            pair = atoi(e) != 0;
#endif
        if (m_tiles < 4)
            pair = 0;
        cluster = pair ? 2 : 1;
#ifdef B200CB_TOOLS
        if (const char* e = getenv("B200CB_CLUSTER"))
            if (!pair)
                cluster = (uint32_t)atoi(e);
#endif
        if (cluster != 1 && cluster != 2 && cluster != 4)
            cluster = 1;
        while (!pair && cluster > 1 && (m_tiles < 2 * cluster || (block_n / cluster) % 8 != 0))
            cluster /= 2;
    }
    const bool pointwise = d.fh == 1 && d.fw == 1 && d.stride_h == 1 && d.stride_w == 1 &&
                           d.pad_h == 0 && d.pad_w == 0 && dd.src_batch_stride == 0;
    // resident filter tile: plain 1x1 layers on single CTAs (a_mode 0)
    bool b_res = false;
    const uint32_t bres_tile = ((g.ic * in_es + kb_bytes - 1) / kb_bytes) * block_n * kb_bytes;
    if (pointwise && !stem && !rows && !pair && cluster == 1 && !nchw4 && !grouped) {
        b_res = pick_bres(d, block_n, g.ic * in_es, kb_bytes, (M + IGEMM_BLOCK_M - 1) / IGEMM_BLOCK_M);
#ifdef B200CB_TOOLS
        if (const char* e = getenv("B200CB_BRES"))
            b_res = atoi(e) != 0 && (d.oc + block_n - 1) / block_n <= 16 && bres_tile <= 96 * 1024;
#endif
    }
    const uint32_t stem_rb = g.stem4 ? 64u : 128u;
    const uint32_t stage_b = stem    ? (g.stem4 ? 1 : 2) * d.fh * STEM_NJ * 16
                             : rows  ? rg.a_stage
                             : b_res ? IGEMM_BLOCK_M * kb_bytes
                                     : (IGEMM_BLOCK_M + (pair ? block_n / 2 : block_n)) * kb_bytes;
    const uint32_t fixed = fixed0 + (stem ? d.fh * block_n * stem_rb : rows ? rg.bres : b_res ? bres_tile : 0);
    uint32_t stages = (MAX_DYN_SMEM - fixed) / stage_b;
    // stem row ring: 16 row pairs + slack for the descriptor rows past OW (read, never stored)
    const bool ring = stem && g.ring;
    const uint32_t ring_bytes = ring ? STEM_RING_PAIRS * 2 * g.ring_nj * 16 + 4096 : 0;
    if (rows && stages > 3)
        stages = 3;
    if (stages > IGEMM_MAX_STAGES)
        stages = IGEMM_MAX_STAGES;
    if (stages < 2)
        return B200CB_UNSUPPORTED_SHAPE;
    // z ring depth: layers with a short k-loop are bound by the z / dst streams; give them z chunks
    // (prefetched a tile or more ahead) instead of a third A/B stage
    uint32_t nz = NZ_MIN, smem_total = fixed + stages * stage_b;
    if (ring) {
        if (fixed + ring_bytes > MAX_DYN_SMEM)
            return B200CB_UNSUPPORTED_SHAPE;
        stages = STEM_RING_AHEAD;
        smem_total = fixed + ring_bytes;
    }
    if (d.has_z && !nchw4 && !stem && !rows) {
        const uint32_t zbuf = 128 * ((out32 || in_es == 2) ? 128 : 64);
        const uint32_t num_kb = d.fh * d.fw * ((g.ic * in_es + kb_bytes - 1) / kb_bytes);
        const uint32_t keep = num_kb >= 4 ? 3u : 2u;  // stages the main loop keeps (measured: l3.c3 wants 3)
        // the ring is GROUPS sub-rings (one per epilogue group): 4 chunks (depth 1) or 8 (depth 2)
        const uint32_t extra = NZ_MAX - NZ_MIN;
        if (MAX_DYN_SMEM >= fixed + extra * zbuf + keep * stage_b) {
            nz = NZ_MAX;
            stages = (MAX_DYN_SMEM - fixed - extra * zbuf) / stage_b;
            if (stages > IGEMM_MAX_STAGES)
                stages = IGEMM_MAX_STAGES;
            smem_total = fixed + extra * zbuf + stages * stage_b;
        }
    }

#ifdef B200CB_TOOLS
    if (const char* s2 = getenv("B200CB_STAGES"))  // pipeline-depth sensitivity (generic modes)
        if (!getenv("B200CB_NZ") && !stem && !rows && (uint32_t)atoi(s2) >= 2 && (uint32_t)atoi(s2) < stages) {
            smem_total -= (stages - (uint32_t)atoi(s2)) * stage_b;
            stages = (uint32_t)atoi(s2);
        }
    if (const char* e = getenv("B200CB_NZ")) {  // force the z ring depth (stages shrink to fit)
        if (d.has_z && !nchw4 && !stem && !rows) {
            const uint32_t zbuf = 128 * ((out32 || in_es == 2) ? 128 : 64);
            const uint32_t want = (uint32_t)atoi(e);
            if (want == 4 || want == 8) {
                const uint32_t fixed_nz = fixed - NZ_MIN * zbuf + want * zbuf;
                stages = (MAX_DYN_SMEM - fixed_nz) / stage_b;
                if (stages > IGEMM_MAX_STAGES)
                    stages = IGEMM_MAX_STAGES;
                if (const char* s2 = getenv("B200CB_STAGES"))
                    if ((uint32_t)atoi(s2) >= 1 && (uint32_t)atoi(s2) < stages)
                        stages = (uint32_t)atoi(s2);
                nz = want;
                smem_total = fixed_nz + stages * stage_b;
            }
        }
    }
#endif
    plan->magic = PLAN_MAGIC;
    plan->kernel = KERNEL_IGEMM;
    plan->block_n = block_n;
    plan->stages = stages;
    plan->grid = 0;  // resolved at launch from the device's SM count
    plan->smem_bytes = smem_total;
    plan->a_mode = stem ? 2 : rows ? 3 : (pointwise ? 0 : 1);
    for (auto& r : plan->reserved)
        r = 0;
    plan->reserved[0] = static_cast<uint32_t>(kind);
    plan->reserved[1] = (out32 ? 1u : 0u) | (out_f32 ? 2u : 0u);
    // bit 0: debugging override (force the im2col path on pointwise convs); bit 1: direct-store epilogue
    plan->reserved[2] = pick_direct_store(d, nchw4, rows) ? 2u : 0u;
#ifdef B200CB_TOOLS
    if (const char* e = getenv("B200CB_DIRECT"))
        plan->reserved[2] = (atoi(e) != 0 && !nchw4 && !rows) ? 2u : 0u;
#endif
#ifdef B200CB_TOOLS
    if (const char* e = getenv("B200CB_ZJIT"))
        if (atoi(e) != 0 && d.has_z && !nchw4 && !rows)
            plan->reserved[2] |= 4u;
#endif
    if (b_res)
        plan->reserved[2] |= 8u;  // bit 3: resident filter tile (a_mode 0)
    plan->reserved[3] = 0;  // 1: `filter` passed to launch is already packed (weight preprocess)
    plan->reserved[4] = (g.folded ? 1u : 0u) | (ring ? 2u : 0u);  // bit 0: W-folded (packed filter), bit 1: row ring
    plan->reserved[5] = kb_bytes | (nz << 16);
    plan->reserved[6] = cluster;
    plan->reserved[7] = pair;
    // folded stem: the re-packed filter lives in the workspace unless it was preprocessed
    const uint32_t fwp = stem ? g.fw_packed : g.fw;
    plan->workspace_bytes =
            g.folded ? (((uint64_t)d.oc * d.fh * fwp * g.ic * in_es + 255) & ~255ull) : 0;
    if (nchw4)  // NHWC copy of src + KRSC copy of the filter (channels padded to d.ic) [+ packed stem filter]
        plan->workspace_bytes = align256((size_t)d.n * d.ih * d.iw * d.ic) +
                                align256((size_t)d.oc * d.fh * d.fw * d.ic) + plan->workspace_bytes;
    plan->desc = d0;
    return B200CB_OK;
}

size_t igemm_packed_filter_bytes(const b200cb_plan& plan) {
    // (NCHW4 filters are re-tiled and packed on every exec: no weight-preprocess form)
    if (!plan.reserved[4] || plan.desc.format == B200CB_FORMAT_NCHW4)
        return 0;
    const b200cb_desc& d = plan.desc;
    const uint32_t in_es = plan.reserved[0] == KIND_S8 ? 1 : 2;
    const Geometry g = make_geometry(d, in_es);
    const uint32_t fwp = plan.a_mode == 2 ? g.fw_packed : g.fw;
    return (size_t)d.oc * d.fh * fwp * g.ic * in_es;
}

int igemm_pack_filter(const b200cb_plan& plan, const void* filter, void* packed, cudaStream_t stream) {
    if (!plan.reserved[4])
        return B200CB_OK;
    // `filter` is KRSC (for NCHW4 plans: the re-tiled workspace copy)
    const b200cb_desc d =
            plan.desc.format == B200CB_FORMAT_NCHW4 ? nchw4_equivalent(plan.desc) : plan.desc;
    const uint32_t in_es = plan.reserved[0] == KIND_S8 ? 1 : 2;
    const Geometry g = make_geometry(d, in_es);
    const uint32_t fwp = plan.a_mode == 2 ? g.fw_packed : g.fw;
    const uint32_t total = d.oc * d.fh * fwp * g.ic;
    const uint32_t blocks = (total + 255) / 256;
    if (in_es == 1)
        fold_filter_kernel<int8_t><<<blocks, 256, 0, stream>>>(
                static_cast<const int8_t*>(filter), static_cast<int8_t*>(packed), d.oc, d.fh, d.fw,
                d.ic, fwp, g.fold_shift);
    else
        fold_filter_kernel<uint16_t><<<blocks, 256, 0, stream>>>(
                static_cast<const uint16_t*>(filter), static_cast<uint16_t*>(packed), d.oc, d.fh,
                d.fw, d.ic, fwp, g.fold_shift);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("fold_filter launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

namespace {
int launch_dense(
        const b200cb_plan& plan, const b200cb_desc& d, uint32_t ld_src, uint32_t ld_dst, const void* src,
        const void* filter, const void* bias, const void* z, void* dst, void* workspace, cudaStream_t stream);
}

int igemm_launch(
        const b200cb_plan& plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* workspace, size_t /*workspace_bytes*/,
        cudaStream_t stream) {
    const bool nchw4 = plan.desc.format == B200CB_FORMAT_NCHW4;
    b200cb_desc d;
    uint32_t ld_src, ld_dst;
    bool bias_as_z;
    if (igemm_normalise(nchw4 ? nchw4_equivalent(plan.desc) : plan.desc, &d, &ld_src, &ld_dst, &bias_as_z) != B200CB_OK)
        return B200CB_ERR_BAD_PLAN;
    if (!src || !filter || !dst || (plan.desc.bias_mode != B200CB_BIAS_NONE && !bias) || (plan.desc.has_z && !z))
        return B200CB_ERR_NULL;
    if (bias_as_z) {
        z = bias;
        bias = nullptr;
    }
    const uint32_t groups = plan.desc.group;
    if (groups <= 1)
        return launch_dense(plan, d, ld_src, ld_dst, src, filter, bias, z, dst, workspace, stream);
    // one dense launch per group on channel-sliced views (pixel pitch = the parent's channel count)
    const uint32_t es = dtype_size(d.src_dtype), oes = dtype_size(d.dst_dtype);
    const size_t flt_g = (size_t)d.oc * d.fh * d.fw * d.ic * es;
    for (uint32_t g = 0; g < groups; ++g) {
        const int rc = launch_dense(
                plan, d, ld_src, ld_dst, static_cast<const uint8_t*>(src) + (size_t)g * d.ic * es,
                static_cast<const uint8_t*>(filter) + g * flt_g,
                bias ? static_cast<const uint8_t*>(bias) + (size_t)g * d.oc * dtype_size(d.bias_dtype) : nullptr,
                z ? static_cast<const uint8_t*>(z) + (size_t)g * d.oc * dtype_size(d.z_dtype) : nullptr,
                static_cast<uint8_t*>(dst) + (size_t)g * d.oc * oes, workspace, stream);
        if (rc != B200CB_OK)
            return rc;
    }
    return B200CB_OK;
}

namespace {
int launch_dense(
        const b200cb_plan& plan, const b200cb_desc& d, uint32_t ld_src, uint32_t ld_dst, const void* src,
        const void* filter, const void* bias, const void* z, void* dst, void* workspace, cudaStream_t stream) {
    const bool nchw4 = plan.desc.format == B200CB_FORMAT_NCHW4;
    const bool grouped = plan.desc.group > 1;
    const int kind = static_cast<int>(plan.reserved[0]);
    const bool out32 = (plan.reserved[1] & 1) != 0;
    const bool out_f32 = (plan.reserved[1] & 2) != 0;
    const uint32_t in_es = kind == KIND_S8 ? 1 : 2;
    const uint32_t out_es = out32 ? 4 : in_es;
    const uint32_t epi_n = out32 ? 32 : 64;
    const uint32_t row_bytes = epi_n * out_es;
    const uint32_t BN = plan.block_n;
    uint32_t a_mode = plan.a_mode;
    if (plan.reserved[2] & 1)
        a_mode = 1;  // debugging override: force the im2col path on pointwise convs

    if (!src || !filter || !dst || (d.bias_mode != B200CB_BIAS_NONE && !bias) || (d.has_z && !z))
        return B200CB_ERR_NULL;
    if (nchw4) {
        // pre-pass: src (N, C/4, H, W, 4) -> NHWC, filter (OC, IC/4, FH, FW, 4) -> KRSC, in the workspace
        const b200cb_desc& o = plan.desc;
        if (!workspace)
            return B200CB_ERR_WORKSPACE;
        uint8_t* ws_src = static_cast<uint8_t*>(workspace);
        uint8_t* ws_flt = ws_src + align256((size_t)d.n * d.ih * d.iw * d.ic);
        int rc = relayout_c4_to_channels_last(
                src, ws_src, o.n, o.ic, d.ic, o.ih * o.iw, o.src_batch_stride, stream);
        if (rc == B200CB_OK)
            rc = relayout_c4_to_channels_last(filter, ws_flt, o.oc, o.ic, d.ic, o.fh * o.fw, 0, stream);
        if (rc != B200CB_OK)
            return rc;
        src = ws_src;
        filter = ws_flt;
    }

    b200cb_desc dd = d;
    const uint64_t src_img = (uint64_t)d.ih * d.iw * ld_src;  // elements between images of a dense parent tensor
    if (d.src_batch_stride == src_img)
        dd.src_batch_stride = 0;
    const Geometry g = make_geometry(dd, in_es, !grouped && !out_f32 && (!nchw4 || dd.ic == 8));
    const uint32_t kb_bytes = pick_kb_bytes(g.ic * in_es);
    const uint32_t kb_elems = kb_bytes / in_es;
    const CUtensorMapSwizzle in_swz = kb_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : kb_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                       : CU_TENSOR_MAP_SWIZZLE_32B;
    const void* flt = filter;
    if (g.folded && !plan.reserved[3]) {
        if (!workspace)
            return B200CB_ERR_WORKSPACE;
        // packed stem filter: the workspace, after the NCHW4 copies if any
        void* ws_pack = workspace;
        if (nchw4)
            ws_pack = static_cast<uint8_t*>(workspace) + align256((size_t)d.n * d.ih * d.iw * d.ic) +
                      align256((size_t)d.oc * d.fh * d.fw * d.ic);
        int rc = igemm_pack_filter(plan, filter, ws_pack, stream);
        if (rc != B200CB_OK)
            return rc;
        flt = ws_pack;
    }

    // Kernel parameters (5 encoded tensor maps + scalars) are a pure function of (device, plan, tensor
    // addresses): a steady-state network launches the same layer on the same buffers every step, so they
    // are cached (a mutex-guarded CUtensorMap cache keyed by (device, plan)).
    IgemmParams p;
    ParamKey key;
    memset(&key, 0, sizeof(key));
    key.plan = plan;
    key.src = src, key.flt = flt, key.bias = bias, key.z = z, key.dst = dst;
    cudaGetDevice(&key.dev);
    auto build_params = [&]() -> int {
    memset(&p, 0, sizeof(p));
    const uint64_t M = (uint64_t)d.n * d.oh * d.ow;
    const CUtensorMapDataType in_dt =
            kind == KIND_S8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                            : (kind == KIND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    const CUtensorMapDataType out_dt = out32 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : in_dt;
    const uint64_t src_bstride = d.src_batch_stride ? d.src_batch_stride : src_img;
    CUresult cr;
    const bool stem = a_mode == 2;
    const bool rows = a_mode == 3;
    const RowsGeom rg = rows ? rows_geometry(dd, in_es, BN) : RowsGeom{};
    if (rows) {
        // whole padded input rows: (C, W, H, N) boxes of PW x RH pixels, zero fill does the padding
        uint64_t dims[4] = {d.ic, d.iw, d.ih, d.n};
        uint64_t strides[3] = {
                (uint64_t)d.ic * in_es, (uint64_t)d.iw * d.ic * in_es, src_bstride * in_es};
        uint32_t box[4] = {kb_elems, rg.pw, rg.rh, 1};
        cr = encode_tiled(
                &p.tmA, in_dt, 4, src, dims, strides, box, nullptr, in_swz,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    } else if (stem && (plan.reserved[4] & 2)) {
        // row ring: rows as contiguous runs of 8-byte elements (a pixel of IC * es == 8 bytes), boxes of one
        // row pair; left / right / top / bottom padding by out-of-bounds zero fill
        uint64_t dims[3] = {(uint64_t)d.iw * d.ic * in_es / 8, d.ih, d.n};
        uint64_t strides[2] = {(uint64_t)d.iw * d.ic * in_es, src_bstride * in_es};
        uint32_t box[3] = {2 * g.ring_nj, 2, 1};
        cr = encode_tiled(
                &p.tmA, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, src, dims, strides, box, nullptr,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    } else if (stem) {
        // even / odd input pixels as two (C, W/2, H, N) tensors whose W stride is one pixel pair
        // (stem4: one tensor whose innermost 16 bytes are the pixel pair itself)
        const uint32_t unit = g.stem4 ? 2 * d.ic : d.ic;
        uint64_t dims[4] = {unit, d.iw / 2, d.ih, d.n};
        uint64_t strides[3] = {
                2ull * d.ic * in_es, (uint64_t)d.iw * d.ic * in_es, src_bstride * in_es};
        uint32_t box[4] = {unit, STEM_NJ, d.fh, 1};
        cr = encode_tiled(
                &p.tmA, in_dt, 4, src, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        if (cr == CUDA_SUCCESS && !g.stem4)
            cr = encode_tiled(
                    &p.tmA2, in_dt, 4, static_cast<const uint8_t*>(src) + d.ic * in_es, dims, strides,
                    box, nullptr, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    } else if (a_mode == 0) {
        uint64_t dims[2] = {d.ic, M};
        uint64_t strides[1] = {(uint64_t)ld_src * in_es};
        uint32_t box[2] = {kb_elems, IGEMM_BLOCK_M};
        cr = encode_tiled(
                &p.tmA, in_dt, 2, src, dims, strides, box, nullptr, in_swz,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    } else {
        // (grouped views are never W-folded: g.ic == d.ic there and the pixel pitch is the parent's)
        const uint64_t pix = g.folded ? (uint64_t)g.ic : (uint64_t)ld_src;
        uint64_t dims[4] = {g.ic, g.iw, d.ih, d.n};
        uint64_t strides[3] = {pix * in_es, (uint64_t)g.iw * pix * in_es, src_bstride * in_es};
        int lower[2] = {-g.pad_w_lo, -(int)d.pad_h};
        int upper[2] = {
                g.pad_w_hi - (int)((g.fw - 1) * d.dilate_w),
                (int)d.pad_h - (int)((d.fh - 1) * d.dilate_h)};
        uint32_t estr[4] = {1, g.stride_w, d.stride_h, 1};
        cr = encode_im2col(
                &p.tmA, in_dt, 4, src, dims, strides, lower, upper, kb_elems, IGEMM_BLOCK_M, estr,
                in_swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, src_bstride * in_es * d.n);
    }
    if (cr != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof(buf), "CUresult %d (activation map, a_mode %u)", (int)cr, a_mode);
        set_last_error("cuTensorMapEncode", buf);
        return B200CB_ERR_DRIVER;
    }
    {
        // packed stem filter: (4 tap pairs x 2*IC = 128 bytes, FH, OC), one 128-byte K row per filter row
        uint64_t dims_stem[3] = {(uint64_t)g.fw_packed * g.ic, d.fh, d.oc};
        const uint64_t rb = (uint64_t)g.fw_packed * g.ic * in_es;  // 128, or 64 for stem4
        uint64_t strides_stem[2] = {rb, (uint64_t)d.fh * rb};
        uint32_t box_stem[3] = {g.fw_packed * g.ic, 1, BN};
        uint64_t dims[3] = {g.ic, (uint64_t)d.fh * g.fw, d.oc};
        uint64_t strides[2] = {(uint64_t)g.ic * in_es, (uint64_t)d.fh * g.fw * g.ic * in_es};
        const uint32_t C = (stem || rows) ? 1u : (plan.reserved[6] ? plan.reserved[6] : 1u);
        uint32_t box[3] = {kb_elems, 1, BN / C};  // clusters: each CTA loads (and multicasts) BN / C rows
        cr = stem ? encode_tiled(
                            &p.tmB, in_dt, 3, flt, dims_stem, strides_stem, box_stem, nullptr,
                            rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B)
                  : encode_tiled(
                &p.tmB, in_dt, 3, flt, dims, strides, box, nullptr, in_swz,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (cr != CUDA_SUCCESS) {
            char buf[64];
            snprintf(buf, sizeof(buf), "CUresult %d (filter map)", (int)cr);
            set_last_error("cuTensorMapEncodeTiled", buf);
            return B200CB_ERR_DRIVER;
        }
    }
    if (!nchw4) {
        const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        // stem tiles are row segments: dst / z as (OC, OW, N*OH) so that a tile's tail is clipped
        uint64_t dims[3] = {d.oc, (stem || rows) ? (uint64_t)d.ow : M, (uint64_t)d.n * d.oh};
        uint64_t strides[2] = {(uint64_t)ld_dst * out_es, (uint64_t)d.ow * ld_dst * out_es};
        uint32_t box_d[3] = {epi_n, 32, 1};
        uint32_t box_z[3] = {epi_n, IGEMM_BLOCK_M, 1};
        const uint32_t rank = (stem || rows) ? 3 : 2;
        cr = encode_tiled(
                &p.tmD, out_dt, rank, dst, dims, strides, box_d, nullptr, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        if (cr == CUDA_SUCCESS && d.has_z)
            cr = encode_tiled(
                    &p.tmZ, out_dt, rank, z, dims, strides, box_z, nullptr, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        if (cr != CUDA_SUCCESS) {
            char buf[64];
            snprintf(buf, sizeof(buf), "CUresult %d (dst/z map)", (int)cr);
            set_last_error("cuTensorMapEncodeTiled", buf);
            return B200CB_ERR_DRIVER;
        }
    }
    p.bias = bias;
    p.dst = dst;
    p.M = (uint32_t)M;
    p.OC = d.oc;
    p.OHW = d.oh * d.ow;
    p.OW = d.ow;
    p.FW = g.fw;
    p.num_taps = d.fh * g.fw;
    p.stride_h = d.stride_h;
    p.stride_w = g.stride_w;
    p.pad_h = d.pad_h;
    p.pad_w = g.pad_w_lo;
    p.dil_h = d.dilate_h;
    p.dil_w = d.dilate_w;
    p.kb_bytes = kb_bytes;
    p.kb_elems = kb_elems;
    p.num_kc = (g.ic + kb_elems - 1) / kb_elems;
    p.m_tiles = (uint32_t)((M + IGEMM_BLOCK_M - 1) / IGEMM_BLOCK_M);
    p.n_tiles = (d.oc + BN - 1) / BN;
    p.OH = d.oh;
    p.nseg = (d.ow + IGEMM_BLOCK_M - 1) / IGEMM_BLOCK_M;
    p.stem_nj = STEM_NJ;
    p.stem_fh = d.fh;
    p.stem_fw2 = g.fw;
    p.stem_pad_lo2 = g.pad_w_lo;
    p.stem4 = (stem && g.stem4) ? 1 : 0;
    if (stem && (plan.reserved[4] & 2)) {
        p.stem_ring = 1;
        p.stem_nj = g.ring_nj;
        p.stem_rowb = g.ring_nj * 16;
        p.stem_np = g.ring_np;
    }
    if (stem)
        p.m_tiles = d.n * d.oh * p.nseg;
    if (rows) {
        p.PW = rg.pw;
        p.rows_rh = rg.rh;
        p.tiles_per_img = rg.tiles_per_img;
        p.m_tiles = d.n * rg.tiles_per_img;
    }
    p.stages = plan.stages;
    p.a_mode = a_mode;
    p.cluster = (stem || rows) ? 1u : (plan.reserved[6] ? plan.reserved[6] : 1u);
    p.pair = (stem || rows) ? 0u : plan.reserved[7];
    p.flip = d.mode == B200CB_MODE_CONVOLUTION;
    p.has_bias = d.bias_mode != B200CB_BIAS_NONE;
    p.has_z = d.has_z && !nchw4;
    p.nchw4 = nchw4;
    p.z_direct = d.has_z && nchw4;
    p.direct_store = ((plan.reserved[2] >> 1) & 1) && ld_dst == d.oc;  // (the st.global path assumes a dense dst)
    p.out_f32 = out_f32 ? 1 : 0;
    p.acc_scale = d.src_scale * d.filter_scale;
    {
        const uint64_t kb = (uint64_t)d.fh * d.fw * d.ic * 16384ull;  // |int8 * int8| <= 2^14 per tap and channel
        p.acc_bound = kb > 0xffffffffull ? 0xffffffffu : (uint32_t)kb;
    }
    p.z = z;
    p.nz = (plan.reserved[5] >> 16) ? (plan.reserved[5] >> 16) : NZ_MIN;
    p.z_jit = (plan.reserved[2] >> 2) & 1u;
    // (set by the query for 1x1 layers only: one tap; the forced-im2col debugging path keeps it)
    p.b_res = (!stem && !rows && !p.pair && p.cluster == 1 && p.num_taps == 1) ? (plan.reserved[2] >> 3) & 1u : 0u;
    p.nonline = d.nonline_mode;
    p.rq = make_requant(d);
#ifdef B200CB_DEBUG_TIMING
    if (const char* e = getenv("B200CB_DBG_PTR"))
        p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 16));
#endif

    return B200CB_OK;
    };
#ifdef B200CB_DEBUG_TIMING
    const bool cached = false;
#else
    const bool cached = param_cache().lookup(key, &p);
#endif
    if (!cached) {
        const int rc = build_params();
        if (rc != B200CB_OK)
            return rc;
        param_cache().store(key, p);
    }

    // the cap on the grid counts M tiles only per N tile: cluster tiles = ceil(m_tiles / C) * n_tiles
    const uint64_t total = (uint64_t)((p.m_tiles + p.cluster - 1) / p.cluster) * p.cluster * p.n_tiles;
    const uint32_t grid = plan.grid / p.cluster * p.cluster;

    cudaError_t e = cudaErrorInvalidValue;
#define B200CB_DISPATCH(KIND_, OUT_, PAIR_)                                                        \
    switch (BN) {                                                                             \
        case 64: e = launch_variant<KIND_, 64, OUT_, PAIR_>(p, grid, total, plan.smem_bytes, stream); break;   \
        case 128: e = launch_variant<KIND_, 128, OUT_, PAIR_>(p, grid, total, plan.smem_bytes, stream); break; \
        case 256: e = launch_variant<KIND_, 256, OUT_, PAIR_>(p, grid, total, plan.smem_bytes, stream); break; \
        default: break;                                                                       \
    }
    if (p.pair) {
        if (kind == KIND_BF16 && out32) {
            B200CB_DISPATCH(KIND_BF16, 1, 1)
        } else if (kind == KIND_BF16) {
            B200CB_DISPATCH(KIND_BF16, 0, 1)
        } else if (kind == KIND_F16) {
            B200CB_DISPATCH(KIND_F16, 0, 1)
        } else if (!out32) {
            B200CB_DISPATCH(KIND_S8, 0, 1)
        } else {
            B200CB_DISPATCH(KIND_S8, 1, 1)
        }
    } else if (kind == KIND_BF16 && out32) {
        B200CB_DISPATCH(KIND_BF16, 1, 0)
    } else if (kind == KIND_BF16) {
        B200CB_DISPATCH(KIND_BF16, 0, 0)
    } else if (kind == KIND_F16) {
        B200CB_DISPATCH(KIND_F16, 0, 0)
    } else if (!out32) {
        B200CB_DISPATCH(KIND_S8, 0, 0)
    } else {
        B200CB_DISPATCH(KIND_S8, 1, 0)
    }
#undef B200CB_DISPATCH
    if (e != cudaSuccess) {
        set_last_error("igemm launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}
}  // namespace

}  // namespace b200cb

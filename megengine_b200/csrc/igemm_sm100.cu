// Dense NHWC ConvBias as an implicit GEMM on Blackwell tensor cores (sm_100a).
//
//   GEMM view: M = N*OH*OW output pixels, N = OC, K = FH*FW*IC.
//   A (activations) tiles: TMA im2col mode straight from the NHWC tensor -- 128 pixels x 128 bytes
//     of channels per (filter tap, channel block); padding / stride / image wrap-around are done
//     by the TMA unit (zero fill), so no im2col buffer ever exists in HBM.
//   B (filter, KRSC) tiles: tiled TMA, BLOCK_N output channels x 128 bytes.
//   MMA: tcgen05.mma cta_group::1, M=128, N=BLOCK_N, K=32 bytes per instruction, fp32 (bf16/fp16)
//     or int32 (int8) accumulators in TMEM, double buffered so the epilogue of tile i overlaps
//     the main loop of tile i+1.
//   Epilogue: tcgen05.ld -> registers -> (+bias, +z, activation, requantize) -> swizzled smem
//     -> TMA store. z tiles are prefetched by TMA into smem by a dedicated warp.
//
// Replaces (does not call) the reference algos listed in SURVEY.md 8(a13):
//   dnn/src/cuda/conv_bias/implicit_gemm_int8_nhwc_imma.cpp, cudnn_conv_bias_activation.cpp,
//   bfloat16.cpp.  Epilogue arithmetic follows the naive oracle
//   dnn/src/naive/conv_bias/opr_impl.cpp:21-98 (see oracle/convbias_oracle.c).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include "common.h"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace b200cb {

struct IgemmParams {
    CUtensorMap tmA;  // activations: im2col (C,W,H,N) or tiled (C, M)
    CUtensorMap tmB;  // filter (IC, FH*FW, OC)
    CUtensorMap tmD;  // dst (OC, M)
    CUtensorMap tmZ;  // z   (OC, M)
    const void* bias;
    uint32_t M, OC, IC;
    uint32_t OH, OW;
    uint32_t FH, FW;
    int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    uint32_t num_kc;  // channel blocks per tap
    uint32_t m_tiles, n_tiles;
    uint32_t stages;
    uint32_t a_mode;  // 0 tiled, 1 im2col
    uint32_t flip;    // mode == CONVOLUTION
    uint32_t has_bias, has_z, nonline;
    float s_bias, s_z, s_dst;
};

namespace {

using namespace ptx;

template <int KIND, int OUT32>
struct Traits {
    static constexpr int IN_ES = (KIND == KIND_S8) ? 1 : 2;
    static constexpr int OUT_ES = OUT32 ? 4 : IN_ES;
    static constexpr int KB_ELEMS = IGEMM_KBYTES / IN_ES;
    static constexpr int EPI_N = OUT32 ? 32 : 64;       // columns per epilogue chunk
    static constexpr int ROW_BYTES = EPI_N * OUT_ES;    // 128 (16-bit, 32-bit) or 64 (int8)
};

template <int KIND, int BN, int OUT32>
__global__ void __launch_bounds__(256, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
    using T = Traits<KIND, OUT32>;
    constexpr uint32_t A_BYTES = IGEMM_BLOCK_M * IGEMM_KBYTES;  // 16 KiB
    constexpr uint32_t B_BYTES = BN * IGEMM_KBYTES;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;  // 128 / 256 / 512: powers of two >= 32
    constexpr uint32_t NUM_CHUNKS = BN / T::EPI_N;
    constexpr uint32_t UMMA_LAYOUT_SW128 = 2;
    constexpr uint32_t IDESC =
            (KIND == KIND_S8) ? umma_idesc(2, 1, 128, BN)
                              : umma_idesc(1, KIND == KIND_BF16 ? 1 : 0, 128, BN);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    // header layout
    const uint32_t bar_full = smem_base;               // [MAX_STAGES] x 8
    const uint32_t bar_empty = smem_base + 64;         // [MAX_STAGES] x 8
    const uint32_t bar_tfull = smem_base + 128;        // [2]
    const uint32_t bar_tempty = smem_base + 144;       // [2]
    const uint32_t bar_zfull = smem_base + 160;        // [2]
    const uint32_t bar_zempty = smem_base + 176;       // [2]
    const uint32_t tmem_slot = smem_base + 192;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + 192);
    uint32_t* sbias = reinterpret_cast<uint32_t*>(smem_gen + 1024);  // BN x 4 bytes

    const uint32_t sD = smem_base + IGEMM_HDR_BYTES;                       // 2 x 16 KiB
    const uint32_t sZ = sD + 2 * IGEMM_EPI_BUF_BYTES;                      // 2 x 16 KiB (if has_z)
    const uint32_t sStage = sZ + (p.has_z ? 2 * IGEMM_EPI_BUF_BYTES : 0);  // stages
    uint8_t* sD_gen = smem_gen + IGEMM_HDR_BYTES;
    uint8_t* sZ_gen = sD_gen + 2 * IGEMM_EPI_BUF_BYTES;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t num_taps = p.FH * p.FW;
    const uint32_t num_kb = num_taps * p.num_kc;
    const uint32_t total_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&p.tmA);
        prefetch_tensormap(&p.tmB);
        prefetch_tensormap(&p.tmD);
        if (p.has_z)
            prefetch_tensormap(&p.tmZ);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t i = 0; i < p.stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);  // one arrive per epilogue warp
            mbar_init(bar_zfull + 8 * i, 1);
            mbar_init(bar_zempty + 8 * i, 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ============================ TMA producer (A, B) ============================
        uint32_t stage = 0, phase = 0;
        const uint32_t OHW = p.OH * p.OW;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const uint32_t m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
            const int32_t m0 = static_cast<int32_t>(m_blk * IGEMM_BLOCK_M);
            const int32_t n0 = static_cast<int32_t>(n_blk * BN);
            const uint32_t img = static_cast<uint32_t>(m0) / OHW;
            const uint32_t rem = static_cast<uint32_t>(m0) % OHW;
            const int32_t base_h = static_cast<int32_t>(rem / p.OW) * p.stride_h - p.pad_h;
            const int32_t base_w = static_cast<int32_t>(rem % p.OW) * p.stride_w - p.pad_w;
            for (uint32_t tap = 0; tap < num_taps; ++tap) {
                const uint32_t fh = tap / p.FW, fw = tap % p.FW;
                // mode CONVOLUTION reads the spatially flipped filter tap for the same src pixel
                const uint32_t btap = p.flip ? (num_taps - 1 - tap) : tap;
                for (uint32_t kc = 0; kc < p.num_kc; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (lane == 0) {
                        const uint32_t full = bar_full + 8 * stage;
                        const uint32_t sa = sStage + stage * STAGE_BYTES;
                        const uint32_t sb = sa + A_BYTES;
                        const int32_t c0 = static_cast<int32_t>(kc * T::KB_ELEMS);
                        mbar_arrive_expect_tx(full, STAGE_BYTES);
                        if (p.a_mode == 0) {
                            tma_load_2d(sa, &p.tmA, full, c0, m0);
                        } else {
                            tma_load_im2col_4d(
                                    sa, &p.tmA, full, c0, base_w, base_h, static_cast<int32_t>(img),
                                    static_cast<uint16_t>(fw * p.dil_w),
                                    static_cast<uint16_t>(fh * p.dil_h));
                        }
                        tma_load_3d(sb, &p.tmB, full, c0, static_cast<int32_t>(btap), n0);
                    }
                    __syncwarp();
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BN;
            for (uint32_t kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = sStage + stage * STAGE_BYTES;
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (uint32_t k = 0; k < IGEMM_KBYTES / 32; ++k) {
                        const uint64_t adesc = umma_smem_desc(sa + k * 32, 1024, UMMA_LAYOUT_SW128);
                        const uint64_t bdesc = umma_smem_desc(sb + k * 32, 1024, UMMA_LAYOUT_SW128);
                        if constexpr (KIND == KIND_S8)
                            mma_i8(taddr, adesc, bdesc, IDESC, (kb | k) != 0);
                        else
                            mma_f16(taddr, adesc, bdesc, IDESC, (kb | k) != 0);
                    }
                    mma_commit(bar_empty + 8 * stage);
                    if (kb == num_kb - 1)
                        mma_commit(bar_tfull + 8 * acc);
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            acc ^= 1;
            if (acc == 0)
                acc_phase ^= 1;
        }
    } else if (warp == 2) {
        // ============================ z prefetch (TMA) ============================
        if (p.has_z) {
            uint32_t zb = 0, zphase = 0;
            for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const uint32_t m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
                const int32_t m0 = static_cast<int32_t>(m_blk * IGEMM_BLOCK_M);
                for (uint32_t c = 0; c < NUM_CHUNKS; ++c) {
                    const uint32_t col0 = n_blk * BN + c * T::EPI_N;
                    if (col0 >= p.OC)
                        break;
                    mbar_wait(bar_zempty + 8 * zb, zphase ^ 1);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(bar_zfull + 8 * zb, IGEMM_BLOCK_M * T::ROW_BYTES);
                        tma_load_2d(
                                sZ + zb * IGEMM_EPI_BUF_BYTES, &p.tmZ, bar_zfull + 8 * zb,
                                static_cast<int32_t>(col0), m0);
                    }
                    __syncwarp();
                    zb ^= 1;
                    if (zb == 0)
                        zphase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ============================ epilogue ============================
        const uint32_t et = threadIdx.x - 128;  // 0..127 == row of the tile == TMEM lane
        const uint32_t quad = warp & 3;
        uint32_t acc = 0, acc_phase = 0, zb = 0, zphase = 0, db = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const uint32_t m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
            const int32_t m0 = static_cast<int32_t>(m_blk * IGEMM_BLOCK_M);
            const uint32_t n0 = n_blk * BN;
            // bias slice of this tile -> smem (as fp32 for float kinds, int32 for int8)
            for (uint32_t i = et; i < BN; i += 128) {
                uint32_t v = 0;
                if (p.has_bias && n0 + i < p.OC) {
                    if constexpr (KIND == KIND_S8) {
                        v = static_cast<uint32_t>(reinterpret_cast<const int32_t*>(p.bias)[n0 + i]);
                    } else if constexpr (KIND == KIND_BF16) {
                        v = static_cast<uint32_t>(
                                    reinterpret_cast<const unsigned short*>(p.bias)[n0 + i])
                            << 16;
                    } else {
                        v = __float_as_uint(
                                __half2float(reinterpret_cast<const __half*>(p.bias)[n0 + i]));
                    }
                }
                sbias[i] = v;
            }
            named_bar_sync(1, 128);
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            uint32_t nchunks = 0;
            for (uint32_t c = 0; c < NUM_CHUNKS; ++c)
                if (n0 + c * T::EPI_N < p.OC)
                    nchunks = c + 1;
            for (uint32_t c = 0; c < nchunks; ++c) {
                const uint32_t col0 = n0 + c * T::EPI_N;
                // staging buffer db must have been drained by the TMA store issued 2 chunks ago
                if (et == 0)
                    tma_store_wait_read<1>();
                named_bar_sync(1, 128);
                if (p.has_z)
                    mbar_wait(bar_zfull + 8 * zb, zphase);
                uint8_t* drow = sD_gen + db * IGEMM_EPI_BUF_BYTES;
                const uint8_t* zrow = sZ_gen + zb * IGEMM_EPI_BUF_BYTES;
#pragma unroll
                for (uint32_t h = 0; h < T::EPI_N / 32; ++h) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((quad * 32) << 16) + acc * BN + c * T::EPI_N + h * 32, r);
                    tmem_ld_wait();
                    if (c == nchunks - 1 && h == T::EPI_N / 32 - 1) {
                        // accumulator fully drained into registers: hand TMEM back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0)
                            mbar_arrive(bar_tempty + 8 * acc);
                    }
                    const uint32_t* bcol = sbias + c * T::EPI_N + h * 32;
                    constexpr uint32_t COLS_PER_16B = 16 / T::OUT_ES;
#pragma unroll
                    for (uint32_t j = 0; j < 32 / COLS_PER_16B; ++j) {
                        // 16-byte piece j of this half-chunk of row `et`
                        const uint32_t off = ptx::swizzle_off<T::ROW_BYTES>(
                                et * T::ROW_BYTES + h * 32 * T::OUT_ES + j * 16);
                        uint4 zv = make_uint4(0, 0, 0, 0);
                        if (p.has_z)
                            zv = *reinterpret_cast<const uint4*>(zrow + off);
                        uint4 ov;
                        if constexpr (KIND == KIND_S8 && OUT32) {
                            // QuantizedS32 destination: dst = acc + bias (IDENTITY, no z)
                            const uint32_t* rr = r + j * 4;
                            const uint32_t* bb = bcol + j * 4;
                            ov.x = rr[0] + bb[0];
                            ov.y = rr[1] + bb[1];
                            ov.z = rr[2] + bb[2];
                            ov.w = rr[3] + bb[3];
                        } else if constexpr (KIND == KIND_S8) {
                            const uint32_t* rr = r + j * 16;
                            const uint32_t* bb = bcol + j * 16;
                            uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
                            uint32_t ow[4];
#pragma unroll
                            for (uint32_t w = 0; w < 4; ++w) {
                                uint32_t packed = 0;
#pragma unroll
                                for (uint32_t b = 0; b < 4; ++b) {
                                    const int32_t sfb = static_cast<int32_t>(rr[w * 4 + b] + bb[w * 4 + b]);
                                    const int32_t zq = static_cast<int32_t>(
                                            static_cast<int8_t>((zw[w] >> (8 * b)) & 0xff));
                                    const int32_t q = requant_s8(
                                            sfb, p.s_bias, p.has_z != 0, zq, p.s_z, p.nonline, p.s_dst);
                                    packed |= (static_cast<uint32_t>(q) & 0xff) << (8 * b);
                                }
                                ow[w] = packed;
                            }
                            ov = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                        } else {
                            const uint32_t* rr = r + j * 8;
                            const uint32_t* bb = bcol + j * 8;
                            uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
                            uint32_t ow[4];
#pragma unroll
                            for (uint32_t w = 0; w < 4; ++w) {
                                float x0 = __uint_as_float(rr[2 * w]) + __uint_as_float(bb[2 * w]);
                                float x1 = __uint_as_float(rr[2 * w + 1]) + __uint_as_float(bb[2 * w + 1]);
                                if (p.has_z) {
                                    if constexpr (KIND == KIND_BF16) {
                                        x0 += bf16lo(zw[w]);
                                        x1 += bf16hi(zw[w]);
                                    } else {
                                        x0 += f16lo(zw[w]);
                                        x1 += f16hi(zw[w]);
                                    }
                                }
                                x0 = act_f(x0, p.nonline);
                                x1 = act_f(x1, p.nonline);
                                ow[w] = (KIND == KIND_BF16) ? pack_bf16x2(x0, x1) : pack_f16x2(x0, x1);
                            }
                            ov = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                        }
                        *reinterpret_cast<uint4*>(drow + off) = ov;
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (et == 0) {
                    tma_store_2d(&p.tmD, sD + db * IGEMM_EPI_BUF_BYTES, static_cast<int32_t>(col0), m0);
                    tma_store_commit();
                    if (p.has_z)
                        mbar_arrive(bar_zempty + 8 * zb);
                }
                db ^= 1;
                if (p.has_z) {
                    zb ^= 1;
                    if (zb == 0)
                        zphase ^= 1;
                }
            }
            acc ^= 1;
            if (acc == 0)
                acc_phase ^= 1;
        }
        if (et == 0)
            tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int KIND, int BN, int OUT32>
cudaError_t launch_variant(const IgemmParams& p, uint32_t grid, uint32_t smem, cudaStream_t stream) {
    auto kern = igemm_kernel<KIND, BN, OUT32>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(MAX_DYN_SMEM));
        if (e != cudaSuccess)
            return e;
        attr_set[dev] = true;
    }
    kern<<<grid, 256, smem, stream>>>(p);
    return cudaGetLastError();
}

uint32_t stage_bytes(uint32_t bn) {
    return IGEMM_BLOCK_M * IGEMM_KBYTES + bn * IGEMM_KBYTES;
}

}  // namespace

// --------------------------------------------------------------------------- host side
int igemm_query(const b200cb_desc& d, uint32_t block_n, b200cb_plan* plan) {
    if (d.format != B200CB_FORMAT_NHWC)
        return B200CB_UNSUPPORTED_FORMAT;
    if (d.sparse != B200CB_SPARSE_DENSE || d.group != 1)
        return B200CB_UNSUPPORTED_PARAM;
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
    bool out32 = false;
    if (kind == KIND_S8) {
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
    }
    if (d.bias_mode == B200CB_BIAS_FULL)
        return B200CB_UNSUPPORTED_PARAM;
    if (block_n != 64 && block_n != 128 && block_n != 256)
        return B200CB_UNSUPPORTED_ALGO;
    const uint32_t in_es = kind == KIND_S8 ? 1 : 2;
    const uint32_t out_es = out32 ? 4 : in_es;
    // TMA needs 16-byte aligned global strides
    if ((d.ic * in_es) % 16 != 0 || (d.oc * out_es) % 16 != 0)
        return B200CB_UNSUPPORTED_ALIGNMENT;
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

    const uint32_t fixed = 1024 + IGEMM_HDR_BYTES + 2 * IGEMM_EPI_BUF_BYTES +
                           (d.has_z ? 2 * IGEMM_EPI_BUF_BYTES : 0);
    uint32_t stages = (MAX_DYN_SMEM - fixed) / stage_bytes(block_n);
    if (stages > IGEMM_MAX_STAGES)
        stages = IGEMM_MAX_STAGES;
    if (stages < 2)
        return B200CB_UNSUPPORTED_SHAPE;

    plan->magic = PLAN_MAGIC;
    plan->kernel = KERNEL_IGEMM;
    plan->block_n = block_n;
    plan->stages = stages;
    plan->grid = 0;  // resolved at launch from the device's SM count
    plan->smem_bytes = fixed + stages * stage_bytes(block_n);
    const bool pointwise = d.fh == 1 && d.fw == 1 && d.stride_h == 1 && d.stride_w == 1 &&
                           d.pad_h == 0 && d.pad_w == 0 &&
                           (d.src_batch_stride == 0 || d.src_batch_stride == src_img);
    plan->a_mode = pointwise ? 0 : 1;
    for (auto& r : plan->reserved)
        r = 0;
    plan->reserved[0] = static_cast<uint32_t>(kind);
    plan->reserved[1] = out32 ? 1 : 0;
    plan->workspace_bytes = 0;
    plan->desc = d;
    return B200CB_OK;
}

int igemm_launch(
        const b200cb_plan& plan, const void* src, const void* filter, const void* bias,
        const void* z, void* dst, void* /*workspace*/, size_t /*workspace_bytes*/,
        cudaStream_t stream) {
    const b200cb_desc& d = plan.desc;
    const int kind = static_cast<int>(plan.reserved[0]);
    const bool out32 = plan.reserved[1] != 0;
    const uint32_t in_es = kind == KIND_S8 ? 1 : 2;
    const uint32_t out_es = out32 ? 4 : in_es;
    const uint32_t kb_elems = IGEMM_KBYTES / in_es;
    const uint32_t epi_n = out32 ? 32 : 64;
    const uint32_t row_bytes = epi_n * out_es;
    const uint32_t BN = plan.block_n;
    uint32_t a_mode = plan.a_mode;
    if (plan.reserved[2] == 1)
        a_mode = 1;  // debugging override: force the im2col path on pointwise convs

    if (!src || !filter || !dst || (d.bias_mode != B200CB_BIAS_NONE && !bias) || (d.has_z && !z))
        return B200CB_ERR_NULL;

    IgemmParams p;
    memset(&p, 0, sizeof(p));
    const uint64_t M = (uint64_t)d.n * d.oh * d.ow;
    const CUtensorMapDataType in_dt =
            kind == KIND_S8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                            : (kind == KIND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    const CUtensorMapDataType out_dt = out32 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : in_dt;
    const uint64_t src_bstride = d.src_batch_stride ? d.src_batch_stride : (uint64_t)d.ih * d.iw * d.ic;
    CUresult cr;
    if (a_mode == 0) {
        uint64_t dims[2] = {d.ic, M};
        uint64_t strides[1] = {(uint64_t)d.ic * in_es};
        uint32_t box[2] = {kb_elems, IGEMM_BLOCK_M};
        cr = encode_tiled(
                &p.tmA, in_dt, 2, src, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    } else {
        uint64_t dims[4] = {d.ic, d.iw, d.ih, d.n};
        uint64_t strides[3] = {
                (uint64_t)d.ic * in_es, (uint64_t)d.iw * d.ic * in_es, src_bstride * in_es};
        int lower[2] = {-(int)d.pad_w, -(int)d.pad_h};
        int upper[2] = {
                (int)d.pad_w - (int)((d.fw - 1) * d.dilate_w),
                (int)d.pad_h - (int)((d.fh - 1) * d.dilate_h)};
        uint32_t estr[4] = {1, d.stride_w, d.stride_h, 1};
        cr = encode_im2col(
                &p.tmA, in_dt, 4, src, dims, strides, lower, upper, kb_elems, IGEMM_BLOCK_M, estr,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                src_bstride * in_es * d.n);
    }
    if (cr != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof(buf), "CUresult %d (activation map, a_mode %u)", (int)cr, a_mode);
        set_last_error("cuTensorMapEncode", buf);
        return B200CB_ERR_DRIVER;
    }
    {
        uint64_t dims[3] = {d.ic, (uint64_t)d.fh * d.fw, d.oc};
        uint64_t strides[2] = {(uint64_t)d.ic * in_es, (uint64_t)d.fh * d.fw * d.ic * in_es};
        uint32_t box[3] = {kb_elems, 1, BN};
        cr = encode_tiled(
                &p.tmB, in_dt, 3, filter, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (cr != CUDA_SUCCESS) {
            char buf[64];
            snprintf(buf, sizeof(buf), "CUresult %d (filter map)", (int)cr);
            set_last_error("cuTensorMapEncodeTiled", buf);
            return B200CB_ERR_DRIVER;
        }
    }
    {
        const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        uint64_t dims[2] = {d.oc, M};
        uint64_t strides[1] = {(uint64_t)d.oc * out_es};
        uint32_t box[2] = {epi_n, IGEMM_BLOCK_M};
        cr = encode_tiled(
                &p.tmD, out_dt, 2, dst, dims, strides, box, nullptr, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        if (cr == CUDA_SUCCESS && d.has_z)
            cr = encode_tiled(
                    &p.tmZ, out_dt, 2, z, dims, strides, box, nullptr, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        if (cr != CUDA_SUCCESS) {
            char buf[64];
            snprintf(buf, sizeof(buf), "CUresult %d (dst/z map)", (int)cr);
            set_last_error("cuTensorMapEncodeTiled", buf);
            return B200CB_ERR_DRIVER;
        }
    }
    p.bias = bias;
    p.M = (uint32_t)M;
    p.OC = d.oc;
    p.IC = d.ic;
    p.OH = d.oh;
    p.OW = d.ow;
    p.FH = d.fh;
    p.FW = d.fw;
    p.stride_h = d.stride_h;
    p.stride_w = d.stride_w;
    p.pad_h = d.pad_h;
    p.pad_w = d.pad_w;
    p.dil_h = d.dilate_h;
    p.dil_w = d.dilate_w;
    p.num_kc = (d.ic + kb_elems - 1) / kb_elems;
    p.m_tiles = (uint32_t)((M + IGEMM_BLOCK_M - 1) / IGEMM_BLOCK_M);
    p.n_tiles = (d.oc + BN - 1) / BN;
    p.stages = plan.stages;
    p.a_mode = a_mode;
    p.flip = d.mode == B200CB_MODE_CONVOLUTION;
    p.has_bias = d.bias_mode != B200CB_BIAS_NONE;
    p.has_z = d.has_z;
    p.nonline = d.nonline_mode;
    p.s_bias = d.bias_scale;
    p.s_z = d.z_scale;
    p.s_dst = d.dst_scale;

    const uint64_t total = (uint64_t)p.m_tiles * p.n_tiles;
    uint32_t grid = plan.grid ? plan.grid : (uint32_t)device_sm_count();
    if (grid > total)
        grid = (uint32_t)total;

    cudaError_t e = cudaErrorInvalidValue;
#define B200CB_DISPATCH(KIND_, OUT_)                                                          \
    switch (BN) {                                                                             \
        case 64: e = launch_variant<KIND_, 64, OUT_>(p, grid, plan.smem_bytes, stream); break;   \
        case 128: e = launch_variant<KIND_, 128, OUT_>(p, grid, plan.smem_bytes, stream); break; \
        case 256: e = launch_variant<KIND_, 256, OUT_>(p, grid, plan.smem_bytes, stream); break; \
        default: break;                                                                       \
    }
    if (kind == KIND_BF16) {
        B200CB_DISPATCH(KIND_BF16, 0)
    } else if (kind == KIND_F16) {
        B200CB_DISPATCH(KIND_F16, 0)
    } else if (!out32) {
        B200CB_DISPATCH(KIND_S8, 0)
    } else {
        B200CB_DISPATCH(KIND_S8, 1)
    }
#undef B200CB_DISPATCH
    if (e != cudaSuccess) {
        set_last_error("igemm launch", cudaGetErrorString(e));
        return B200CB_ERR_CUDA;
    }
    return B200CB_OK;
}

}  // namespace b200cb

// tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (bf16 operands, fp32 accumulate in TMEM)
// with the fused epilogues the speech-transformer hot path needs (bias / ReLU / GLU / x*sqrt(d)+PE /
// residual / residual+LayerNorm / swish / gelu / tanh) and an implicit-GEMM mode for the 3x3 stride-2
// Conv2d subsampling layer (A gathered by a 5-D TMA tensor map, one 64-channel tap per k-block).
//
// Replaces the cuBLAS/oneDNN calls behind nn.Linear / nn.Conv2d at
//   otrans/module/attention.py:68,128-129,44   otrans/module/ffn.py:39-41
//   otrans/frontend/conv.py:63-64,146          otrans/decoder/transformer.py:181
//
// Structure (persistent, one CTA per SM, 384 threads):
//   warp 0 lane 0 : TMA producer   (A tile 128x64, B tile BNx64 per k-block, SWIZZLE_128B)
//   warp 1 lane 0 : tcgen05.mma issuer (UMMA 128 x BN x 16, 4 per k-block), tcgen05.commit -> mbarriers
//   warp 2        : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4..11   : epilogue, two warpgroups: warp w reads TMEM lanes 32*(w%4).., warpgroup (w-4)/4 takes
//                   one half of the tile's columns.  tcgen05.ld -> registers -> fused math (bias / LN
//                   parameters broadcast from smem) -> staged in smem -> coalesced 16-byte global stores
//                   (one full output row segment per warp instruction; the residual tile is fetched the
//                   same way).  Round-1 profile: per-thread row stores were 32 sectors/request and the
//                   epilogue, not the MMA, bounded every GEMM (profiles/r1_*).
// Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static
// round-robin tile scheduler, so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "otb_internal.h"
#include "ptx.cuh"

namespace otb {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kThreads = 384;
static constexpr int kEpiThreads = 256;
static constexpr int STG_PITCH_MAX = 512 + 16;          // bytes per staged row (+16 B pad: conflict-free 16 B accesses)
static constexpr int STG_BYTES = BM * STG_PITCH_MAX;    // 67,584
static constexpr int AUX_BYTES = 6144;                  // barriers, TMEM slot, bias / gamma / beta, row map, LN stats

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 3 : (BN == 128 ? 4 : 6);
    static constexpr int TMEM_COLS = 2 * BN;  // power of two for BN in {64,128,256}
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + AUX_BYTES + 1024;
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 fp32 values -> staged row (bf16 or f32), 16-byte shared stores
__device__ __forceinline__ void stage32(uint8_t* row_ptr, int col, int out_f32, const float (&v)[32]) {
    if (out_f32) {
        float4* d = reinterpret_cast<float4*>(row_ptr + col * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        uint4* d = reinterpret_cast<uint4*>(row_ptr + col * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16(v[8 * i], v[8 * i + 1]);
            u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
            d[i] = u;
        }
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BN_OUT = (EPI == EPI_GLU) ? BN / 2 : BN;  // output columns per tile
    constexpr int HALF = BN_OUT / 2;                        // output columns per epilogue warpgroup

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stg = smem + STAGES * Cfg::STAGE_BYTES;
    uint8_t* aux = stg + STG_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(aux + 256);   // [256]  (GLU: value half | gate half)
    float* s_gamma = s_bias + 256;                          // [256]
    float* s_beta = s_gamma + 256;                          // [256]
    int* s_rowmap = reinterpret_cast<int*>(s_beta + 256);   // [128] output row of each tile row, -1 = skip
    float2* s_stats = reinterpret_cast<float2*>(s_rowmap + 128);  // [2][128] partial (sum, sumsq)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = p.conv ? (p.conv_B * p.conv_T1h + p.conv_R - 1) / p.conv_R : (p.M + BM - 1) / BM;
    const int n_tiles = (p.N + BN_OUT - 1) / BN_OUT;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.K + BK - 1) / BK;
    const uint32_t a_tx = p.conv ? (uint32_t)(p.conv_R * p.conv_F2 * BK * 2) : (uint32_t)Cfg::A_BYTES;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sb = sa + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], a_tx + (uint32_t)Cfg::B_BYTES);
                if (p.conv) {
                    // k-block -> (tap, channel chunk); tap (kh,kw): input row 2t'+kh, col 2f'+kw-1
                    const int tap = kb / p.conv_cchunks, cc = kb % p.conv_cchunks;
                    const int kh = tap / 3, kw = tap % 3;
                    const int par_f = (kw == 1) ? 0 : 1, f0 = (kw == 0) ? -1 : 0;
                    const int par_t = (kh == 1) ? 1 : 0, dt = (kh == 2) ? 1 : 0;
                    tma_load_5d(sa, &tmA, &full_bar[stage], cc * BK, par_f, f0, par_t, m_blk * p.conv_R + dt);
                } else {
                    tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                }
                if (EPI == EPI_GLU) {
                    tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN_OUT);
                    tma_load_2d(sb + Cfg::B_BYTES / 2, &tmB, &full_bar[stage], kb * BK, p.N + n_blk * BN_OUT);
                } else {
                    tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_bf16(BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(&tempty_bar[as], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                              (uint32_t)((kb | k) != 0));
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (256 threads)
        const int et = threadIdx.x - 128;       // 0..255
        const int ewarp = warp - 4;             // 0..7
        const int quad = ewarp & 3;             // TMEM lane quadrant == warp % 4
        const int half = ewarp >> 2;            // column half handled by this warpgroup
        const int row_in_tile = quad * 32 + lane;
        const int esize = p.out_f32 ? 4 : 2;
        const int pitch = BN_OUT * esize + 16;
        uint8_t* my_row = stg + row_in_tile * pitch;

        if (EPI == EPI_RESID_LN) {
            for (int i = et; i < BN; i += kEpiThreads) {
                s_gamma[i] = p.gamma[i];
                s_beta[i] = p.beta[i];
            }
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int col_base = n_blk * BN_OUT;             // first output column of the tile
            const int ncols = min(BN_OUT, p.N - col_base);   // valid output columns

            // ---- (a) per-tile smem setup: bias slice, row map, residual tile (all coalesced)
            for (int i = et; i < BN; i += kEpiThreads) {
                int c;
                if (EPI == EPI_GLU) c = (i < BN_OUT) ? col_base + i : p.N + col_base + (i - BN_OUT);
                else c = col_base + i;
                const bool ok = (EPI == EPI_GLU) ? ((i % BN_OUT) < ncols) : (i < ncols);
                s_bias[i] = (p.bias != nullptr && ok) ? p.bias[c] : 0.f;
            }
            if (et < BM) {
                int out_row;
                bool ok;
                if (p.conv) {
                    const int r = et / p.conv_F2, f = et % p.conv_F2;
                    const int bt = m_blk * p.conv_R + r;
                    const int b = bt / p.conv_T1h, t = bt % p.conv_T1h;
                    ok = (r < p.conv_R) && (b < p.conv_B) && (t < p.conv_T2);
                    out_row = (b * p.conv_T2 + t) * p.conv_F2 + f;
                } else {
                    out_row = m_blk * BM + et;
                    ok = out_row < p.M;
                }
                s_rowmap[et] = ok ? out_row : -1;
            }
            if (EPI == EPI_RESID || EPI == EPI_RESID_LN) {
                // residual rows are contiguous (no conv mode): warp copies one row segment per instruction
                for (int r = ewarp; r < BM; r += 8) {
                    const int grow = m_blk * BM + r;
                    if (grow >= p.M) continue;
                    const bf16* src = p.resid + (size_t)grow * p.ldr + col_base;
                    uint8_t* dst = stg + r * pitch;
                    for (int c = lane * 8; c < ncols; c += 256) {
                        if (c + 8 <= ncols) {
                            *reinterpret_cast<uint4*>(dst + c * 2) = *reinterpret_cast<const uint4*>(src + c);
                        } else {
                            for (int j = c; j < ncols; ++j) reinterpret_cast<bf16*>(dst)[j] = src[j];
                        }
                    }
                }
            }
            epi_bar();

            // row validity / padding mask of this thread's row
            const int out_row = s_rowmap[row_in_tile];
            bool row_live = true;
            if (p.row_len != nullptr && out_row >= 0) {
                const int b = out_row / p.row_period, t = out_row % p.row_period;
                row_live = t < p.row_len[b];
            }

            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);

            if (EPI == EPI_RESID_LN) {
                // pass 1: v = resid + acc + bias parked back in TMEM; row statistics over both column halves
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(t_row + c, r);
                    tmem_ld_wait();
                    const uint4* rs = reinterpret_cast<const uint4*>(my_row + c * 2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 u = rs[i];
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = unpack_bf16(w[j]);
                            const int e = 8 * i + 2 * j;
                            float v0 = __uint_as_float(r[e]) + s_bias[c + e];
                            float v1 = __uint_as_float(r[e + 1]) + s_bias[c + e + 1];
                            v0 = (row_live ? v0 : 0.f) + f.x;
                            v1 = (row_live ? v1 : 0.f) + f.y;
                            s1 += v0 + v1;
                            s2 += v0 * v0 + v1 * v1;
                            r[e] = __float_as_uint(v0);
                            r[e + 1] = __float_as_uint(v1);
                        }
                    }
                    tmem_st32(t_row + c, r);
                }
                tmem_st_wait();
                s_stats[half * BM + row_in_tile] = make_float2(s1, s2);
                epi_bar();
                const float2 o = s_stats[(half ^ 1) * BM + row_in_tile];
                const float mean = (s1 + o.x) * (1.0f / BN);
                const float var = fmaxf((s2 + o.y) * (1.0f / BN) - mean * mean, 0.f);
                const float rstd = rsqrtf(var + p.eps);
#pragma unroll 1
                for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(t_row + c, r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        v[i] = (__uint_as_float(r[i]) - mean) * rstd * s_gamma[c + i] + s_beta[c + i];
                    stage32(my_row, c, p.out_f32, v);
                }
            } else if (EPI == EPI_GLU) {
#pragma unroll 1
                for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
                    uint32_t ra[32], rg[32];
                    tmem_ld32(t_row + c, ra);
                    tmem_ld32(t_row + BN_OUT + c, rg);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float a = __uint_as_float(ra[i]) + s_bias[c + i];
                        const float g = __uint_as_float(rg[i]) + s_bias[BN_OUT + c + i];
                        v[i] = row_live ? a * fast_sigmoid(g) : 0.f;
                    }
                    stage32(my_row, c, p.out_f32, v);
                }
            } else {
#pragma unroll 1
                for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(t_row + c, r);
                    tmem_ld_wait();
                    if (c >= ncols) continue;  // warp-uniform
                    float v[32];
                    float res[32];
                    if (EPI == EPI_RESID) {
                        const uint4* rs = reinterpret_cast<const uint4*>(my_row + c * 2);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 u = rs[i];
                            const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z),
                                         f3 = unpack_bf16(u.w);
                            res[8 * i] = f0.x; res[8 * i + 1] = f0.y; res[8 * i + 2] = f1.x; res[8 * i + 3] = f1.y;
                            res[8 * i + 4] = f2.x; res[8 * i + 5] = f2.y; res[8 * i + 6] = f3.x; res[8 * i + 7] = f3.y;
                        }
                    }
                    const float* trow = nullptr;
                    if (EPI == EPI_TABLE)
                        trow = p.table + (size_t)((out_row >= 0 ? out_row : 0) % p.period) * p.N + col_base + c;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float x = __uint_as_float(r[i]) + s_bias[c + i];
                        if (EPI == EPI_RELU) x = fmaxf(x, 0.f);
                        if (EPI == EPI_SWISH) x = x * fast_sigmoid(x);
                        if (EPI == EPI_GELU) x = 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
                        if (EPI == EPI_TANH) x = tanhf(x);
                        if (EPI == EPI_TABLE) x = x * p.alpha + ((c + i < ncols) ? __ldg(trow + i) : 0.f);
                        if (!row_live) x = 0.f;
                        if (EPI == EPI_RESID) x = res[i] + p.alpha * x;
                        v[i] = x;
                    }
                    stage32(my_row, c, p.out_f32, v);
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[as]);  // TMEM stage is free for the next tile's MMAs
            epi_bar();

            // ---- (e) coalesced copy-out: one output row segment per warp instruction
            const int row_bytes = ncols * esize;
            for (int r = ewarp; r < BM; r += 8) {
                const int orow = s_rowmap[r];
                if (orow < 0) continue;
                const uint8_t* src = stg + r * pitch;
                uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((size_t)orow * p.ldc + col_base) * esize;
                for (int off = lane * 16; off < row_bytes; off += 512) {
                    if (off + 16 <= row_bytes) {
                        *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(src + off);
                    } else {
                        for (int j = off; j < row_bytes; j += 2)
                            *reinterpret_cast<uint16_t*>(dst + j) = *reinterpret_cast<const uint16_t*>(src + j);
                    }
                }
            }
            epi_bar();  // staging / bias smem may be rewritten for the next tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

const char* encode_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems,
                           uint32_t box_cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled unavailable (no CUDA driver?)";
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_elems & 7)) return "TMA operand must be 16-byte aligned with ld % 8 == 0";
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled(2d) failed";
}

// conv1 output buffer [B, 2*T1h, 2*F1h, C] bf16 viewed as (c, f-parity, f/2, t-parity, b*T1h + t/2)
const char* encode_tmap_conv5d(CUtensorMap* m, const void* base, int C, int F1h, int T1h_total, uint32_t boxF,
                               uint32_t boxR) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled unavailable (no CUDA driver?)";
    cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)F1h, 2, (cuuint64_t)T1h_total};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 4, (cuuint64_t)F1h * 2 * C * 2,
                             (cuuint64_t)F1h * 2 * C * 4};
    cuuint32_t box[5] = {64, 1, boxF, 1, boxR};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled(5d) failed";
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN, int EPI>
static const char* launch_inst(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                               int num_tiles) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES) != cudaSuccess)
            return "cudaFuncSetAttribute(max dynamic smem) failed";
        attr_set = true;
    }
    int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    gemm_tc_kernel<BN, EPI><<<grid, kThreads, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <int BN>
static const char* launch_bn(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                             int epi, int num_tiles) {
    switch (epi) {
        case EPI_BIAS: return launch_inst<BN, EPI_BIAS>(st, ta, tb, p, num_tiles);
        case EPI_RELU: return launch_inst<BN, EPI_RELU>(st, ta, tb, p, num_tiles);
        case EPI_GLU: return launch_inst<BN, EPI_GLU>(st, ta, tb, p, num_tiles);
        case EPI_TABLE: return launch_inst<BN, EPI_TABLE>(st, ta, tb, p, num_tiles);
        case EPI_RESID: return launch_inst<BN, EPI_RESID>(st, ta, tb, p, num_tiles);
        case EPI_RESID_LN: return launch_inst<BN, EPI_RESID_LN>(st, ta, tb, p, num_tiles);
        case EPI_SWISH: return launch_inst<BN, EPI_SWISH>(st, ta, tb, p, num_tiles);
        case EPI_GELU: return launch_inst<BN, EPI_GELU>(st, ta, tb, p, num_tiles);
        case EPI_TANH: return launch_inst<BN, EPI_TANH>(st, ta, tb, p, num_tiles);
    }
    return "unknown epilogue";
}

// Pick the N tile that wastes the fewest MMA cycles across the persistent grid.
static int choose_bn(int m_tiles, int n_cols, int epi, int out_f32) {
    if (epi == EPI_RESID_LN) return n_cols;  // tile must span the row
    const int sms = num_sms();
    int best = 0;
    double best_cost = 1e30;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const int bn_out = (epi == EPI_GLU) ? bn / 2 : bn;
        if (bn_out * (out_f32 ? 4 : 2) > 512) continue;  // staged row must fit the staging pitch
        if (bn_out < 64) continue;                       // each epilogue warpgroup needs >= 32 columns
        const int n_tiles = (n_cols + bn_out - 1) / bn_out;
        const long tiles = (long)m_tiles * n_tiles;
        const long waves = (tiles + sms - 1) / sms;
        const double cost = (double)waves * (bn + 48);  // +48: per-tile fixed overhead in "column" units
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

const char* gemm_launch(cudaStream_t st, const void* A, int lda, const void* W, int ldw, int w_rows, int epi,
                        GemmParams p, const CUtensorMap* conv_map) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) return "gemm: empty problem";
    if (p.K % 8) return "gemm: K must be a multiple of 8";
    if (p.ldc % (p.out_f32 ? 4 : 8)) return "gemm: ldc must keep rows 16-byte aligned (ldc % 8 == 0 for bf16, % 4 for f32)";
    if (reinterpret_cast<uintptr_t>(p.out) & 15) return "gemm: out must be 16-byte aligned";
    if ((epi == EPI_RESID || epi == EPI_RESID_LN) && ((p.ldr % 8) || (reinterpret_cast<uintptr_t>(p.resid) & 15)))
        return "gemm: residual must be 16-byte aligned with ldr % 8 == 0";
    const int m_tiles = p.conv ? (p.conv_B * p.conv_T1h + p.conv_R - 1) / p.conv_R : (p.M + BM - 1) / BM;
    if (epi == EPI_RESID_LN && !(p.N == 64 || p.N == 128 || p.N == 256))
        return "gemm: fused residual+LayerNorm epilogue needs N in {64,128,256}";
    if (epi == EPI_RESID_LN && p.out_f32) return "gemm: fused residual+LayerNorm epilogue writes bf16";
    const int bn = choose_bn(m_tiles, p.N, epi, p.out_f32);
    if (bn == 0) return "gemm: no tile configuration";
    const int bn_out = (epi == EPI_GLU) ? bn / 2 : bn;
    const int n_tiles = (p.N + bn_out - 1) / bn_out;
    const int num_tiles = m_tiles * n_tiles;

    CUtensorMap ta, tb;
    const char* err;
    if (p.conv) {
        if (!conv_map) return "gemm: conv mode needs a conv tensor map";
        ta = *conv_map;
    } else {
        if ((err = encode_tmap_2d(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda, BK, BM))) return err;
    }
    const uint32_t box_rows = (epi == EPI_GLU) ? (uint32_t)(bn / 2) : (uint32_t)bn;
    if ((err = encode_tmap_2d(&tb, W, (uint64_t)p.K, (uint64_t)w_rows, (uint64_t)ldw, BK, box_rows))) return err;

    switch (bn) {
        case 256: return launch_bn<256>(st, ta, tb, p, epi, num_tiles);
        case 128: return launch_bn<128>(st, ta, tb, p, epi, num_tiles);
        case 64: return launch_bn<64>(st, ta, tb, p, epi, num_tiles);
    }
    return "gemm: bad tile";
}

}  // namespace otb

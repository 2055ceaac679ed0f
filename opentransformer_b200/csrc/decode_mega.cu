// Persistent beam-search decode kernel for sm_100a: the WHOLE decode loop of SpeechToTextRecognizer.recognize
// (otrans/recognize/speech2text.py:60-68 -> decode_step :95-153 -> TransformerDecoder.inference,
// otrans/decoder/transformer.py:185-208) in ONE launch.
//
// Why: the per-step graph of ~52 kernels (GEMMs on 12..100 CTAs, 4..32 k-blocks each) is bound by per-kernel
// fixed latency -- 0.50 ms per beam step, 95 % of a recognize pass (profiles/r1_launches_bench_v4.csv).  But the
// decode loop of one utterance never needs another utterance: self-attention is per hypothesis, cross-attention and
// the beam pruning are per utterance, everything else is row-wise.  So one thread-block CLUSTER of 4 CTAs owns one
// utterance for all steps; there is no grid-wide synchronisation and no launch inside the loop.
//
//   * rows: the `beam` (<= 16) hypotheses of the utterance = one m16 MMA tile.  With 10 rows the GEMMs are pure
//     weight streaming (25.9 MB of bf16 decoder weights per step, L2-resident): mma.sync.m16n8k16 fed by 16-byte
//     read-only loads straight from L2 (16 in flight per lane), activations in shared memory.  tcgen05 (M >= 64
//     rows per instruction) has nothing to offer a 10-row problem; the bound is L2 -> SM bandwidth.
//   * cluster rank c == attention head c (d_model 256 = 4 x 64): QKV / q projections are split by head, the other
//     projections by output column (64 per CTA), the GLU feed-forward by hidden feature (512 per CTA).  Results every
//     CTA needs (attention context, pre-LayerNorm rows, GLU activations, soft-max statistics, top-k candidates) are
//     written into ALL four CTAs' shared memory through DSMEM stores followed by one cluster barrier; LayerNorm and
//     the beam step are then computed redundantly by each CTA, so the search state (scores, flags, newest tokens,
//     ancestry table) lives replicated in shared memory and never crosses the cluster.
//   * self-attention K/V cache in HBM/L2 ([layer, step, hyp, d], addressed through the 1-byte ancestry table: beam
//     reordering never moves K/V rows); cross-attention K/V are projected once per utterance by the tcgen05 GEMM.
//   * log-softmax + top-k + finished-hypothesis masking + beam^2 pruning follow beam.cu exactly (ties -> lower
//     index), so ids / parents are bit-exact with the oracle's beam_step driven by this kernel's log-probs.
//
// Supported: post-norm decoder (normalize_before False), GLU feed-forward, d_model 256, 4 heads, d_ff % 512 == 0,
// beam <= 16, max_len <= 128.  Other configurations use the per-step graph path (recognize.BeamDecoder.step).
#include <string.h>

#include "beam_common.cuh"
#include "otb_internal.h"
#include "ptx.cuh"

namespace otb {

static constexpr int MG_THREADS = 256;
static constexpr int MG_C = 4;        // CTAs per cluster == attention heads
static constexpr int MG_D = 256;      // d_model
static constexpr int MG_XP = 576;     // bytes per activation row in smem: 512 + 64 (pitch % 128 == 64 -> conflict-free LDS.128)
static constexpr int MG_LMAX = 128;   // cached positions

unsigned long long* g_mega_dbg = nullptr;   // otb_debug_mega_timing: clock64 stamps of cluster 0 / rank 0 at step g_mega_dbg_step
int g_mega_dbg_step = 0;
#define MG_STAMP() do { if (dbg_on) p.dbg_clk[dbg_n++] = clock64(); } while (0)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {   // read-only, do not pollute L1 (every weight byte is used once per step)
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ float mg_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// One warp streams `ntile` n8-tiles of Y[16 x N] = X[16 x K] W^T.
//   X: bf16 rows in shared memory, `xpitch` bytes apart (xpitch % 128 == 64); K % 256 == 0
//   W: bf16 [rows, ldw] row-major in global memory; row_of(i) = first of the 8 consecutive W rows of the warp's i-th
//      tile; rows >= row_lim are clamped (their outputs are garbage and must be ignored by `epi`)
//   epi(i, acc): acc[0..1] = (row g, cols 2t, 2t+1), acc[2..3] = (row g+8, same cols) of tile i, g = lane/4, t = lane%4
// A 16-byte weight load per lane feeds two MMAs: within a 32-wide k slice, lane t's chunk k = 8t..8t+7 is used as
// logical k pairs (2t, 2t+8) of the first MMA and of the second -- the same permutation is applied to the A fragment, so
// the contraction is unchanged.  Work unit = (tile, 256-wide k chunk) = 8 loads per lane, three units buffered in registers.
// `rot` rotates the order in which the warp walks its tiles (tile (i + rot) % ntile first): the 32 clusters of a launch
// stream the SAME weights in the same order at the same time, i.e. they all hit the same few L2 slices at any moment;
// starting each cluster at a different tile spreads the requests over the slices.
template <class RowFn, class EpiFn>
__device__ __forceinline__ void stream_tiles(const uint8_t* xs, int xpitch, int K, const bf16* __restrict__ W, int ldw,
                                             int row_lim, int ntile, int rot, RowFn row_of_, EpiFn epi_) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int kch = K >> 8;
    const int nunit = ntile * kch;
    if (nunit <= 0) return;
    // a single tile (the K = 2048 projection: 8 chunks): rotate the order of its k chunks instead (fp32 sum order only)
    const int krot = (ntile == 1) ? rot % kch : 0;
    rot = rot % ntile;
    auto kc_of = [&](int kc) { int k2 = kc + krot; return k2 >= kch ? k2 - kch : k2; };
    auto row_of = [&](int i) { int j = i + rot; if (j >= ntile) j -= ntile; return row_of_(j); };
    auto epi = [&](int i, const float(&a)[4]) { int j = i + rot; if (j >= ntile) j -= ntile; epi_(j, a); };
    uint4 w0[8], w1[8], w2[8];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    auto issue = [&](uint4(&w)[8], int u) {
        const int tile = u / kch, kc = u - tile * kch;
        int row = row_of(tile) + g;
        row = row < row_lim ? row : row_lim - 1;
        const uint4* src = reinterpret_cast<const uint4*>(W + (size_t)row * ldw + kc_of(kc) * 256 + t * 8);
#pragma unroll
        for (int s = 0; s < 8; ++s) w[s] = ldg_stream(src + s * 4);
    };
    auto compute = [&](const uint4(&w)[8], int u) {
        const int tile = u / kch, kc = u - tile * kch;
        if (kc == 0) { acc[0] = 0.f; acc[1] = 0.f; acc[2] = 0.f; acc[3] = 0.f; }
        const uint8_t* xa = xs + g * xpitch + kc_of(kc) * 512 + t * 16;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const uint4 lo = *reinterpret_cast<const uint4*>(xa + s * 64);
            const uint4 hi = *reinterpret_cast<const uint4*>(xa + 8 * xpitch + s * 64);
            mma16816(acc, lo.x, hi.x, lo.y, hi.y, w[s].x, w[s].y);
            mma16816(acc, lo.z, hi.z, lo.w, hi.w, w[s].z, w[s].w);
        }
        if (kc == kch - 1) epi(tile, acc);
    };
    // Three units (= 24 sixteen-byte loads per lane, 96 KB per SM) in flight: with two, every unit waited ~500 cycles for
    // its weights and the stream reached ~20 B/clk/SM (profiles/r1_mega_phases.txt); four spill (255 registers).
    issue(w0, 0);
    if (nunit > 1) issue(w1, 1);
    for (int u = 0; u < nunit; u += 3) {
        if (u + 2 < nunit) issue(w2, u + 2);
        compute(w0, u);
        if (u + 1 < nunit) {
            if (u + 3 < nunit) issue(w0, u + 3);
            compute(w1, u + 1);
        }
        if (u + 2 < nunit) {
            if (u + 4 < nunit) issue(w1, u + 4);
            compute(w2, u + 2);
        }
    }
}

__device__ __forceinline__ void mg_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mg_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct MegaSmem {   // byte offsets into dynamic shared memory (identical in every CTA of the cluster)
    int xs, ctxf, yf, big, qs, kcur, vcur, sc, stats, cand_v, cand_i, misc, total;
    int sc_pitch, hf_pitch, lg_pitch;
    int keys_pad;   // cached positions per warp of the self-attention V staging (multiple of 32)
    int kv_tile;    // bytes of one cross-attention K (or V) tile inside `big`: K at +0 (later the P.V partials), V at +kv_tile
};

__host__ __device__ inline MegaSmem mega_layout(int dff, int V, int T, int Lmax) {
    MegaSmem L;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 127) & ~127; return r; };
    const int ntv = (V + 7) / 8, tpc = (ntv + MG_C - 1) / MG_C;
    L.hf_pitch = dff * 2 + 64;
    L.lg_pitch = tpc * 8;
    int big = 16 * L.hf_pitch;                                         // GLU activations [16][dff] bf16 (peer-written)
    if (16 * L.lg_pitch * 4 > big) big = 16 * L.lg_pitch * 4;          // | local logits slice [16][lg_pitch] f32
    L.keys_pad = (Lmax + 31) / 32 * 32;
    if (8 * L.keys_pad * 128 > big) big = 8 * L.keys_pad * 128;        // | self-attention V rows staged per warp [8][keys_pad][128 B]
    L.kv_tile = T * 128 > 8 * 16 * 64 * 4 ? T * 128 : 8 * 16 * 64 * 4; // | cross-attention K tile (then P.V partials [8][16][64] f32) + V tile
    if (2 * L.kv_tile > big) big = 2 * L.kv_tile;
    const int sc_cols = ((T > MG_LMAX ? T : MG_LMAX) + 31) / 32 * 32 + 1;
    L.sc_pitch = sc_cols;
    L.xs = take(16 * MG_XP);
    L.ctxf = take(16 * MG_XP);
    L.yf = take(16 * MG_D * 4);
    L.big = take(big);
    L.qs = take(16 * 64 * 4);
    L.kcur = take(16 * 64 * 2);
    L.vcur = take(16 * 64 * 2);
    L.sc = take(16 * sc_cols * 4);
    L.stats = take(MG_C * 16 * 8);
    L.cand_v = take(MG_C * 16 * KMAX * 4);
    L.cand_i = take(MG_C * 16 * KMAX * 4);
    L.misc = take(12288);
    L.total = o + 128;
    return L;
}

struct MegaMisc {   // replicated search state + scratch (lives at MegaSmem::misc)
    float scores[16];
    int last_tok[16];
    int flag[16];
    float inv_l[16];
    float lse[16];
    float row_v[16][KMAX];
    int row_i[16][KMAX];
    float c_val[KMAX * KMAX];
    int c_tok[KMAX * KMAX];
    float sel_v[KMAX];
    int sel_i[KMAX];
    unsigned char anc[2][16][MG_LMAX];
};
static_assert(sizeof(MegaMisc) <= 12288, "MegaMisc must fit its smem slot");

__global__ void __cluster_dims__(MG_C, 1, 1) __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const MegaSmem L = mega_layout(p.dff, p.V, p.T, p.st.Lmax);
    uint8_t* xs = smem + L.xs;
    uint8_t* ctxf = smem + L.ctxf;
    float* yf = reinterpret_cast<float*>(smem + L.yf);
    uint8_t* hf = smem + L.big;
    float* lg = reinterpret_cast<float*>(smem + L.big);
    float* red = reinterpret_cast<float*>(smem + L.big);          // aliases the cross-attention K tile once the scores are done
    uint8_t* sKx = smem + L.big;                                  // cross-attention K tile [T][128 B], 16-byte chunks XOR-swizzled by row
    uint8_t* sVx = smem + L.big + L.kv_tile;                      // cross-attention V tile, same layout
    float* qs = reinterpret_cast<float*>(smem + L.qs);
    bf16* kcur = reinterpret_cast<bf16*>(smem + L.kcur);
    bf16* vcur = reinterpret_cast<bf16*>(smem + L.vcur);
    float* sc = reinterpret_cast<float*>(smem + L.sc);
    float2* stats = reinterpret_cast<float2*>(smem + L.stats);
    float* cand_v = reinterpret_cast<float*>(smem + L.cand_v);
    int* cand_i = reinterpret_cast<int*>(smem + L.cand_i);
    MegaMisc& ms = *reinterpret_cast<MegaMisc*>(smem + L.misc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int c = (int)cluster_ctarank();          // == head
    const int b = blockIdx.x / MG_C;               // utterance
    const int beam = p.st.beam, N = p.st.N, Lmax = p.st.Lmax;
    const int n0 = b * beam;
    const int d = MG_D, dff = p.dff, V = p.V;
    const int scp = L.sc_pitch;
    const int kv_len = min(p.mem_len[b], p.T);
    const int ntv = (V + 7) / 8, tpc = (ntv + MG_C - 1) / MG_C;
    const int my_t0 = c * tpc, my_t1 = min(ntv, my_t0 + tpc);      // this CTA's vocabulary tiles
    const int my_ncol = max(0, min(V, my_t1 * 8) - my_t0 * 8);     // valid local vocabulary columns

    // remote (cluster) base addresses of the peer-written buffers
    const uint32_t a_ctxf = smem_u32(ctxf), a_yf = smem_u32(yf), a_hf = smem_u32(hf), a_stats = smem_u32(stats),
                   a_cv = smem_u32(cand_v), a_ci = smem_u32(cand_i);

    // ---- initial search state (speech2text.py:54-58)
    if (tid < 16) {
        ms.scores[tid] = (tid == 0) ? 0.f : -INFINITY;
        ms.last_tok[tid] = (int)EOS_ID;   // BOS == EOS == 1
        ms.flag[tid] = 0;
    }
    for (int i = tid; i < 16 * MG_XP / 4; i += MG_THREADS) {
        reinterpret_cast<uint32_t*>(xs)[i] = 0u;
        reinterpret_cast<uint32_t*>(ctxf)[i] = 0u;
    }
    __syncthreads();
    cluster_sync_all();   // every CTA of the cluster is resident before the first DSMEM store

    int steps_done = 0;
    bool all_ended = false;
    for (int step = 0; step < p.max_steps; ++step) {
        const int cur = step & 1, nxt = cur ^ 1;
        const bool dbg_on = p.dbg_clk != nullptr && blockIdx.x == 0 && tid == 0 && step == p.dbg_step;
        int dbg_n = 0;
        MG_STAMP();
        // ---- S0: embedding + positional encoding (decoder/transformer.py:163,169; pos.py:56) -> xs (bf16, replicated)
        for (int r = warp; r < beam; r += 8) {
            int tok = ms.last_tok[r];
            if (tok < 0 || tok >= V) tok = 0;
            const uint4 u = *reinterpret_cast<const uint4*>(p.emb + (size_t)tok * d + lane * 8);
            const float* pe = p.pe + (size_t)step * d + lane * 8;
            const float4 p0 = *reinterpret_cast<const float4*>(pe), p1 = *reinterpret_cast<const float4*>(pe + 4);
            const float2 e0 = unpack_bf16(u.x), e1 = unpack_bf16(u.y), e2 = unpack_bf16(u.z), e3 = unpack_bf16(u.w);
            const float xsc = 16.0f;   // sqrt(256)
            uint4 o;
            o.x = pack_bf16(e0.x * xsc + p0.x, e0.y * xsc + p0.y);
            o.y = pack_bf16(e1.x * xsc + p0.z, e1.y * xsc + p0.w);
            o.z = pack_bf16(e2.x * xsc + p1.x, e2.y * xsc + p1.y);
            o.w = pack_bf16(e3.x * xsc + p1.z, e3.y * xsc + p1.w);
            *reinterpret_cast<uint4*>(xs + r * MG_XP + lane * 16) = o;
        }
        __syncthreads();
        MG_STAMP();   // 1: S0 embedding

        // LayerNorm(resid + y) computed redundantly by every CTA from the gathered pre-norm rows (transformer.py:54-56 etc.)
        auto layer_norm = [&](const float* gamma, const float* beta) {
            for (int r = warp; r < beam; r += 8) {
                const float4 y0 = *reinterpret_cast<const float4*>(yf + r * d + lane * 8);
                const float4 y1 = *reinterpret_cast<const float4*>(yf + r * d + lane * 8 + 4);
                const uint4 xu = *reinterpret_cast<const uint4*>(xs + r * MG_XP + lane * 16);
                const float2 x0 = unpack_bf16(xu.x), x1 = unpack_bf16(xu.y), x2 = unpack_bf16(xu.z), x3 = unpack_bf16(xu.w);
                float v[8] = {y0.x + x0.x, y0.y + x0.y, y0.z + x1.x, y0.w + x1.y, y1.x + x2.x, y1.y + x2.y, y1.z + x3.x, y1.w + x3.y};
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += v[j];
                const float mean = warp_sum(s) * (1.0f / MG_D);
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float dd = v[j] - mean; q += dd * dd; }
                const float rstd = rsqrtf(warp_sum(q) * (1.0f / MG_D) + p.eps);
                const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(beta + lane * 8), b1 = *reinterpret_cast<const float4*>(beta + lane * 8 + 4);
                uint4 o;
                o.x = pack_bf16((v[0] - mean) * rstd * g0.x + b0.x, (v[1] - mean) * rstd * g0.y + b0.y);
                o.y = pack_bf16((v[2] - mean) * rstd * g0.z + b0.z, (v[3] - mean) * rstd * g0.w + b0.w);
                o.z = pack_bf16((v[4] - mean) * rstd * g1.x + b1.x, (v[5] - mean) * rstd * g1.y + b1.y);
                o.w = pack_bf16((v[6] - mean) * rstd * g1.z + b1.z, (v[7] - mean) * rstd * g1.w + b1.w);
                *reinterpret_cast<uint4*>(xs + r * MG_XP + lane * 16) = o;
            }
        };
        // Y[:, c*64 + warp*8 ..] = X W^T + bias -> fp32 into every CTA's yf (each warp owns one n8 tile of the CTA's 64 columns)
        auto proj_to_yf = [&](const uint8_t* X, int xpitch, int K, const bf16* W, const float* bias) {
            stream_tiles(X, xpitch, K, W, K, d, 1, b, [&](int) { return c * 64 + warp * 8; },
                         [&](int, const float(&acc)[4]) {
                             const int col = c * 64 + warp * 8 + 2 * t;
                             const float b0 = bias[col], b1 = bias[col + 1];
                             const uint32_t o0 = a_yf + (uint32_t)((g * d + col) * 4), o1 = a_yf + (uint32_t)(((g + 8) * d + col) * 4);
#pragma unroll
                             for (int j = 0; j < MG_C; ++j) {
                                 st_cluster_f32x2(mapa_shared(o0, j), acc[0] + b0, acc[1] + b1);
                                 st_cluster_f32x2(mapa_shared(o1, j), acc[2] + b0, acc[3] + b1);
                             }
                         });
        };
        // context of head c (rows < beam, 2 dims per call) -> bf16 into every CTA's ctxf
        auto put_ctx = [&](int r, int dimpair, float v0, float v1) {
            const uint32_t o = a_ctxf + (uint32_t)(r * MG_XP + (c * 64 + 2 * dimpair) * 2);
            const uint32_t u = pack_bf16(v0, v1);
#pragma unroll
            for (int j = 0; j < MG_C; ++j) st_cluster_u32(mapa_shared(o, j), u);
        };

        for (int l = 0; l < p.n_layers; ++l) {
            const MegaLayer& ly = p.layers[l];
            // ---- S1: Q,K,V of head c (attention.py:68-73): warp w -> tile w of Q, of K and of V
            stream_tiles(xs, MG_XP, d, ly.wqkv, d, 3 * d, 3, b, [&](int i) { return i * d + c * 64 + warp * 8; },
                         [&](int i, const float(&acc)[4]) {
                             const int col = warp * 8 + 2 * t;                 // column inside the head
                             const float b0 = ly.bqkv[i * d + c * 64 + col], b1 = ly.bqkv[i * d + c * 64 + col + 1];
                             if (i == 0) {
                                 *reinterpret_cast<float2*>(qs + g * 64 + col) = make_float2(acc[0] + b0, acc[1] + b1);
                                 *reinterpret_cast<float2*>(qs + (g + 8) * 64 + col) = make_float2(acc[2] + b0, acc[3] + b1);
                             } else {
                                 bf16* cs = (i == 1) ? kcur : vcur;
                                 bf16* cg = ((i == 1) ? p.kc : p.vc) + (((size_t)l * Lmax + step) * N + n0) * d + c * 64 + col;
                                 const uint32_t u0 = pack_bf16(acc[0] + b0, acc[1] + b1), u1 = pack_bf16(acc[2] + b0, acc[3] + b1);
                                 *reinterpret_cast<uint32_t*>(cs + g * 64 + col) = u0;
                                 *reinterpret_cast<uint32_t*>(cs + (g + 8) * 64 + col) = u1;
                                 if (g < beam) *reinterpret_cast<uint32_t*>(cg + (size_t)g * d) = u0;
                                 if (g + 8 < beam) *reinterpret_cast<uint32_t*>(cg + (size_t)(g + 8) * d) = u1;
                             }
                         });
            __syncthreads();
            MG_STAMP();   // S1 qkv
            // ---- S2: self-attention of head c over the cached prefix (the cache the reference stubbed out,
            //      decoder/transformer.py:92-126); warp = hypothesis row.  Lanes run over KEYS: a lane resolves its cached
            //      positions through the ancestry table and pulls the whole 128-byte K and V rows with sixteen independent
            //      16-byte loads, V rows are staged in the warp's slice of `big` so that P.V is a column sum over shared
            //      memory -- one or two L2 round trips per row instead of one per key.
            {
                const int nkeys = step + 1;
                uint8_t* stage = hf + (size_t)warp * L.keys_pad * 128;
                for (int r = warp; r < beam; r += 8) {
                    const unsigned char* an = ms.anc[cur][r];
                    float mx = -INFINITY;
                    for (int s = lane; s < nkeys; s += 32) {
                        const bf16* krow = (s == step) ? (kcur + r * 64) : (p.kc + (((size_t)l * Lmax + s) * N + n0 + an[s]) * d + c * 64);
                        const bf16* vrow = (s == step) ? (vcur + r * 64) : (p.vc + (((size_t)l * Lmax + s) * N + n0 + an[s]) * d + c * 64);
                        uint4 ku[8], vu[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) ku[i] = *reinterpret_cast<const uint4*>(krow + 8 * i);
#pragma unroll
                        for (int i = 0; i < 8; ++i) vu[i] = *reinterpret_cast<const uint4*>(vrow + 8 * i);
                        float dot = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 k0 = unpack_bf16(ku[i].x), k1 = unpack_bf16(ku[i].y), k2 = unpack_bf16(ku[i].z), k3 = unpack_bf16(ku[i].w);
                            const float4 q0 = *reinterpret_cast<const float4*>(qs + r * 64 + 8 * i);
                            const float4 q1 = *reinterpret_cast<const float4*>(qs + r * 64 + 8 * i + 4);
                            dot += q0.x * k0.x + q0.y * k0.y + q0.z * k1.x + q0.w * k1.y + q1.x * k2.x + q1.y * k2.y + q1.z * k3.x + q1.w * k3.y;
                        }
                        dot *= 0.125f;
                        sc[r * scp + s] = dot;
                        mx = fmaxf(mx, dot);
#pragma unroll
                        for (int i = 0; i < 8; ++i)      // chunks rotated by the row index: conflict-free 16-byte stores
                            *reinterpret_cast<uint4*>(stage + (size_t)s * 128 + (((i + s) & 7) << 4)) = vu[i];
                    }
                    mx = warp_max(mx);
                    float lsum = 0.f;
                    for (int s = lane; s < nkeys; s += 32) {
                        const float e = __expf(sc[r * scp + s] - mx);
                        sc[r * scp + s] = e;
                        lsum += e;
                    }
                    lsum = warp_sum(lsum);
                    __syncwarp();
                    float ax = 0.f, ay = 0.f;
#pragma unroll 4
                    for (int s = 0; s < nkeys; ++s) {
                        // dims 2*lane.. live in chunk lane/4 (rotated by s), word lane%4
                        const float2 vv = unpack_bf16(*reinterpret_cast<const uint32_t*>(stage + (size_t)s * 128 + ((((lane >> 2) + s) & 7) << 4) + (lane & 3) * 4));
                        const float pw = sc[r * scp + s];
                        ax = fmaf(pw, vv.x, ax);
                        ay = fmaf(pw, vv.y, ay);
                    }
                    const float inv = 1.0f / lsum;
                    put_ctx(r, lane, ax * inv, ay * inv);
                    __syncwarp();      // the next row of this warp re-uses the staging slice
                }
            }
            __syncthreads();           // every warp is done with its staging slice: `big` may receive the cross-attention tiles
            {   // cross-attention K / V tiles of (layer l, utterance b, head c) -> shared memory, asynchronously: they arrive
                // while the out-projection, LayerNorm and the query projection run (S3, LN1, S4)
                const bf16* kb = p.kvx + ((size_t)l * p.B * p.T + (size_t)b * p.T) * (2 * d) + c * 64;
                for (int i = tid; i < kv_len * 8; i += MG_THREADS) {
                    const int j = i >> 3, ch = i & 7;
                    const int dst = j * 128 + ((ch ^ (j & 7)) << 4);
                    mg_cp_async16(sKx + dst, kb + (size_t)j * (2 * d) + ch * 8);
                    mg_cp_async16(sVx + dst, kb + (size_t)j * (2 * d) + d + ch * 8);
                }
            }
            MG_STAMP();   // S2 self-attention
            cluster_sync_all();                                                         // #1 context gathered
            MG_STAMP();   // barrier 1
            proj_to_yf(ctxf, MG_XP, d, ly.wo, ly.bo);                                   // S3 (attention.py:44)
            MG_STAMP();   // S3 out-proj
            cluster_sync_all();                                                         // #2 pre-norm rows gathered
            MG_STAMP();   // barrier 2
            layer_norm(ly.g1, ly.be1);
            __syncthreads();
            MG_STAMP();   // LN1
            // ---- S4: cross-attention query of head c (attention.py:128)
            stream_tiles(xs, MG_XP, d, ly.wq, d, d, 1, 0, [&](int) { return c * 64 + warp * 8; },
                         [&](int, const float(&acc)[4]) {
                             const int col = warp * 8 + 2 * t;
                             const float b0 = ly.bq[c * 64 + col], b1 = ly.bq[c * 64 + col + 1];
                             *reinterpret_cast<float2*>(qs + g * 64 + col) = make_float2(acc[0] + b0, acc[1] + b1);
                             *reinterpret_cast<float2*>(qs + (g + 8) * 64 + col) = make_float2(acc[2] + b0, acc[3] + b1);
                         });
            __syncthreads();
            MG_STAMP();   // S4 q-proj
            // ---- S5: cross-attention of head c over the utterance's memory (attention.py:129-141,34-41), K / V tiles in
            //      shared memory (staged above): thread per key for the scores, warp per row for the soft-max, warps over
            //      key groups for P.V
            {
                mg_cp_async_wait_all();
                __syncthreads();
                for (int j = tid; j < kv_len; j += MG_THREADS) {
                    const uint8_t* krow = sKx + j * 128;
                    const int sw = j & 7;
                    for (int r = 0; r < beam; ++r) {          // rows outer: few live registers, the K row is re-read from smem
                        float dot = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint4 u = *reinterpret_cast<const uint4*>(krow + ((i ^ sw) << 4));
                            const float2 k0 = unpack_bf16(u.x), k1 = unpack_bf16(u.y), k2 = unpack_bf16(u.z), k3 = unpack_bf16(u.w);
                            const float4 q0 = *reinterpret_cast<const float4*>(qs + r * 64 + 8 * i);       // broadcast
                            const float4 q1 = *reinterpret_cast<const float4*>(qs + r * 64 + 8 * i + 4);
                            dot += q0.x * k0.x + q0.y * k0.y + q0.z * k1.x + q0.w * k1.y + q1.x * k2.x + q1.y * k2.y + q1.z * k3.x + q1.w * k3.y;
                        }
                        sc[r * scp + j] = dot * 0.125f;
                    }
                }
                __syncthreads();
                MG_STAMP();   // S5 scores
                for (int r = warp; r < beam; r += 8) {                 // soft-max statistics, warp per row
                    float mx = -INFINITY;
                    for (int j = lane; j < kv_len; j += 32) mx = fmaxf(mx, sc[r * scp + j]);
                    mx = warp_max(mx);
                    float lsum = 0.f;
                    for (int j = lane; j < kv_len; j += 32) {
                        const float e = __expf(sc[r * scp + j] - mx);
                        sc[r * scp + j] = e;
                        lsum += e;
                    }
                    lsum = warp_sum(lsum);
                    if (lane == 0) ms.inv_l[r] = (lsum > 0.f) ? 1.0f / lsum : 0.f;
                }
                __syncthreads();
                MG_STAMP();   // S5 softmax
                float ax[16], ay[16];                                   // P.V: warp = key group, lane = dim pair
#pragma unroll
                for (int r = 0; r < 16; ++r) { ax[r] = 0.f; ay[r] = 0.f; }
                for (int j = warp; j < kv_len; j += 8) {
                    const float2 vv = unpack_bf16(*reinterpret_cast<const uint32_t*>(sVx + j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4));
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        if (r < beam) {
                            const float pw = sc[r * scp + j];
                            ax[r] = fmaf(pw, vv.x, ax[r]);
                            ay[r] = fmaf(pw, vv.y, ay[r]);
                        }
                    }
                }
                // the K tile is dead since the scores: its region now holds the partials of the 8 key groups
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (r < beam) *reinterpret_cast<float2*>(red + (warp * 16 + r) * 64 + 2 * lane) = make_float2(ax[r], ay[r]);
                __syncthreads();
                for (int idx = tid; idx < beam * 32; idx += MG_THREADS) {
                    const int r = idx >> 5, dp = idx & 31;
                    float sx = 0.f, sy = 0.f;
#pragma unroll
                    for (int kg = 0; kg < 8; ++kg) {
                        const float2 v2 = *reinterpret_cast<const float2*>(red + (kg * 16 + r) * 64 + 2 * dp);
                        sx += v2.x;
                        sy += v2.y;
                    }
                    put_ctx(r, dp, sx * ms.inv_l[r], sy * ms.inv_l[r]);
                }
            }
            MG_STAMP();   // S5 PV
            cluster_sync_all();                                                         // #3
            MG_STAMP();   // barrier 3
            proj_to_yf(ctxf, MG_XP, d, ly.wo2, ly.bo2);                                 // S6
            MG_STAMP();   // S6 out-proj2
            cluster_sync_all();                                                         // #4
            MG_STAMP();   // barrier 4
            layer_norm(ly.g2, ly.be2);
            __syncthreads();
            MG_STAMP();   // LN2
            // ---- S7: GLU feed-forward, hidden features [c*dff/4, (c+1)*dff/4) (ffn.py:18,39): value tile then gate tile
            {
                const int fpc = dff / MG_C;                 // hidden features per CTA
                const int pairs = fpc / 8 / 8;              // (value, gate) tile pairs per warp
                float va[4] = {0.f, 0.f, 0.f, 0.f};
                stream_tiles(xs, MG_XP, d, ly.w1, d, 2 * dff, 2 * pairs, 2 * (b % pairs),
                             [&](int i) { return ((i & 1) ? dff : 0) + c * fpc + (warp + 8 * (i >> 1)) * 8; },
                             [&](int i, const float(&acc)[4]) {
                                 const int feat = c * fpc + (warp + 8 * (i >> 1)) * 8 + 2 * t;
                                 if (!(i & 1)) {
                                     const float b0 = ly.b1[feat], b1 = ly.b1[feat + 1];
                                     va[0] = acc[0] + b0; va[1] = acc[1] + b1; va[2] = acc[2] + b0; va[3] = acc[3] + b1;
                                 } else {
                                     const float b0 = ly.b1[dff + feat], b1 = ly.b1[dff + feat + 1];
                                     const uint32_t u0 = pack_bf16(va[0] * mg_sigmoid(acc[0] + b0), va[1] * mg_sigmoid(acc[1] + b1));
                                     const uint32_t u1 = pack_bf16(va[2] * mg_sigmoid(acc[2] + b0), va[3] * mg_sigmoid(acc[3] + b1));
                                     const uint32_t o0 = a_hf + (uint32_t)(g * L.hf_pitch + feat * 2);
                                     const uint32_t o1 = a_hf + (uint32_t)((g + 8) * L.hf_pitch + feat * 2);
#pragma unroll
                                     for (int j = 0; j < MG_C; ++j) {
                                         st_cluster_u32(mapa_shared(o0, j), u0);
                                         st_cluster_u32(mapa_shared(o1, j), u1);
                                     }
                                 }
                             });
            }
            MG_STAMP();   // S7 GLU
            cluster_sync_all();                                                         // #5 hidden activations gathered
            MG_STAMP();   // barrier 5
            proj_to_yf(hf, L.hf_pitch, dff, ly.w2, ly.b2);                              // S8
            MG_STAMP();   // S8 w2
            cluster_sync_all();                                                         // #6
            MG_STAMP();   // barrier 6
            layer_norm(ly.g3, ly.be3);
            __syncthreads();
            MG_STAMP();   // LN3
        }

        // ---- S9: logits of this CTA's vocabulary slice (decoder/transformer.py:181) -> lg (fp32, local)
        {
            const int mine = (my_t1 - my_t0 - warp + 7) / 8;      // tiles warp, warp+8, ... of the slice
            stream_tiles(xs, MG_XP, d, p.wout, d, V, mine > 0 ? mine : 0, b, [&](int i) { return (my_t0 + warp + 8 * i) * 8; },
                         [&](int i, const float(&acc)[4]) {
                             const int lc = (warp + 8 * i) * 8 + 2 * t;
                             const int gc = my_t0 * 8 + lc;
                             const float b0 = (gc < V && p.bout) ? p.bout[gc] : 0.f, b1 = (gc + 1 < V && p.bout) ? p.bout[gc + 1] : 0.f;
                             lg[g * L.lg_pitch + lc] = (gc < V) ? acc[0] + b0 : -INFINITY;
                             lg[g * L.lg_pitch + lc + 1] = (gc + 1 < V) ? acc[1] + b1 : -INFINITY;
                             lg[(g + 8) * L.lg_pitch + lc] = (gc < V) ? acc[2] + b0 : -INFINITY;
                             lg[(g + 8) * L.lg_pitch + lc + 1] = (gc + 1 < V) ? acc[3] + b1 : -INFINITY;
                         });
        }
        __syncthreads();
        MG_STAMP();   // S9 logits
        // partial log-sum-exp of the slice -> every CTA (decoder/transformer.py:206)
        for (int r = warp; r < beam; r += 8) {
            const float* row = lg + r * L.lg_pitch;
            float mx = -INFINITY;
            for (int j = lane; j < my_ncol; j += 32) mx = fmaxf(mx, row[j]);
            mx = warp_max(mx);
            float s = 0.f;
            for (int j = lane; j < my_ncol; j += 32) s += expf(row[j] - mx);
            s = warp_sum(s);
            if (lane < MG_C) st_cluster_f32x2(mapa_shared(a_stats + (uint32_t)((c * 16 + r) * 8), lane), mx, s);
        }
        MG_STAMP();   // partial lse
        cluster_sync_all();                                                             // #7
        MG_STAMP();   // barrier 7
        for (int r = warp; r < beam; r += 8) {
            float M = -INFINITY;
#pragma unroll
            for (int j = 0; j < MG_C; ++j) M = fmaxf(M, stats[j * 16 + r].x);
            float S = 0.f;
#pragma unroll
            for (int j = 0; j < MG_C; ++j) S += stats[j * 16 + r].y * expf(stats[j * 16 + r].x - M);
            const float lse = M + logf(S);
            float* row = lg + r * L.lg_pitch;
            float* dump = p.dbg_logp ? p.dbg_logp + ((size_t)step * N + n0 + r) * V + my_t0 * 8 : nullptr;
            for (int j = lane; j < my_ncol; j += 32) {
                const float v = row[j] - lse;
                row[j] = v;
                if (dump) dump[j] = v;
            }
            __syncwarp();
            // local top-k (speech2text.py:112): k rounds of (scan, warp arg-max, knock out), ties -> lower index
            for (int k = 0; k < beam; ++k) {
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int j = lane; j < my_ncol; j += 32) {
                    const float v = row[j];
                    if (v > bv) { bv = v; bi = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                if (bi != 0x7fffffff && (bi & 31) == lane) row[bi] = -INFINITY;
                if (lane < MG_C) {
                    const uint32_t off = (uint32_t)(((c * 16 + r) * KMAX + k) * 4);
                    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(mapa_shared(a_cv + off, lane)), "f"(bv) : "memory");
                    st_cluster_u32(mapa_shared(a_ci + off, lane), (uint32_t)(bi == 0x7fffffff ? bi : bi + my_t0 * 8));
                }
                __syncwarp();
            }
        }
        MG_STAMP();   // log-probs + local top-k
        cluster_sync_all();                                                             // #8 candidates gathered
        MG_STAMP();   // barrier 8
        // ---- merge the 4 sorted candidate lists per row, then the beam step (redundantly in every CTA; beam.cu semantics)
        for (int r = warp; r < beam; r += 8) {
            if (ms.flag[r]) {                       // mask_finished_scores / mask_finished_preds (speech2text.py:156-192)
                if (lane < beam) {
                    ms.row_v[r][lane] = (lane == 0) ? 0.f : -INFINITY;
                    ms.row_i[r][lane] = (int)EOS_ID;
                }
            } else {
                int pos = 0;
                for (int k = 0; k < beam; ++k) {
                    float bv = (lane < MG_C && pos < beam) ? cand_v[(lane * 16 + r) * KMAX + pos] : -INFINITY;
                    int bi = (lane < MG_C && pos < beam) ? cand_i[(lane * 16 + r) * KMAX + pos] : 0x7fffffff;
                    const float mv = bv;
                    const int mi = bi;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                    }
                    if (mi == bi && mv == bv && lane < MG_C) ++pos;
                    if (lane == 0) { ms.row_v[r][k] = bv; ms.row_i[r][k] = bi; }
                }
            }
            __syncwarp();
            const float base = ms.scores[r];
            if (lane < beam) {
                ms.c_val[r * beam + lane] = base + ms.row_v[r][lane];      // scores + last_k_scores (:118)
                ms.c_tok[r * beam + lane] = ms.row_i[r][lane];
            }
        }
        __syncthreads();
        if (warp == 0) warp_topk([&](int idx) { return ms.c_val[idx]; }, beam * beam, beam, ms.sel_v, ms.sel_i);   // (:119-122)
        __syncthreads();
        for (int r = warp; r < beam; r += 8) {      // ancestry of the surviving hypotheses (:126-140)
            const int parent = ms.sel_i[r] / beam;
            for (int s = lane; s < step; s += 32) ms.anc[nxt][r][s] = ms.anc[cur][parent][s];
            if (lane == 0) ms.anc[nxt][r][step] = (unsigned char)parent;
        }
        __syncthreads();
        int ended = 0;
        if (tid < beam) {
            const int off = ms.sel_i[tid];
            const int tok = ms.c_tok[off];
            ms.scores[tid] = ms.sel_v[tid];
            ms.last_tok[tid] = tok;
            ms.flag[tid] = (tok == (int)EOS_ID) ? 1 : 0;
            ended = (tok == (int)EOS_ID) ? 1 : 0;
            if (c == 0) {
                p.st.tok_hist[(size_t)step * N + n0 + tid] = tok;
                p.st.par_hist[(size_t)step * N + n0 + tid] = n0 + off / beam;
                if (p.dbg_scores) p.dbg_scores[(size_t)step * N + n0 + tid] = ms.sel_v[tid];
            }
        }
        MG_STAMP();   // merge + beam step
        steps_done = step + 1;
        all_ended = __syncthreads_and(tid >= beam || ended) != 0;
        if (all_ended) break;       // every hypothesis of this utterance ended (uniform across the cluster: replicated state)
    }

    // An utterance that ended early keeps emitting EOS from sorted hypotheses (identity parents) while the reference
    // loops on for the other utterances (speech2text.py:62-68): fill the rest of its history so that any global step
    // count >= its own gives the reference's tokens.
    if (c == 0) {
        for (int i = tid; i < (p.max_steps - steps_done) * beam; i += MG_THREADS) {
            const int s = steps_done + i / beam, r = i % beam;
            p.st.tok_hist[(size_t)s * N + n0 + r] = (int)EOS_ID;
            p.st.par_hist[(size_t)s * N + n0 + r] = n0 + r;
            if (p.dbg_scores) p.dbg_scores[(size_t)s * N + n0 + r] = ms.scores[r];
        }
        if (tid < beam) {
            p.st.scores[n0 + tid] = ms.scores[tid];
            p.st.flag[n0 + tid] = (unsigned char)ms.flag[tid];
            p.st.last_tok[n0 + tid] = ms.last_tok[tid];
        }
        if (tid == 0) {
            atomicMax(&p.st.ctrl[0], all_ended ? steps_done : p.max_steps);   // global step count = max over utterances
            if (!all_ended) atomicAdd(&p.st.ctrl[2], 1);                       // utterances that hit max_len unfinished
        }
    }
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

const char* decode_mega_launch(cudaStream_t st, const MegaParams& p) {
    if (p.d != MG_D || p.H != MG_C) return "decode_mega: needs d_model 256 with 4 heads";
    if (p.dff < 512 || p.dff % 512) return "decode_mega: d_ff must be a positive multiple of 512";
    if (p.st.beam < 1 || p.st.beam > KMAX) return "decode_mega: beam must be in [1,16]";
    if (p.st.Lmax > MG_LMAX || p.max_steps > p.st.Lmax || p.max_steps < 1) return "decode_mega: max_steps <= Lmax <= 128";
    if (p.n_layers < 1 || p.n_layers > OTB_MEGA_MAX_LAYERS_INT) return "decode_mega: too many layers";
    if (p.B < 1 || p.T < 1 || p.st.N != p.B * p.st.beam) return "decode_mega: bad batch geometry";
    if (p.V < 8) return "decode_mega: vocabulary too small";
    const MegaSmem L = mega_layout(p.dff, p.V, p.T, p.st.Lmax);
    if (L.total > 227 * 1024) return "decode_mega: shared-memory budget exceeded (memory too long / d_ff too large)";
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return "cudaFuncSetAttribute(decode_mega) failed";
        attr_set = true;
    }
    MegaParams q = p;
    q.dbg_clk = g_mega_dbg;
    q.dbg_step = g_mega_dbg_step;
    decode_mega_kernel<<<p.B * MG_C, MG_THREADS, L.total, st>>>(q);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace otb

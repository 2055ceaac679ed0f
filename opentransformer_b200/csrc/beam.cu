// Batched beam-search step and finalisation on the device.
// Replaces the ~10 ATen launches + host sync per step of SpeechToTextRecognizer.decode_step
//   otrans/recognize/speech2text.py:102-153 (LM add, top-k per hypothesis, finished masking :156-192,
//   score add, top-k over beam^2, gather tokens/parents, EOS flags)
// and the tail of recognize() (:70-91: lengths, length penalty applied once, sort, n-best gather).
//
// Integer outputs (token ids, parent rows, top-k offsets) are bit-exact with the reference on
// tie-free inputs; ties are broken towards the LOWER index (torch.topk's tie order is
// implementation-defined).  Scores are fp32 and use the same additions as the reference.
//
// Device-resident search state (N = B * beam hypotheses):
//   tok_hist  i32 [Lmax, N]   token chosen at step s for (post-reorder) hypothesis n
//   par_hist  i32 [Lmax, N]   its parent row (global, pre-reorder) at step s
//   last_tok  i64 [N]         newest token of every hypothesis (decoder input of the next step)
//   scores    f32 [N], flag u8 [N]
//   anc       i32 [2, N, Lmax] ping-pong ancestry table used by the KV-cache attention
//   ctrl      i32 [4]: {step, done, ended_now, unused}
#include <math.h>

#include "beam_common.cuh"
#include "launch.cuh"
#include "otb_internal.h"
#include "ptx.cuh"

namespace otb {

// one CTA per utterance, `beam` warps
__global__ void __launch_bounds__(KMAX * 32) beam_step_kernel(const float* __restrict__ logp, int ldl, int V,
                                                              const float* __restrict__ lm_logp, int ld_lm,
                                                              float lm_weight, BeamState st, long long* dbg_ktok,
                                                              int* dbg_offs, const float* __restrict__ pre_val,
                                                              const int* __restrict__ pre_idx) {
    PDL_TRIGGER();
    PDL_WAIT();
    __shared__ float c_val[KMAX * KMAX];
    __shared__ int c_tok[KMAX * KMAX];
    __shared__ float sel_v[KMAX];
    __shared__ int sel_i[KMAX];
    __shared__ float row_v[KMAX][KMAX];
    __shared__ int row_i[KMAX][KMAX];
    const int beam = st.beam;
    const int u = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (st.ctrl[1] || st.ctrl[0] >= st.Lmax) return;  // search already ended: state stays frozen (reference breaks out, speech2text.py:67)
    const int step = st.ctrl[0];

    // ---- phase 1: per-hypothesis top-k over the vocabulary (speech2text.py:112), finished masking (:114-115)
    if (warp < beam) {
        const int n = u * beam + warp;
        const bool fin = st.flag[n] != 0;
        if (fin) {
            if (lane < beam) {
                row_v[warp][lane] = (lane == 0) ? 0.f : -INFINITY;
                row_i[warp][lane] = (int)EOS_ID;
            }
        } else if (pre_val != nullptr) {
            if (lane < beam) {
                row_v[warp][lane] = pre_val[(size_t)n * beam + lane];
                row_i[warp][lane] = pre_idx[(size_t)n * beam + lane];
            }
        } else {
            const float* lp = logp + (size_t)n * ldl;
            const float* lm = lm_logp ? lm_logp + (size_t)n * ld_lm : nullptr;
            warp_topk([&](int idx) { return lm ? lp[idx] + lm_weight * lm[idx] : lp[idx]; }, V, beam, row_v[warp],
                      row_i[warp]);
        }
        __syncwarp();
        const float base = st.scores[n];
        if (lane < beam) {
            c_val[warp * beam + lane] = base + row_v[warp][lane];  // scores + last_k_scores (:118)
            c_tok[warp * beam + lane] = row_i[warp][lane];
            if (dbg_ktok) dbg_ktok[(size_t)n * beam + lane] = row_i[warp][lane];
        }
    }
    __syncthreads();
    // ---- phase 2: top-k over the beam*beam candidates of this utterance (:119-122)
    if (warp == 0) {
        warp_topk([&](int idx) { return c_val[idx]; }, beam * beam, beam, sel_v, sel_i);
    }
    __syncthreads();
    // ---- phase 3: gather tokens / parents, update state (:126-146)
    const int cur = step & 1, nxt = cur ^ 1;
    int ended = 0;
    for (int r = warp; r < beam; r += (blockDim.x >> 5)) {
        const int off = sel_i[r];
        const int parent = u * beam + off / beam;
        const int tok = c_tok[off];
        const int nn = u * beam + r;
        const int* a_old = st.anc + ((size_t)cur * st.N + parent) * st.Lmax;
        int* a_new = st.anc + ((size_t)nxt * st.N + nn) * st.Lmax;
        for (int s = lane; s < step; s += 32) a_new[s] = a_old[s];
        if (lane == 0) {
            a_new[step] = parent;
            st.tok_hist[(size_t)step * st.N + nn] = tok;
            st.par_hist[(size_t)step * st.N + nn] = parent;
            if (dbg_offs) dbg_offs[nn] = off;
        }
    }
    __syncthreads();  // every read of the old scores / flags of this utterance happened in phase 1
    if (threadIdx.x < beam) {
        const int r = threadIdx.x;
        const int nn = u * beam + r;
        const int tok = c_tok[sel_i[r]];
        st.scores[nn] = sel_v[r];
        st.last_tok[nn] = tok;
        st.flag[nn] = (tok == (int)EOS_ID) ? 1 : 0;
        ended = (tok == (int)EOS_ID) ? 1 : 0;
    }
    if (threadIdx.x < 32) {
        ended = (int)warp_sum((float)ended);
        if (threadIdx.x == 0) {
            if (ended) atomicAdd(&st.ctrl[2], ended);
            __threadfence();
            // last CTA of this step advances {step, done}: replaces a separate 1-thread kernel per step
            if (atomicAdd(&st.ctrl[3], 1) == (int)gridDim.x - 1) {
                const int total = atomicAdd(&st.ctrl[2], 0);
                st.ctrl[2] = 0;
                st.ctrl[3] = 0;
                __threadfence();
                if (total == st.N) st.ctrl[1] = 1;      // every hypothesis ended (speech2text.py:66-67)
                st.ctrl[0] = step + 1;
            }
        }
    }
}

// single thread: step += 1; done when every hypothesis ended at this step (speech2text.py:66-67)
__global__ void beam_advance_kernel(int* ctrl, int N, int Lmax) {
    if (ctrl[1] || ctrl[0] >= Lmax) return;
    const int ended = ctrl[2];
    ctrl[2] = 0;
    ctrl[0] += 1;
    if (ended == N) ctrl[1] = 1;
}

const char* beam_step_launch(cudaStream_t stream, const float* logp, int ldl, int V, const float* lm_logp, int ld_lm,
                             float lm_weight, BeamState st, long long* dbg_ktok, int* dbg_offs, const float* pre_val,
                             const int* pre_idx) {
    if (st.beam < 1 || st.beam > KMAX) return "beam_step: beam must be in [1,16]";
    if (st.N % st.beam) return "beam_step: N must be a multiple of beam";
    if (!logp && !pre_val) return "beam_step: need log-probs or a precomputed top-k";
    if ((pre_val == nullptr) != (pre_idx == nullptr)) return "beam_step: pre_val / pre_idx must both be given";
    cudaError_t e = launch_pdl(beam_step_kernel, dim3(st.N / st.beam), dim3(st.beam * 32), 0, stream, logp, ldl, V, lm_logp, ld_lm,
                               lm_weight, st, dbg_ktok, dbg_offs, pre_val, pre_idx);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// ------------------------------------------------------------------------------------------------
// preds reconstruction: walk the (token, parent) back-pointers.  preds [N, ld] i64, column 0 = BOS.
// ------------------------------------------------------------------------------------------------
__global__ void beam_reconstruct_kernel(BeamState st, long long* preds, int ld, int steps) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= st.N) return;
    long long* row = preds + (size_t)n * ld;
    row[0] = 1;  // BOS
    int cur = n;
    for (int s = steps - 1; s >= 0; --s) {
        row[s + 1] = st.tok_hist[(size_t)s * st.N + cur];
        cur = st.par_hist[(size_t)s * st.N + cur];
    }
}

const char* beam_reconstruct_launch(cudaStream_t stream, BeamState st, long long* preds, int ld, int steps) {
    if (steps < 0 || steps > st.Lmax || ld < steps + 1) return "beam_reconstruct: bad steps / ld";
    beam_reconstruct_kernel<<<(st.N + 127) / 128, 128, 0, stream>>>(st, preds, ld, steps);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// ------------------------------------------------------------------------------------------------
// finalisation (speech2text.py:70-91): lengths = #(tok != EOS) (BOS == EOS so BOS is not counted),
// scores /= ((lamda + len)/(lamda + 1))^penalty, sort descending, gather, strip BOS, take nbest.
// One CTA per utterance; thread r < beam handles hypothesis r.  steps is read from ctrl[0].
//   out_preds i64 [B, nbest, Lmax]  (columns >= steps are filled with EOS), out_scores f32 [B, nbest]
// ------------------------------------------------------------------------------------------------
__global__ void beam_finalize_kernel(BeamState st, float penalty, float lamda, int nbest, long long* out_preds,
                                     float* out_scores) {
    __shared__ float s_score[KMAX];
    __shared__ int s_rank[KMAX];
    const int u = blockIdx.x, r = threadIdx.x, beam = st.beam;
    const int steps = st.ctrl[0];
    const int n = u * beam + r;
    float sc = -INFINITY;
    if (r < beam) {
        int cur = n, len = 0;
        for (int s = steps - 1; s >= 0; --s) {
            len += (st.tok_hist[(size_t)s * st.N + cur] != (int)EOS_ID) ? 1 : 0;
            cur = st.par_hist[(size_t)s * st.N + cur];
        }
        sc = st.scores[n];
        if (penalty != 0.f) sc = sc / powf((lamda + (float)len) / (lamda + 1.0f), penalty);
        s_score[r] = sc;
    }
    __syncthreads();
    if (r < beam) {
        int rank = 0;
        for (int j = 0; j < beam; ++j)
            if (better(s_score[j], j, sc, r)) ++rank;
        s_rank[rank] = r;
    }
    __syncthreads();
    if (r < nbest && r < beam) {
        const int src = u * beam + s_rank[r];
        out_scores[(size_t)u * nbest + r] = s_score[s_rank[r]];
        long long* row = out_preds + ((size_t)u * nbest + r) * st.Lmax;
        for (int s = steps; s < st.Lmax; ++s) row[s] = EOS_ID;
        int cur = src;
        for (int s = steps - 1; s >= 0; --s) {
            row[s] = st.tok_hist[(size_t)s * st.N + cur];
            cur = st.par_hist[(size_t)s * st.N + cur];
        }
    }
}

const char* beam_finalize_launch(cudaStream_t stream, BeamState st, float penalty, float lamda, int nbest,
                                 long long* out_preds, float* out_scores) {
    if (nbest < 1) return "beam_finalize: nbest < 1";
    beam_finalize_kernel<<<st.N / st.beam, 32, 0, stream>>>(st, penalty, lamda, nbest, out_preds, out_scores);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// initial state (speech2text.py:54-58): preds = BOS, scores = [0, -inf, ...] per utterance, flags = 0
__global__ void beam_init_kernel(BeamState st) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 4) st.ctrl[n] = 0;
    if (n >= st.N) return;
    st.last_tok[n] = 1;
    st.scores[n] = (n % st.beam == 0) ? 0.f : -INFINITY;
    st.flag[n] = 0;
}

const char* beam_init_launch(cudaStream_t stream, BeamState st) {
    int n = st.N < 4 ? 4 : st.N;
    beam_init_kernel<<<(n + 127) / 128, 128, 0, stream>>>(st);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// ------------------------------------------------------------------------------------------------
// Fused  log_softmax(logits[row]) (+ lm_weight * lm_logp[row])  ->  top-k (values, vocabulary ids)
// decoder/transformer.py:206 + recognize/speech2text.py:102-112 in ONE pass over the logits: the row is
// staged in smem once (coalesced), so the [N,V] log-prob matrix is never written or re-read.  Values are
// formed exactly as the stand-alone log-softmax kernel does (x - (max + log sum exp(x - max))) before the
// top-k, so the selected ids equal top-k over the materialised log-probs (ties -> lower index).
// One CTA per hypothesis row, 8 warps; per-warp top-k over a slice, then one warp merges 8*k candidates.
// ------------------------------------------------------------------------------------------------
static constexpr int TOPK_THREADS = 256;

__global__ void __launch_bounds__(TOPK_THREADS) logsoftmax_topk_kernel(const float* __restrict__ logits, int ldl, int V,
                                                                       const float* __restrict__ lm_logp, int ld_lm,
                                                                       float lm_weight, int k, float* __restrict__ out_val,
                                                                       int* __restrict__ out_idx,
                                                                       float* __restrict__ out_logp, int ld_logp) {
    PDL_TRIGGER();
    PDL_WAIT();
    extern __shared__ float srow[];   // [V]
    __shared__ float red[8];
    __shared__ float bc;
    __shared__ float cand_v[8 * KMAX];
    __shared__ int cand_i[8 * KMAX];
    const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* x = logits + (size_t)row * ldl;
    float m = -INFINITY;
    for (int i = tid; i < V; i += TOPK_THREADS) {
        const float v = x[i];
        srow[i] = v;
        m = fmaxf(m, v);
    }
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        float t = red[0];
        for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
        bc = t;
    }
    __syncthreads();
    m = bc;
    float sum = 0.f;
    for (int i = tid; i < V; i += TOPK_THREADS) sum += expf(srow[i] - m);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bc = m + logf(t);
    }
    __syncthreads();
    const float lse = bc;
    const float* lm = lm_logp ? lm_logp + (size_t)row * ld_lm : nullptr;
    for (int i = tid; i < V; i += TOPK_THREADS) {
        float v = srow[i] - lse;
        if (lm) v = v + lm_weight * lm[i];
        srow[i] = v;
        if (out_logp) out_logp[(size_t)row * ld_logp + i] = v;
    }
    __syncthreads();
    // per-warp top-k over a contiguous slice: k rounds of (lane-local scan of smem, warp arg-max, knock out)
    const int per = (V + 7) / 8;
    const int lo = warp * per, hi = min(V, lo + per);
    for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int idx = lo + lane; idx < hi; idx += 32) {
            const float v = srow[idx];
            if (v > bv) { bv = v; bi = idx; }      // ascending scan + strict '>' keeps the lower index on ties
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (bi != 0x7fffffff && ((bi - lo) & 31) == lane) srow[bi] = -INFINITY;
        if (lane == 0) { cand_v[warp * KMAX + r] = bv; cand_i[warp * KMAX + r] = bi; }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
        // merge: 8 sorted lists of k -> global top-k (lane w < 8 walks list w)
        int pos = 0;
        for (int r = 0; r < k; ++r) {
            float bv = (lane < 8 && pos < k) ? cand_v[lane * KMAX + pos] : -INFINITY;
            int bi = (lane < 8 && pos < k) ? cand_i[lane * KMAX + pos] : 0x7fffffff;
            const float mv = bv;
            const int mi = bi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (mi == bi && mv == bv && lane < 8) ++pos;
            if (lane == 0) { out_val[(size_t)row * k + r] = bv; out_idx[(size_t)row * k + r] = bi; }
        }
    }
}

const char* logsoftmax_topk_launch(cudaStream_t stream, const float* logits, int ldl, int V, const float* lm_logp,
                                   int ld_lm, float lm_weight, int k, int rows, float* out_val, int* out_idx,
                                   float* out_logp, int ld_logp) {
    if (k < 1 || k > KMAX) return "logsoftmax_topk: k must be in [1,16]";
    if (V < 1 || V > 48 * 1024) return "logsoftmax_topk: vocabulary too large for one CTA's shared memory";
    const size_t smem = (size_t)V * sizeof(float);
    static bool attr_set = false;
    if (!attr_set && smem > 40 * 1024) {
        if (cudaFuncSetAttribute(logsoftmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return "cudaFuncSetAttribute(logsoftmax_topk) failed";
        attr_set = true;
    }
    cudaError_t e = launch_pdl(logsoftmax_topk_kernel, dim3(rows), dim3(TOPK_THREADS), smem, stream, logits, ldl, V, lm_logp, ld_lm,
                               lm_weight, k, out_val, out_idx, out_logp, ld_logp);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace otb

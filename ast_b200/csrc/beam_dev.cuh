// Device-side beam-search bookkeeping (nn.py:235-322), shared by the per-step kernels (beam.cu) and the persistent
// beam-search kernel (beam_seq.cu).  Every read of mutable beam state goes through L2 (__ldcg): inside the persistent
// kernel the state is written by other CTAs between grid barriers and L1 is not coherent.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ast {

// Row r of the live beam (whole CTA).  lp = z - (max + log(sum exp(z - max)))  (Chainer F.log_softmax); top-K in
// descending lp; exact ties -> larger token id first (a stable ascending argsort reversed).  lp: V floats of smem.
__device__ __forceinline__ void beam_topk_row(const float* z, int ldz, int V, int K, const BeamState& bs, float* cand_lp,
                                              int* cand_tok, int r, float* lp, float* scratch, int* iscratch) {
    if (r >= __ldcg(bs.n_active) || __ldcg(bs.finished + r)) return;
    const float* zr = z + (size_t)r * ldz;
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < V; n += blockDim.x) mx = fmaxf(mx, __ldcg(zr + n));
    mx = block_max(mx, scratch);
    float sum = 0.f;
    for (int n = threadIdx.x; n < V; n += blockDim.x) sum += expf(__ldcg(zr + n) - mx);
    sum = block_sum(sum, scratch);
    const float lse = mx + logf(sum);
    __syncthreads();
    for (int n = threadIdx.x; n < V; n += blockDim.x) lp[n] = __ldcg(zr + n) - lse;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = 0; k < K; ++k) {
        float bv = -INFINITY; int bi = -1;
        for (int n = threadIdx.x; n < V; n += blockDim.x) {
            const float v = lp[n];
            if (v > bv || (v == bv && n > bi)) { bv = v; bi = n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { scratch[w] = bv; iscratch[w] = bi; }
        __syncthreads();
        bv = (lane < nw) ? scratch[lane] : -INFINITY;
        bi = (lane < nw) ? iscratch[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) {
            cand_lp[r * K + k] = bv; cand_tok[r * K + k] = bi;
            if (bi >= 0) lp[bi] = -INFINITY;      // NaN/inf-safe removal from the pool
        }
        __syncthreads();
    }
}

// One CTA: candidate list in the reference's order (finished hypotheses carried over, nn.py:314-318), stable descending
// rank (nn.py:320), keep N.  sraw: N*K*(4+3*4) bytes of smem.
__device__ __forceinline__ void beam_prune_cta(const BeamState& bs, const float* cand_lp, const int* cand_tok, int N, int K, int step,
                                               int eos, int* hist_parent, int* hist_tok, unsigned char* sraw, int* ncand_sh) {
    const int maxc = N * K;
    float* cs = reinterpret_cast<float*>(sraw);
    int* cpar = reinterpret_cast<int*>(cs + maxc);
    int* ctok = cpar + maxc;
    int* cfin = ctok + maxc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        const int na = __ldcg(bs.n_active);
        for (int e = 0; e < na; ++e) {
            if (__ldcg(bs.finished + e)) { cs[n] = __ldcg(bs.score + e); cpar[n] = e; ctok[n] = -1; cfin[n] = 1; ++n; }
            else for (int k = 0; k < K; ++k) {
                const int tk = __ldcg(cand_tok + e * K + k);
                // float32 accumulation, exactly `score + pred_probs[pi]` (nn.py:289)
                cs[n] = __fadd_rn(__ldcg(bs.score + e), __ldcg(cand_lp + e * K + k));
                cpar[n] = e; ctok[n] = tk; cfin[n] = (tk == eos) ? 1 : 0; ++n;
            }
        }
        *ncand_sh = n;
    }
    __syncthreads();
    const int n = *ncand_sh;
    const int keep = min(N, n);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float si = cs[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = cs[j];
            rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
        }
        if (rank < keep) {
            bs.new_score[rank] = si;
            bs.new_parent[rank] = cpar[i];
            bs.new_tok[rank] = ctok[i];
            bs.new_finished[rank] = cfin[i];
            hist_parent[(size_t)step * N + rank] = cpar[i];
            hist_tok[(size_t)step * N + rank] = ctok[i];
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int allfin = 1;
        for (int r = 0; r < keep; ++r) allfin &= __ldcg(bs.new_finished + r);
        bs.n_active[0] = keep;
        bs.steps_done[0] = step + 1;
        for (int r = 0; r < keep; ++r) {
            bs.score[r] = __ldcg(bs.new_score + r);
            bs.finished[r] = __ldcg(bs.new_finished + r);
        }
        if (allfin) bs.done[0] = 1;       // honoured from the NEXT step on (nn.py:308-311)
    }
}

// Gather decoder state by parent for beam slot r (whole CTA): dst slot r <- (carried ? current state of the parent :
// post-step state of the parent); alpha history row; last token.
__device__ __forceinline__ void beam_gather_row(const BeamState& bs, const BeamGather& gd, int N, int step, int Tp,
                                                const float* alpha_step, float* alpha_hist, const int* last_tok_prev,
                                                int* last_tok_next, int r) {
    if (r >= __ldcg(bs.n_active)) return;
    const int par = __ldcg(bs.new_parent + r);
    const bool carry = __ldcg(bs.new_tok + r) < 0;
    for (int t = 0; t < gd.n; ++t) {
        const float* src = (carry ? gd.cur[t] : gd.post[t]) + (size_t)par * gd.width[t];
        float* dst = gd.nxt[t] + (size_t)r * gd.width[t];
        for (int j = threadIdx.x; j < gd.width[t]; j += blockDim.x) dst[j] = __ldcg(src + j);
    }
    float* ah = alpha_hist + ((size_t)step * N + r) * Tp;
    for (int j = threadIdx.x; j < Tp; j += blockDim.x) ah[j] = carry ? 0.f : __ldcg(alpha_step + (size_t)par * Tp + j);
    if (threadIdx.x == 0) last_tok_next[r] = carry ? __ldcg(last_tok_prev + par) : __ldcg(bs.new_tok + r);
}

}  // namespace ast

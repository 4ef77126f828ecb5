// On-device beam search bookkeeping (nn.py:235-322).  The N live hypotheses are the decoder batch;
// per step: log-softmax + per-hypothesis top-K (nn.py:269-270), candidate merge with finished-hyp
// carry-over and the reference's stable descending sort (nn.py:314-321), then a state gather by
// parent index.  No host round trip inside the search; the host reads tokens / parents / scores /
// attention history once at the end.
#include "common.cuh"
#include "kernels.h"

namespace ast {

// One CTA per hypothesis row.  lp = z - (max + log(sum exp(z - max)))  (Chainer F.log_softmax).
// Top-K in descending lp; exact ties -> larger token id first (a stable ascending argsort reversed).
__global__ void beam_topk_kernel(const float* __restrict__ z, int ldz, int V, int K, const BeamState bs,
                                 float* __restrict__ cand_lp, int* __restrict__ cand_tok) {
    if (bs.done[0]) return;
    const int r = blockIdx.x;
    if (r >= bs.n_active[0] || bs.finished[r]) return;
    extern __shared__ float lp[];          // V
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    const float* zr = z + (size_t)r * ldz;
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < V; n += blockDim.x) mx = fmaxf(mx, zr[n]);
    mx = block_max(mx, scratch);
    float sum = 0.f;
    for (int n = threadIdx.x; n < V; n += blockDim.x) sum += expf(zr[n] - mx);
    sum = block_sum(sum, scratch);
    const float lse = mx + logf(sum);
    for (int n = threadIdx.x; n < V; n += blockDim.x) lp[n] = zr[n] - lse;
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

// Single CTA: candidate list in the reference's order, stable descending rank, keep N.
__global__ void beam_prune_kernel(BeamState bs, const float* __restrict__ cand_lp, const int* __restrict__ cand_tok,
                                  int N, int K, int step, int eos, int* __restrict__ hist_parent,
                                  int* __restrict__ hist_tok) {
    if (bs.done[0]) return;
    extern __shared__ unsigned char sraw[];
    const int maxc = N * K;
    float* cs = reinterpret_cast<float*>(sraw);
    int* cpar = reinterpret_cast<int*>(cs + maxc);
    int* ctok = cpar + maxc;
    int* cfin = ctok + maxc;
    __shared__ int ncand;
    if (threadIdx.x == 0) {
        int n = 0;
        const int na = bs.n_active[0];
        for (int e = 0; e < na; ++e) {
            if (bs.finished[e]) { cs[n] = bs.score[e]; cpar[n] = e; ctok[n] = -1; cfin[n] = 1; ++n; }
            else for (int k = 0; k < K; ++k) {
                const int tk = cand_tok[e * K + k];
                // float32 accumulation, exactly `score + pred_probs[pi]` (nn.py:289)
                cs[n] = __fadd_rn(bs.score[e], cand_lp[e * K + k]);
                cpar[n] = e; ctok[n] = tk; cfin[n] = (tk == eos) ? 1 : 0; ++n;
            }
        }
        ncand = n;
    }
    __syncthreads();
    const int n = ncand;
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
    __syncthreads();
    if (threadIdx.x == 0) {
        int allfin = 1;
        for (int r = 0; r < keep; ++r) allfin &= bs.new_finished[r];
        bs.n_active[0] = keep;
        bs.steps_done[0] = step + 1;
        for (int r = 0; r < keep; ++r) {
            bs.score[r] = bs.new_score[r];
            bs.finished[r] = bs.new_finished[r];
        }
        if (allfin) bs.done[0] = 1;       // honoured from the NEXT step on (nn.py:308-311)
    }
}

// Gather decoder state by parent: dst slot r <- (carried ? cur state of parent : post-step state of parent).
// State vectors are laid out [slot][width]; one launch per state tensor family via the descriptor.
__global__ void beam_gather_kernel(const BeamState bs, BeamGather gd, int N, int step, int Tp,
                                   const float* __restrict__ alpha_step, float* __restrict__ alpha_hist,
                                   const int* __restrict__ last_tok_prev, int* __restrict__ last_tok_next) {
    if (bs.steps_done[0] != step + 1) return;     // this step's prune did not run (search already over)
    const int r = blockIdx.x;
    if (r >= bs.n_active[0]) return;
    const int par = bs.new_parent[r];
    const bool carry = bs.new_tok[r] < 0;
    for (int t = 0; t < gd.n; ++t) {
        const float* src = (carry ? gd.cur[t] : gd.post[t]) + (size_t)par * gd.width[t];
        float* dst = gd.nxt[t] + (size_t)r * gd.width[t];
        for (int j = threadIdx.x; j < gd.width[t]; j += blockDim.x) dst[j] = src[j];
    }
    float* ah = alpha_hist + ((size_t)step * N + r) * Tp;
    for (int j = threadIdx.x; j < Tp; j += blockDim.x) ah[j] = carry ? 0.f : alpha_step[(size_t)par * Tp + j];
    if (threadIdx.x == 0) last_tok_next[r] = carry ? last_tok_prev[par] : bs.new_tok[r];
}

int beam_topk(cudaStream_t st, const float* z, int ldz, int V, int K, int N, const BeamState& bs, float* cand_lp, int* cand_tok) {
    beam_topk_kernel<<<N, 256, sizeof(float) * V, st>>>(z, ldz, V, K, bs, cand_lp, cand_tok);
    AST_LAUNCH_OK();
    return 0;
}
int beam_prune(cudaStream_t st, const BeamState& bs, const float* cand_lp, const int* cand_tok, int N, int K, int step,
               int eos, int* hist_parent, int* hist_tok) {
    const size_t smem = (size_t)N * K * (sizeof(float) + 3 * sizeof(int));
    beam_prune_kernel<<<1, 128, smem, st>>>(bs, cand_lp, cand_tok, N, K, step, eos, hist_parent, hist_tok);
    AST_LAUNCH_OK();
    return 0;
}
int beam_gather(cudaStream_t st, const BeamState& bs, const BeamGather& gd, int N, int step, int Tp,
                const float* alpha_step, float* alpha_hist, const int* last_tok_prev, int* last_tok_next) {
    beam_gather_kernel<<<N, 128, 0, st>>>(bs, gd, N, step, Tp, alpha_step, alpha_hist, last_tok_prev, last_tok_next);
    AST_LAUNCH_OK();
    return 0;
}

}  // namespace ast

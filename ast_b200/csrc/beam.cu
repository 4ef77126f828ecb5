// On-device beam search bookkeeping (nn.py:235-322), one kernel per phase.  The N live hypotheses are the decoder batch;
// per step: log-softmax + per-hypothesis top-K (nn.py:269-270), candidate merge with finished-hyp carry-over and the
// reference's stable descending sort (nn.py:314-321), then a state gather by parent index.  No host round trip inside the
// search.  The persistent single-launch version of the whole search is beam_seq.cu; the device bodies are shared
// (beam_dev.cuh).
#include "beam_dev.cuh"

namespace ast {

__global__ void beam_topk_kernel(const float* __restrict__ z, int ldz, int V, int K, const BeamState bs,
                                 float* __restrict__ cand_lp, int* __restrict__ cand_tok) {
    extern __shared__ float lp[];          // V
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    if (bs.done[0]) return;
    beam_topk_row(z, ldz, V, K, bs, cand_lp, cand_tok, blockIdx.x, lp, scratch, iscratch);
}

__global__ void beam_prune_kernel(BeamState bs, const float* __restrict__ cand_lp, const int* __restrict__ cand_tok,
                                  int N, int K, int step, int eos, int* __restrict__ hist_parent,
                                  int* __restrict__ hist_tok) {
    extern __shared__ unsigned char sraw[];
    __shared__ int ncand;
    if (bs.done[0]) return;
    beam_prune_cta(bs, cand_lp, cand_tok, N, K, step, eos, hist_parent, hist_tok, sraw, &ncand);
}

__global__ void beam_gather_kernel(const BeamState bs, BeamGather gd, int N, int step, int Tp,
                                   const float* __restrict__ alpha_step, float* __restrict__ alpha_hist,
                                   const int* __restrict__ last_tok_prev, int* __restrict__ last_tok_next) {
    if (bs.steps_done[0] != step + 1) return;     // this step's prune did not run (search already over)
    beam_gather_row(bs, gd, N, step, Tp, alpha_step, alpha_hist, last_tok_prev, last_tok_next, blockIdx.x);
}

// ---- batched over utterances: blockIdx.y = utterance g, every per-utterance array is g-major ------------------------------
// BeamState arrays: score / finished / new_* are [G][N]; n_active / done / steps_done are [G]; cand_* are [G][N][K];
// hist_* are [G][stop_limit][N]; alpha_hist is [G][stop_limit][N][Tp_ld]; decoder rows are g * N + slot.
__device__ __forceinline__ BeamState beam_state_of(const BeamState& b, int g, int N) {
    BeamState o;
    o.score = b.score + (size_t)g * N; o.finished = b.finished + (size_t)g * N;
    o.n_active = b.n_active + g; o.done = b.done + g; o.steps_done = b.steps_done + g;
    o.new_score = b.new_score + (size_t)g * N; o.new_parent = b.new_parent + (size_t)g * N;
    o.new_tok = b.new_tok + (size_t)g * N; o.new_finished = b.new_finished + (size_t)g * N;
    return o;
}

__global__ void beam_topk_batch_kernel(const float* __restrict__ z, int ldz, int V, int K, int N, const BeamState bs,
                                       float* __restrict__ cand_lp, int* __restrict__ cand_tok) {
    extern __shared__ float lp[];          // V
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    const int g = blockIdx.y;
    const BeamState b = beam_state_of(bs, g, N);
    if (__ldcg(b.done)) return;
    beam_topk_row(z + (size_t)g * N * ldz, ldz, V, K, b, cand_lp + (size_t)g * N * K, cand_tok + (size_t)g * N * K, blockIdx.x, lp,
                  scratch, iscratch);
}

__global__ void beam_prune_batch_kernel(BeamState bs, const float* __restrict__ cand_lp, const int* __restrict__ cand_tok, int N, int K,
                                        int step, int eos, int stop_limit, int* __restrict__ hist_parent, int* __restrict__ hist_tok) {
    extern __shared__ unsigned char sraw[];
    __shared__ int ncand;
    const int g = blockIdx.x;
    const BeamState b = beam_state_of(bs, g, N);
    if (__ldcg(b.done)) return;
    beam_prune_cta(b, cand_lp + (size_t)g * N * K, cand_tok + (size_t)g * N * K, N, K, step, eos,
                   hist_parent + (size_t)g * stop_limit * N, hist_tok + (size_t)g * stop_limit * N, sraw, &ncand);
}

__global__ void beam_gather_batch_kernel(const BeamState bs, BeamGather gd, int N, int step, int Tp_ld, int stop_limit,
                                         const float* __restrict__ alpha_step, float* __restrict__ alpha_hist,
                                         const int* __restrict__ last_tok_prev, int* __restrict__ last_tok_next) {
    const int g = blockIdx.y;
    const BeamState b = beam_state_of(bs, g, N);
    const bool ran = __ldcg(b.steps_done) == step + 1;     // this step's prune ran for this utterance
    if (!ran) {
        // finished search: its rows must survive the bank toggle of the utterances that are still running
        const int r = blockIdx.x;
        for (int t = 0; t < gd.n; ++t) {
            const float* src = gd.cur[t] + ((size_t)g * N + r) * gd.width[t];
            float* dst = gd.nxt[t] + ((size_t)g * N + r) * gd.width[t];
            for (int j = threadIdx.x; j < gd.width[t]; j += blockDim.x) dst[j] = __ldcg(src + j);
        }
        if (threadIdx.x == 0) last_tok_next[(size_t)g * N + r] = __ldcg(last_tok_prev + (size_t)g * N + r);
        return;
    }
    BeamGather o = gd;
    for (int t = 0; t < gd.n; ++t) {
        const size_t off = (size_t)g * N * gd.width[t];
        o.cur[t] = gd.cur[t] + off; o.post[t] = gd.post[t] + off; o.nxt[t] = gd.nxt[t] + off;
    }
    beam_gather_row(b, o, N, step, Tp_ld, alpha_step + (size_t)g * N * Tp_ld, alpha_hist + (size_t)g * stop_limit * N * Tp_ld,
                    last_tok_prev + (size_t)g * N, last_tok_next + (size_t)g * N, blockIdx.x);
}

int beam_step_batch(cudaStream_t st, int G, const float* z, int ldz, int V, int K, int N, const BeamState& bs, float* cand_lp,
                    int* cand_tok, int step, int eos, int stop_limit, int* hist_parent, int* hist_tok, const BeamGather& gd, int Tp_ld,
                    const float* alpha_step, float* alpha_hist, const int* last_tok_prev, int* last_tok_next) {
    beam_topk_batch_kernel<<<dim3(N, G), 256, sizeof(float) * V, st>>>(z, ldz, V, K, N, bs, cand_lp, cand_tok);
    AST_LAUNCH_OK();
    const size_t smem = (size_t)N * K * (sizeof(float) + 3 * sizeof(int));
    beam_prune_batch_kernel<<<G, 128, smem, st>>>(bs, cand_lp, cand_tok, N, K, step, eos, stop_limit, hist_parent, hist_tok);
    AST_LAUNCH_OK();
    beam_gather_batch_kernel<<<dim3(N, G), 128, 0, st>>>(bs, gd, N, step, Tp_ld, stop_limit, alpha_step, alpha_hist, last_tok_prev,
                                                         last_tok_next);
    AST_LAUNCH_OK();
    return 0;
}

// all_done[0] = AND over utterances of done[g]
__global__ void beam_all_done_kernel(const int* __restrict__ done, int G, int* __restrict__ all_done) {
    int v = 1;
    for (int g = threadIdx.x; g < G; g += blockDim.x) v &= (__ldcg(done + g) != 0);
    v = __syncthreads_and(v);
    if (threadIdx.x == 0) all_done[0] = v;
}
int beam_all_done(cudaStream_t st, const int* done, int G, int* all_done) {
    beam_all_done_kernel<<<1, 64, 0, st>>>(done, G, all_done);
    AST_LAUNCH_OK();
    return 0;
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

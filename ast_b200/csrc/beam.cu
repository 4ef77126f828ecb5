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

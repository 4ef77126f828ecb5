// Persistent beam search: the whole of NN.decode_beam's step loop (nn.py:299-321; per hypothesis: decode_step, log-softmax,
// top-K, then the global stable sort / prune and the per-hypothesis state hand-over) as ONE cooperative launch per
// utterance.  The N live hypotheses are the decoder batch.  A step is 12 phases separated by grid barriers instead of 13
// kernel launches (the per-launch version in model.cu::ast_beam_search spends ~140 us per step, mostly launch and drain
// latency around 5-10 us kernels).  Arithmetic is the fp32-faithful path of the decode_step protocol (3xTF32-split mma.sync,
// accurate transcendental functions), so hypotheses stay identical to the fp32 reference arithmetic; every phase body is the same device
// function the per-step kernels run (decoder_dev.cuh, beam_dev.cuh).
#include <cooperative_groups.h>
#include "beam_dev.cuh"
#include "decoder_dev.cuh"

namespace cg = cooperative_groups;

namespace ast {

__global__ void __launch_bounds__(SK_THREADS, 1)
beam_seq_kernel(BeamSeq p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ SkinnySmem sm;
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    __shared__ int ncand;
    extern __shared__ float dsm[];
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int N = p.N, H = p.H, E = p.E, A = p.A, Tp = p.Tp, NL = p.NL, V = p.V, Vp = p.Vp;
    const int ldx0 = E + A;

    for (int s = 0; s < p.stop_limit; ++s) {
        if (__ldcg(p.bs.done)) break;                 // set by the previous step's prune (uniform: read after a barrier)
        const int bank = s & 1;
        // ---- decoder input: x0[r] = [Emb[last token of hyp r] ; attn_v of hyp r]  (seq2seq.py:365-372, eval mode) ----
        for (int r = cta; r < N; r += ncta) {
            const int word = min(max(__ldcg(p.words[bank] + r), 0), V - 1);
            float* dst = p.x0 + (size_t)r * ldx0;
            for (int j = threadIdx.x; j < E; j += blockDim.x) dst[j] = __ldg(p.emb + (size_t)word * E + j);
            for (int j = threadIdx.x; j < A; j += blockDim.x) dst[E + j] = __ldcg(p.ht[bank] + (size_t)r * A + j);
        }
        grid.sync();
        // ---- LSTM stack (seq2seq.py:375); the step's new states go to the "post" bank, the prune decides who keeps them ----
        for (int l = 0; l < NL; ++l) {
            SkinnyArgs a{};
            const int in = l == 0 ? ldx0 : H;
            a.X[0] = l == 0 ? p.x0 : p.hd[l - 1]; a.ldx[0] = l == 0 ? ldx0 : H;
            a.K[0] = in; a.W[0] = p.Wup[l]; a.ldw[0] = in;
            a.X[1] = p.h[bank][l]; a.ldx[1] = H; a.K[1] = H; a.W[1] = p.Wlat[l]; a.ldw[1] = H;
            a.bias = p.bup[l]; a.B = N; a.N = 4 * H; a.epi = EPI_LSTM;
            a.Y = p.act; a.ldy = 4 * H;
            a.c_prev = p.c[bank][l]; a.c_out = p.cpost[l]; a.h_out = p.hpost[l];
            if (l == NL - 1) { a.hd_out = p.cvh + H; a.ld_hd = 2 * H; }
            else { a.hd_out = p.hd[l]; a.ld_hd = H; }
            a.drop = 0.f; a.seed = 0; a.drop_stream = 0; a.drop_base = 0;
            for (int g = cta; g < (4 * H) / SK_COLS; g += ncta) skinny_tile<1, true>(a, g * SK_COLS, sm);
            grid.sync();
        }
        {   // q = attn_Wa(h)  (:341)
            SkinnyArgs a{};
            a.X[0] = p.cvh + H; a.ldx[0] = 2 * H; a.K[0] = H; a.W[0] = p.Wa; a.ldw[0] = H; a.bias = p.ba;
            a.B = N; a.N = H; a.epi = EPI_NONE; a.Y = p.q; a.ldy = H;
            for (int g = cta; g < H / SK_COLS; g += ncta) skinny_tile<1, true>(a, g * SK_COLS, sm);
            grid.sync();
        }
        {   // scores (:342); one utterance -> enc batch stride 0
            const int ntb = (Tp + 7) / 8;
            for (int i = cta; i < ntb * N; i += ncta) attn_dot_block(p.enc, 0, p.q, H, p.scores, Tp, H, i / ntb, i % ntb);
            grid.sync();
        }
        {   // softmax over T' + context (:351-355)
            const int njb = (H + 127) / 128;
            for (int i = cta; i < njb * N; i += ncta)
                attn_ctx_block(p.enc, 0, p.scores, p.alpha, p.cvh, 2 * H, Tp, H, i / njb, i % njb, dsm, dsm + ((Tp + 3) & ~3), scratch);
            grid.sync();
        }
        {   // ht = tanh(context([cv ; h]))  (:386-390)
            SkinnyArgs a{};
            a.X[0] = p.cvh; a.ldx[0] = 2 * H; a.K[0] = 2 * H; a.W[0] = p.Wc; a.ldw[0] = 2 * H; a.bias = p.bc;
            a.B = N; a.N = A; a.epi = EPI_TANH; a.Y = p.htout; a.ldy = A;
            for (int g = cta; g < A / SK_COLS; g += ncta) skinny_tile<1, true>(a, g * SK_COLS, sm);
            grid.sync();
        }
        {   // logits = out(ht)  (:394)
            SkinnyArgs a{};
            a.X[0] = p.htout; a.ldx[0] = A; a.K[0] = A; a.W[0] = p.Wo; a.ldw[0] = A; a.bias = p.bo;
            a.B = N; a.N = V; a.epi = EPI_NONE; a.Y = p.logits; a.ldy = Vp;
            for (int g = cta; g < (V + SK_COLS - 1) / SK_COLS; g += ncta) skinny_tile<1, true>(a, g * SK_COLS, sm);
            grid.sync();
        }
        // ---- log-softmax + top-K per live hypothesis (nn.py:269-270) -------------------------------------------------
        for (int r = cta; r < N; r += ncta) {
            __syncthreads();
            beam_topk_row(p.logits, Vp, V, p.K, p.bs, p.cand_lp, p.cand_tok, r, dsm, scratch, iscratch);
        }
        grid.sync();
        // ---- merge, stable sort, keep N (nn.py:314-321) ----------------------------------------------------------------
        if (cta == 0)
            beam_prune_cta(p.bs, p.cand_lp, p.cand_tok, N, p.K, s, p.eos, p.hist_parent, p.hist_tok, reinterpret_cast<unsigned char*>(dsm), &ncand);
        grid.sync();
        // ---- hand the decoder state over to the kept hypotheses ---------------------------------------------------------
        {
            BeamGather gd{}; gd.n = 0;
            for (int l = 0; l < NL; ++l) {
                gd.cur[gd.n] = p.h[bank][l]; gd.post[gd.n] = p.hpost[l]; gd.nxt[gd.n] = p.h[bank ^ 1][l]; gd.width[gd.n++] = H;
                gd.cur[gd.n] = p.c[bank][l]; gd.post[gd.n] = p.cpost[l]; gd.nxt[gd.n] = p.c[bank ^ 1][l]; gd.width[gd.n++] = H;
            }
            gd.cur[gd.n] = p.ht[bank]; gd.post[gd.n] = p.htout; gd.nxt[gd.n] = p.ht[bank ^ 1]; gd.width[gd.n++] = A;
            for (int r = cta; r < N; r += ncta)
                beam_gather_row(p.bs, gd, N, s, Tp, p.alpha, p.alpha_hist, p.words[bank], p.words[bank ^ 1], r);
        }
        grid.sync();
    }
}

int beam_seq(cudaStream_t st, const BeamSeq& p) {
    AST_CHECK(p.N >= 1 && p.N <= 16, "beam_seq: beam width %d unsupported by the persistent kernel (1..16)", p.N);
    AST_CHECK(p.H % 16 == 0 && p.A % 16 == 0 && p.E % 16 == 0 && p.Vp % 16 == 0, "beam_seq: H, A, E, Vp must be multiples of 16");
    int dev = 0, sms = 0, occ = 0;
    AST_CUDA_OK(cudaGetDevice(&dev));
    AST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t want = std::max<size_t>(std::max<size_t>(sizeof(float) * (size_t)p.V, sizeof(float) * ((size_t)p.Tp + 4 + 8 * 128)),
                                         (size_t)p.N * p.K * 16);
    const size_t dsm = (want + 15) & ~(size_t)15;
    AST_CUDA_OK(cudaFuncSetAttribute(beam_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
    AST_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, beam_seq_kernel, SK_THREADS, dsm));
    AST_CHECK(occ >= 1, "beam search kernel does not fit on an SM");
    BeamSeq pp = p;
    void* args[] = {&pp};
    AST_CUDA_OK(cudaLaunchCooperativeKernel((const void*)beam_seq_kernel, dim3(sms), dim3(SK_THREADS), args, dsm, st));
    ++g_kernel_launches;
    return 0;
}

}  // namespace ast

// Internal launcher interface between model.cu (orchestration) and the kernel translation units.
// Every launcher returns 0 on success, <0 on error (message via ast::get_last_error()).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <algorithm>

#define AST_MAX_CHAINS 4

namespace ast {

// ---- dense GEMM (row-major, see gemm_simt.cu) -----------------------------------------------
int sgemm_simt(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha, const float* A, int lda,
               const float* B, int ldb, float beta, float* C, int ldc, const float* bias);
// TF32 tcgen05/TMEM/TMA GEMM for C = A(MxK, k-contig) * B(NxK, k-contig)^T (+bias); returns 1 when the
// shape is not supported by the tensor-core kernel (caller falls back to sgemm_simt).
int gemm_tc_nt(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
               int ldc, const float* bias, float beta);
// General form (any operand major); split_k: 0 = none, -1 = automatic (weight gradients), >0 = that many.
int gemm_tc(cudaStream_t st, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            float* C, int ldc, const float* bias, float beta, int split_k);

// gated variant for the persistent encoder wavefront (gemm_tc.cu).  Rows of A are (step, batch) pairs, B batch rows per step; the
// kernel producing A counts its CTAs into wait[chunk] per chunk of `chunk` steps in its processing order (rev: from step T-1
// down, and the m-tiles are then walked from the last one).  A tile's rows are read once wait[its chunk] >= target; every
// finished tile adds 4 to done[m-tile of 128 rows].
// 2-CTA (cta_group::2) 256 x 256 tiles, all operand majors, optional split-K (0 none, -1 automatic, > 0 count): C = A . op(B) (+bias, + beta*C); returns 1 if it cannot take the operands
int gemm_tc2(cudaStream_t st, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             const float* bias, float beta, int split_k);
// up to 4 same-shape problems in one 2-CTA launch (tile space: problem x split x m x n)
int gemm_tc2_grouped(cudaStream_t st, int n, bool ta, bool tb, int M, int N, int K, const float* const* A, int lda, const float* const* B, int ldb,
                     float* const* C, int ldc, const float* const* bias, float beta, int split_k);
void gemm_tc_set_cta_cap(int cap);   // 0 = no cap; applies to gemm_tc() launches issued afterwards by this thread's caller
struct TcGate { const unsigned* wait; unsigned target; int B, chunk, T; bool rev; unsigned* done; };
int gemm_tc_tiles_per_row(int N);
int gemm_tc_gated(cudaStream_t st, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* bias, const TcGate& gate, int ctas);
// fp32-faithful 3xTF32 NT GEMM (gemm_tc.cu) on (hi, lo) = split_tf32(operand); returns 1 if the shape is unsupported
int split_tf32(cudaStream_t st, const float* x, float* hi, float* lo, size_t n);
int gemm_tc3_nt(cudaStream_t st, int M, int N, int K, const float* A, const float* Alo, int lda, const float* B, const float* Blo,
                int ldb, float* C, int ldc, const float* bias, bool allow_split = false);

// ---- CNN front-end ---------------------------------------------------------------------------
int im2col0(cudaStream_t st, const float* X, float* cols, int B, int T, int D, int Fp, int T1, int kh, int kw, int sh,
            int sw, int ph, int ldc);
int bn_stats(cudaStream_t st, const float* x, double* stats, int rows, int C, int seg_rows, int seg_valid, double* partials,
             int partial_blocks);
int bn_stats_finalize(cudaStream_t st, const float* x, double* stats, int rows, int C, int seg_rows, int seg_valid, double* partials,
                      int partial_blocks, float* mean, float* invstd, float* avg_mean, float* avg_var, double m, float eps, float decay,
                      bool update_running);      // bn_stats + bn_finalize with the partial-sum pass and the finalisation in one launch
int bn_finalize(cudaStream_t st, const double* stats, float* mean, float* invstd, float* avg_mean, float* avg_var,
                int C, double m, float eps, float decay, bool update_running);
int bn_eval_prepare(cudaStream_t st, const float* avg_mean, const float* avg_var, float* mean, float* invstd, int C,
                    float eps);
int bn_relu_pad(cudaStream_t st, const float* raw, float* out, const float* mean, const float* invstd,
                const float* gamma, const float* beta, int nseg, int T1, int S0, int pad, int C);
int bn_relu_to_rnn(cudaStream_t st, const float* raw, float* rnn_in, float* rnn_rev, const float* mean,
                   const float* invstd, const float* gamma, const float* beta, int B, int Fp, int Rs, int Tp, int C);
int bn_bwd_from_rnn(cudaStream_t st, const float* d_in, const float* d_rev, const float* raw, float* dx,
                    const float* mean, const float* invstd, const float* gamma, const float* beta, double* stats,
                    float* dgamma, float* dbeta, int B, int Fp, int Rs, int Tp, int C, double* partials, int partial_blocks);
int bn_bwd_from_padded(cudaStream_t st, const float* da0p, const float* raw, float* dx, const float* mean,
                       const float* invstd, const float* gamma, const float* beta, double* stats,
                       float* dgamma, float* dbeta, int nseg, int T1, int S0, int pad, int C, double* partials,
                       int partial_blocks);
int col2im1(cudaStream_t st, const float* dA, float* da0p, int nseg, int S0, int Rs, int Tp, int C0, int kh, int sh);
int permute_w1(cudaStream_t st, const float* src, float* dst, int Co, int Ci, int Kt, bool to_p);
int build_w1t(cudaStream_t st, const float* W1p, float* Wt, int Co, int Ci, int Kt, int p);   // transposed-convolution weights, parity p

// ---- persistent LSTM recurrence ----------------------------------------------------------------
struct LstmChain {
    float* G;            // (T x B x 4h): in  x-projection + bias ; out  activated gates (fwd) / dG (bwd)
    const float* Wl;     // (4h x h) lateral weight
    float* Hs;           // ((T+1) x B x h) link state h, slot 0 = initial state
    float* Cs;           // ((T+1) x B x h) cell state,  slot 0 = initial state
    float* out;          // fwd: post-dropout output, element (i,b,j) at out[i*out_si + b*out_sb + j]
    const float* dout;   // bwd: gradient w.r.t. that output, same addressing
    long long out_si, out_sb;
    const float* dh_fin; // bwd: gradient w.r.t. the final link state, element (b,j) at dh_fin[b*ld_dh_fin + j], or null
    const float* dc_fin;
    int ld_dh_fin, ld_dc_fin;
    float* dh0;          // bwd: gradient w.r.t. the initial state (B x h) or null
    float* dc0;
    unsigned drop_stream;
    unsigned drop_off;   // added to the dropout counter: (t0 * B * h) when this launch covers steps [t0, t0+T) of a longer sequence
    int b0, nb;          // batch rows [b0, b0+nb) of the B-row buffers handled by this chain (nb = 0: all B rows)
    // optional: the kernel's per-step input (forward: the x-projection G; backward: dout) is being written by a gated GEMM
    // (TcGate) while this kernel runs - its 128-row tile m is complete once tile_ready[m] >= tile_target.  null: complete at launch.
    const unsigned* tile_ready; unsigned tile_target;
};
struct LstmChains { LstmChain c[AST_MAX_CHAINS]; };
// Chunk signalling of a recurrence kernel that covers a whole sequence while its consumers are already running (the persistent
// encoder wavefront, model.cu).  Steps are numbered in PROCESSING order k (forward: k = t, backward: k = T-1-t); every CTA adds 1
// to done[k / chunk] (after a fence) once its outputs of that chunk are in global memory, and a gated GEMM (TcGate) waits for all
// CTAs of the launch.  null: no signalling.  (The kernel's own inputs are gated per 128-row tile: LstmChain::tile_ready.)
struct LstmGate { unsigned* done; int chunk; unsigned long long* ts; unsigned* resident; unsigned long long* probe; };   // ts (diagnostics, may be null): %globaltimer of block 0 at each chunk signal; resident (may be null): every CTA counts itself in when it starts
// One-CTA kernel that returns when *counter >= target (bounded spin): in front of the gated GEMMs of a wavefront, so that their single
// CTAs are dispatched only after every 8-CTA recurrence cluster is resident (single CTAs placed first can leave no GPC with 8 free SMs;
// the cluster then waits for another layer's kernel to end)
int wait_resident(cudaStream_t st, const unsigned* counter, unsigned target);
int lstm_seq_fwd(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop,
                 unsigned long long seed, bool exact);
int lstm_seq_bwd(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop,
                 unsigned long long seed, bool exact);
// tcgen05 versions (lstm_seq_tc.cu): TF32 mode, h == 256, chains already split into 16-row slices
int lstm_seq_tc_max_clusters(bool backward);     // cudaOccupancyMaxActiveClusters of the tcgen05 recurrence kernels (cached)
int lstm_seq_tc_cluster_size();
int lstm_seq_fwd_tc(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed);
void lstm_tc_set_prof(unsigned long long* p);   // diagnostics: device buffer of >= 128 u64 for the forward kernel's cycle probe
int lstm_seq_bwd_tc(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed);
// gated whole-sequence launches (TF32 tcgen05 kernels only: h == 256, !exact); *ncta = CTAs that will arrive on each done counter
bool lstm_seq_gated_supported(int h, bool exact);
int lstm_seq_fwd_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop, unsigned long long seed,
                       const LstmGate& gate, int* ncta);
int lstm_seq_bwd_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop, unsigned long long seed,
                       const LstmGate& gate, int* ncta);
int lstm_seq_fwd_tc_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed, const LstmGate& gate);
int lstm_seq_bwd_tc_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed, const LstmGate& gate);

// ---- decoder step kernels ------------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_TANH = 1, EPI_LSTM = 2, EPI_TANHBWD = 3, EPI_CELLBWD = 4 };
struct SkinnyArgs {
    const float* X[2]; int ldx[2]; int K[2];
    const float* W[2]; int ldw[2];
    const float* bias;
    int B, N, epi;
    float* Y; int ldy;
    const float* add; int ld_add;      // optional additive term (before the activation)
    const float* aux; int ld_aux;      // EPI_TANHBWD: tanh output
    // EPI_LSTM
    const float* c_prev; float* c_out; float* h_out; float* hd_out; int ld_hd;
    float drop; unsigned long long seed; unsigned drop_stream; size_t drop_base;
    float* Y2; int ldy2;               // optional second copy of the output (e.g. ht -> next step's input feeding slot)
    // EPI_CELLBWD: the output value (after `add`) is d(out) of hidden unit n of the layer below; the epilogue does that
    // layer's LSTM cell backward in place (act -> dG) for rows x units of this column group.
    float* cb_act; const float* cb_c; const float* cb_c_prev; float* cb_dc; const float* cb_dh_rec; int cb_ld_dh_rec; int cb_H;
    int cb_ncols;                      // the cell backward applies to output columns n < cb_ncols
    float* sc_demb; const int* sc_words; int sc_E;   // optional EmbedID scatter-add of columns n < sc_E
};

// Persistent decoder-sequence kernels (dec_seq.cu): everything a forward_loss decoder pass touches.
#define AST_MAXL 4
struct DecSeq {
    int B, S, L, H, E, A, V, Vp, Tp, NL;
    const float* emb; const float* Wup[AST_MAXL]; const float* bup[AST_MAXL]; const float* Wlat[AST_MAXL];
    const float* Wa; const float* ba; const float* Wc; const float* bc; const float* Wo; const float* bo;
    const float* WoT; const float* WcT; const float* WaT; const float* WcatT[AST_MAXL];
    const int* y; const unsigned char* use_true;
    const float* enc; float* d_enc;
    const float* dzw; float* dcv_all; float* ds_all;   // dec_seq2 backward: dz . Wo for all steps (in); per-step dcv and ds (out)
    const float* encW; const float* encb;   // dec_seq2: enc . W_a (B*T' x H) and enc . b_a (B x T'), precomputed per sequence
    float* x0; float* act[AST_MAXL]; float* Hd[AST_MAXL]; float* Cd[AST_MAXL]; float* hdd[AST_MAXL];
    float* q; float* scores; float* alpha; float* cvh; float* ht; float* logits; float* row_loss;
    int* words_used; int* argmax_steps;
    float* du; float* dcvh; float* dalpha; float* dq; float* dxh[AST_MAXL]; float* dcd[AST_MAXL]; float* demb;
    float drop_embed, drop_rnn; unsigned long long seed;
    int emb_done;                      // dec_seq2: teacher-forced embedding rows already written by embed_all()
    unsigned* bar;                     // grid-barrier counter (null: cooperative_groups grid.sync)
    int sync_all;                      // dec_seq2: grid barrier after every phase as well (debugging; the hand-off is by sentinel polling)
    // dec_seq2 backward hand-off slots, one per step, sentinel-filled before the launch: dG of every layer (the forward gates in
    // act[] stay intact), dh_rec of layer l produced at step s (consumed at s-1), d(ht) fed back into layer 0, h-half of du.Wc
    float* dgd[AST_MAXL]; float* dxr[AST_MAXL]; float* dfeed; float* dhh_all;
    unsigned long long* prof;          // optional phase-timing probe: CTA 0 stores %globaltimer at the end of each phase
    int prof_fine;                     // 1 + index of the CTA whose thread 0 stamps clock64 inside the phases of step 6 (0: off)
};
int dec_seq_fwd(cudaStream_t st, const DecSeq& p, bool exact);
int dec_seq_bwd(cudaStream_t st, const DecSeq& p, bool exact);
// second-generation forward (dec_seq2.cu): TMEM-resident weights, cluster K-split, logits deferred to the caller
bool dec_seq2_supported(const DecSeq& p);
int dec_seq2_fwd(cudaStream_t st, const DecSeq& p);
int dec_seq2_bwd(cudaStream_t st, const DecSeq& p);
// sentinel fill of the per-step hand-off slots the dec_seq2 kernels poll; ordered before init_dec_state / the kernel
int dec_seq2_prepare_fwd(cudaStream_t st, const DecSeq& p);
int dec_seq2_prepare_bwd(cudaStream_t st, const DecSeq& p);
int dec_seq2_sampled_argmax(cudaStream_t st, const DecSeq& p);   // argmax_steps of sampled steps <- the token the loop actually fed
int embed_all(cudaStream_t st, const DecSeq& p);   // x0[:, :E] / words_used for every step from the ground-truth tokens
int attn_denc(cudaStream_t st, const float* alpha, const float* ds, const float* dcv, const float* q, float* d_enc, int S, int B, int Tp, int H);
// softmax-CE (+ gradient in place, argmax) for every (step, row) of a decoder pass in one launch
int softmax_ce_all(cudaStream_t st, float* z, int ldz, const int* y, int ldy_tok, float* row_loss, int* argmax_out, int S, int B, int V);
int skinny(cudaStream_t st, const SkinnyArgs& p, bool exact);
int embed_concat(cudaStream_t st, const float* emb, const int* y, int ldy_tok, const unsigned char* use_true,
                 const int* prev_argmax, const int* forced_words, const float* ht_prev, int ld_ht, float* x0,
                 int* words_used, int B, int E, int A, int V, int step, float drop, unsigned long long seed,
                 unsigned drop_stream);
int embed_scatter(cudaStream_t st, float* demb, const float* dx0, int ld_dx, const int* words, int B, int E, int step,
                  float drop, unsigned long long seed, unsigned drop_stream);
int attn_dot(cudaStream_t st, const float* enc, long long enc_bs, const float* v, int ldv, float* s, int B, int Tp, int H);
int attn_grouped(cudaStream_t st, const float* enc, long long enc_bs, int rows_per_enc, const int* lens, const float* q, int ldq,
                 float* scores, float* alpha, float* cv, int ld_cv, int rows, int Tp_ld, int H);
int attn_ctx(cudaStream_t st, const float* enc, long long enc_bs, const float* s, float* alpha, float* cv, int ld_cv,
             int B, int Tp, int H);
int attn_bwd(cudaStream_t st, const float* enc, float* d_enc, long long enc_bs, const float* alpha, const float* dalpha,
             const float* dcv, int ld_dcv, const float* qv, int ld_q, float* dq, int ld_dq, int B, int Tp, int H);
int softmax_ce(cudaStream_t st, float* z, int ldz, const int* y, int ldy_tok, int step_next, float* row_loss,
               int* argmax_out, int B, int V, bool write_grad);
int loss_reduce(cudaStream_t st, const float* row_loss, int n, float* loss);
int split_concat2(cudaStream_t st, const float* src0, int ld0, int w0, const float* src1, int ld1, int w1, float* hi, float* lo, int ldo, int R);
int lstm_cell_rows(cudaStream_t st, float* act, const float* c_prev, float* c_out, float* h_out, float* hd_out, int ld_hd, int R, int H);
int tanh_rows(cudaStream_t st, float* x, size_t n);
int lstm_cell_bwd(cudaStream_t st, float* act, const float* c, const float* c_prev, const float* d_out, int ld_dout,
                  const float* dh_rec, int ld_dhrec, float* dc, int B, int H, int step_row0, float drop,
                  unsigned long long seed, unsigned drop_stream);

// ---- optimizer / data / misc -----------------------------------------------------------------------
struct FrozenRanges { int n; size_t begin[24]; size_t end[24]; };   // merged [begin, end) float ranges, ascending
int opt_sqnorm(cudaStream_t st, const float* g, const float* p, size_t n, float gscale, float wd, double* norm_sq);
int opt_amsgrad(cudaStream_t st, float* p, const float* g, float* m, float* v, float* vhat, size_t n, float gscale,
                float wd, float clip, const double* norm_sq, float alpha_t, float beta1, float beta2, float eps,
                const FrozenRanges& fr, float noise_sigma = 0.f, unsigned long long noise_seed = 0);
int opt_sgd(cudaStream_t st, float* p, const float* g, size_t n, float gscale, float wd, float clip, const double* norm_sq, float lr,
            const FrozenRanges& fr, float noise_sigma = 0.f, unsigned long long noise_seed = 0);
int pack_cmvn(cudaStream_t st, const float* raw, const long long* row_off, const int* lens, const float* scale,
              const float* offset, const unsigned char* keep, const float* noise, float noise_sigma,
              unsigned long long seed, float* X, int B, int T, int D);
int scale_inplace(cudaStream_t st, float* x, float a, size_t n);
int mul_noise(cudaStream_t st, const float* X, float* Y, const float* noise, float sigma, unsigned long long seed, size_t n);
int transpose(cudaStream_t st, const float* src, int ld_src, float* dst, int ld_dst, int R, int C);
int colsum(cudaStream_t st, const float* x, int ld, float* out, int rows, int C, bool accumulate);
int copy2d(cudaStream_t st, const float* src, long long ld_src, float* dst, long long ld_dst, int R, int C);
struct Copy2DJob { const float* src; long long ld_src; float* dst; long long ld_dst; int R, C; };
struct Copy2DBatch { int n; Copy2DJob job[16]; };
struct ZeroBatch { int n; float* ptr[16]; size_t count[16]; };
int zero_multi(cudaStream_t st, const ZeroBatch& b);   // zero up to 16 float ranges (counts % 4 == 0) in one launch
int copy2d_multi(cudaStream_t st, const Copy2DBatch& b);
int add_inplace(cudaStream_t st, float* dst, const float* src, size_t n);
int greedy_track(cudaStream_t st, const int* argmax, int* preds, int* seen, int* done_step, int B, int step, int eos);

// ---- beam search ---------------------------------------------------------------------------------------
struct BeamState {
    float* score; int* finished; int* n_active; int* done; int* steps_done;
    float* new_score; int* new_parent; int* new_tok; int* new_finished;
};
struct BeamGather { int n; const float* cur[8]; const float* post[8]; float* nxt[8]; int width[8]; };
// Persistent beam search (beam_seq.cu): everything one utterance's search touches.
struct BeamSeq {
    int N, K, V, Vp, H, E, A, Tp, NL, stop_limit, eos;
    const float* emb; const float* Wup[AST_MAXL]; const float* bup[AST_MAXL]; const float* Wlat[AST_MAXL];
    const float* Wa; const float* ba; const float* Wc; const float* bc; const float* Wo; const float* bo;
    const float* enc;                                   // (T' x H), one utterance
    float* h[2][AST_MAXL]; float* c[2][AST_MAXL]; float* ht[2]; int* words[2];     // two state banks (slot r = hypothesis r)
    float* hpost[AST_MAXL]; float* cpost[AST_MAXL];     // this step's new states before the hand-over
    float* x0; float* act; float* hd[AST_MAXL]; float* q; float* scores; float* alpha; float* cvh; float* htout; float* logits;
    BeamState bs; float* cand_lp; int* cand_tok;
    int* hist_parent; int* hist_tok; float* alpha_hist;
};
int beam_seq(cudaStream_t st, const BeamSeq& p);
int beam_step_batch(cudaStream_t st, int G, const float* z, int ldz, int V, int K, int N, const BeamState& bs, float* cand_lp,
                    int* cand_tok, int step, int eos, int stop_limit, int* hist_parent, int* hist_tok, const BeamGather& gd, int Tp_ld,
                    const float* alpha_step, float* alpha_hist, const int* last_tok_prev, int* last_tok_next);
int beam_all_done(cudaStream_t st, const int* done, int G, int* all_done);
int beam_topk(cudaStream_t st, const float* z, int ldz, int V, int K, int N, const BeamState& bs, float* cand_lp, int* cand_tok);
int beam_prune(cudaStream_t st, const BeamState& bs, const float* cand_lp, const int* cand_tok, int N, int K, int step,
               int eos, int* hist_parent, int* hist_tok);
int beam_gather(cudaStream_t st, const BeamState& bs, const BeamGather& gd, int N, int step, int Tp,
                const float* alpha_step, float* alpha_hist, const int* last_tok_prev, int* last_tok_next);

}  // namespace ast

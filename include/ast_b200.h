/* ast_b200 — C ABI of the B200-native hot path of 0xSameer/ast.
 *
 * The reference has no FFI layer (pure Python on Chainer/CuPy); its boundary is the
 * Python object protocol between nn.py / beam.py and seq2seq.py::SpeechEncoderDecoder.  Each entry
 * point below names the reference interface it replaces (file:line under /root/reference).  The
 * Python drop-in (ast_b200/seq2seq.py, ast_b200/nn.py) binds these through ctypes; INTEGRATION.md
 * shows the stub.
 *
 * Conventions: every function returns 0 on success, <0 on error (ast_last_error() gives the text);
 * nothing throws across the ABI.  All tensors are caller-owned DEVICE pointers, fp32 row-major unless
 * noted; `stream` is a cudaStream_t passed as void*.  No hidden allocation: parameters, gradients,
 * optimizer moments and the workspace are caller-owned buffers whose sizes are queried first.
 * One ast_model per (device, stream); not thread-safe.
 */
#ifndef AST_B200_H
#define AST_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ast_model ast_model;

/* model_cfg.json (+ config.py:25 dec_vocab_size) flattened; seq2seq.py:23-156 */
typedef struct ast_config {
    int feat_dim;                 /* D: 13 (shipped MFCC configs) or 40 (fbank) */
    int cnn_cout[2], cnn_kh[2], cnn_kw[2], cnn_sh[2], cnn_sw[2], cnn_ph[2], cnn_pw[2];
    int enc_layers, dec_layers;
    int hidden_units, embedding_units, attn_units, vocab;
    float drop_embed, drop_rnn, drop_out;
} ast_config;

const char* ast_last_error(void);
int ast_abi_version(void);
/* number of CUDA kernels this library has enqueued so far (optionally reset) */
unsigned long long ast_launch_count(int reset);

/* SpeechEncoderDecoder.__init__ (seq2seq.py:23-33); to_gpu (nn.py:133) is the `device` argument. */
int ast_create(const ast_config* cfg, int device, ast_model** out);
int ast_destroy(ast_model* m);

/* Flat parameter buffer layout.  Tensors keep chainer's shapes and serializer names (Appendix A.9:
 * "CNN_0/W", "L0_enc/upward/W", ...), each starting at a 256-byte aligned offset.  Used by
 * serializers.load_npz/save_npz (nn.py:150, train.py:75) and copy_params.py:26-43. */
long long ast_param_floats(const ast_model* m);
int ast_param_count(const ast_model* m);
int ast_param_info(const ast_model* m, int idx, char* name, int name_cap, long long* offset, int* ndim, int* shape4);
/* params/grads: ast_param_floats() floats each; bn_state: per CNN layer avg_mean[C] then avg_var[C]. */
int ast_bn_state_floats(const ast_model* m);
int ast_bind_params(ast_model* m, float* params, float* grads, float* bn_state);
/* Must be called after params change outside ast_opt_step (init, load_npz, copy_params). */
int ast_weights_changed(ast_model* m);

/* Workspace for batches up to (B, T frames, L target tokens) and beam width N. */
long long ast_workspace_bytes(const ast_model* m, int B, int T, int L, int beam_n, int max_steps);
int ast_bind_workspace(ast_model* m, void* ws, long long bytes, int B, int T, int L, int beam_n, int max_steps);

/* options: "exact" (1: fp32-faithful 3xTF32 / fp32 FMA everywhere, 0: single-pass TF32 tensor cores, FP16-operand encoder recurrences),
 *          "seed" (dropout / noise RNG), "tc_gemm" (1: tcgen05 GEMM for the batched contractions).
 * Schedule of the persistent encoder wavefront (defaults from tools/sweep_sched.sh on the benchmarked step):
 *          "enc_pchunk" (steps per hand-off chunk, 4), "enc_gemm_ctas" / "enc_gemm_ctas_bwd" (CTAs of each gated projection /
 *          data-gradient GEMM, 8 / 4), "enc_l0dx_ctas" (> 0: layer 0's data gradient as a gated GEMM beside the recurrences, 0),
 *          "enc_side_ctas" (CTA cap of the side stream's GEMMs beside the backward wavefront; 0 = every SM it leaves free).
 *          The same five can be preset from the environment: AST_ENC_PCHUNK, AST_ENC_GEMM_CTAS, AST_ENC_GEMM_CTAS_BWD,
 *          AST_ENC_L0DX_CTAS, AST_ENC_SIDE_CTAS (read at ast_create). */
int ast_set_option(ast_model* m, const char* key, double value);
double ast_get_option(const ast_model* m, const char* key);

/* encode (seq2seq.py:293-315): X (B,T,D).  train!=0: BN batch statistics + running-stat update,
 * dropout and optional multiplicative input noise (explicit tensor or N(1,sigma) generated on device).
 * Sets the model's enc_states (B,T',H) and the encoder link states. */
int ast_encode(ast_model* m, const float* X, int B, int T, int train, const float* noise, float noise_sigma,
               void* stream);
int ast_enc_len(const ast_model* m, int T);                      /* T' for T input frames */
int ast_get_enc_states(ast_model* m, float* out, void* stream); /* (B,T',H) */

/* forward_loss (seq2seq.py:399-473; nn.py:175-179): y (B,L) int32, use_true (L-1 bytes, device, may be
 * NULL = teach_ratio 1: the host's random.random() draws of :431-436), loss_out: 1 float (device).
 * Also init_decoder_state (:318-334), decode_step x (L-1) and the PAD-weighted softmax-CE (:468). */
int ast_forward_loss(ast_model* m, const float* X, const int* y, int B, int T, int L,
                     const unsigned char* use_true, const float* noise, float noise_sigma, float* loss_out,
                     void* stream);
/* loss.backward() (nn.py:181) after cleargrads (:180): fills the bound grads buffer. */
int ast_backward(ast_model* m, void* stream);
/* Data-parallel hook (no reference counterpart: the reference is single-GPU).  backward completes the flat
 * gradient buffer in three contiguous buckets - 0: decoder (attn_Wa .. out), 1: encoder stacks, 2: CNN - and records an
 * event as each one becomes final.  ast_grad_bucket_wait makes `stream` wait for that event of the LAST enqueued
 * ast_backward, so a caller can all-reduce bucket 0 while the encoder/CNN backward still runs.  range: in floats. */
int ast_grad_bucket_count(const ast_model* m);
int ast_grad_bucket_range(const ast_model* m, int bucket, long long* offset, long long* count);
int ast_grad_bucket_wait(ast_model* m, int bucket, void* stream);
/* per-step argmax tokens of the last forward_loss: (L-1, B) int32 */
int ast_get_step_argmax(ast_model* m, int* out, void* stream);

/* optimizer.update() (nn.py:81-119,182): WeightDecay -> GradientClipping(global L2) -> AMSGrad.
 * moments m1,v,vhat: ast_param_floats() floats each; grad_scale multiplies grads first (1/world_size
 * after a data-parallel sum all-reduce); t = 1-based update count; frozen tensors by index list. */
int ast_opt_step(ast_model* m, float* m1, float* v, float* vhat, int t, float lr, float l2, float clip,
                 float beta1, float beta2, float eps, float grad_scale, const int* frozen_idx, int n_frozen,
                 void* stream);
/* optimizers.SGD(lr) (nn.py:91-93, optimizer.type = 1) behind the same hooks: WeightDecay -> GradientClipping(global L2) -> p -= lr * g.
 * The GradientNoise hook (nn.py:107-110, grad_noise_eta > 0) applies to both update rules: ast_set_option(m, "grad_noise_sigma", s)
 * sets the standard deviation of the N(0, s^2) noise added to every (clipped) gradient element at the NEXT update
 * (Chainer: s = sqrt(eta / (1 + t)^0.55), t = updates made so far); device counter RNG, a fresh stream per update. */
int ast_opt_step_sgd(ast_model* m, float lr, float l2, float clip, float grad_scale, const int* frozen_idx, int n_frozen,
                     void* stream);
/* re-record all bucket events on `stream`: the gradients were modified after ast_backward (ast_scale_grads) or no backward ran this
 * step (empty shard), so a collective gated on ast_grad_bucket_wait must wait for work enqueued on `stream` up to here */
int ast_grad_buckets_mark(ast_model* m, void* stream);
/* data parallelism (new work): grads *= weight in place before the all-reduce - a rank whose shard of a tail batch is
 * smaller than the others' weights its replica-mean gradient by n_local * world / n_global (0 for an empty shard: exact zeros) */
int ast_scale_grads(ast_model* m, float weight, void* stream);
double ast_last_grad_norm(ast_model* m, void* stream);   /* synchronises; for logging / tests */

/* decode_step / get|set_decoder_states / get_encoder_states (seq2seq.py:361-396, 529-569; nn.py:238-274).
 * Decoder state lives in the model: per layer c (Bd,H) and h (Bd,H).  states layout: [layer][c|h][Bd][H]. */
int ast_init_decoder_state(ast_model* m, int Bd, void* stream);     /* from encoder finals; Bd = B, or broadcast */
int ast_get_encoder_states(ast_model* m, float* states, void* stream);
int ast_get_decoder_states(ast_model* m, float* states, int Bd, void* stream);
int ast_set_decoder_states(ast_model* m, const float* states, int Bd, void* stream);
/* word (Bd) int32, ht_in (Bd,A) -> logits (Bd,V), ht_out (Bd,A), alphas (Bd,T'). Eval mode. */
int ast_decode_step(ast_model* m, const int* word, const float* ht_in, int Bd, float* logits, float* ht_out,
                    float* alphas, void* stream);

/* compute_context_vector(dec_h, attn_Wa) (seq2seq.py:336-358; called by decode_step :379,382): dec_h (Bd,H) -> cv (Bd,H),
 * alphas (Bd,T') (may be NULL) over the encoder states of the last encode.  Wa (H,H) / ba (H): device pointers to the attention
 * link's W and b (the `attn_Wa` argument of the reference), both NULL = the model's own attn_Wa. */
int ast_attention(ast_model* m, const float* dec_h, int Bd, const float* Wa, const float* ba, float* cv, float* alphas,
                  void* stream);

/* predict (seq2seq.py:475-527; nn.py:217-220): greedy batched decode; preds (stop_limit,B) int32 device,
 * n_steps (host) = rows actually produced (not cut at EOS per row, as the reference). */
int ast_predict(ast_model* m, const float* X, int B, int T, int start_token, int end_token, int stop_limit,
                int* preds, int* n_steps, void* stream);

/* decode_beam (nn.py:235-322; beam.py:117): one utterance X (1,T,D), N kept hyps, K expansions each.
 * Outputs (host pointers): n_steps, n_hyps; device outputs: hist_parent/hist_tok (stop_limit,N) int32,
 * scores (N) f32, alpha_hist (stop_limit,N,T') f32 (may be NULL), final states (see get_decoder_states)
 * and attn_v (N,A) (may be NULL). */
int ast_beam_search(ast_model* m, const float* X, int T, int stop_limit, int N, int K, int go_token, int eos_token,
                    int* n_steps, int* n_hyps, int* hist_parent, int* hist_tok, float* scores, float* alpha_hist,
                    float* final_states, float* final_attn_v, void* stream);

/* The beam.py:110-124 loop over utterances as ONE lock-step search (throughput mode; new work - the reference decodes one utterance
 * at a time): G <= 32 utterances, X = their features back to back (sum of lens[g] x D floats, device), lens (host) in frames.
 * Same hypotheses, scores and attention history per utterance as ast_beam_search on it alone; rows of different utterances never mix
 * (one decoder pass over rows = G x N, grouped attention with per-utterance length, per-utterance top-K / prune / gather).
 * Outputs are utterance-major: n_steps / n_hyps / enc_lens_out (host, G), hist_parent / hist_tok (G, stop_limit, N), scores (G, N),
 * alpha_hist (G, stop_limit, N, Tp_ld) with Tp_ld >= the largest T', final_states (2*layers, G*N, H) and final_attn_v (G*N, A) or NULL.
 * Workspace: ast_bind_workspace with B >= the largest equal-length run (or 1), T >= max lens, N >= G * N, steps >= stop_limit. */
int ast_beam_search_batch(ast_model* m, const float* X, const int* lens, int G, int stop_limit, int N, int K, int go_token,
                          int eos_token, int* n_steps, int* n_hyps, int* enc_lens_out, int* hist_parent, int* hist_tok, float* scores,
                          float* alpha_hist, int Tp_ld, float* final_states, float* final_attn_v, void* stream);

/* ---- stateless kernels (unit-testable pieces; also what a foreign host would call directly) ---- */
/* Kaldi apply-cmvn + dataloader.py:103,156 pad_sequence (+ :83-93 frame zeroing, seq2seq.py:300 noise) */
int ast_pack_cmvn(const float* raw, const long long* row_off, const int* lens, const float* scale,
                  const float* offset, const unsigned char* keep, const float* noise, float noise_sigma,
                  unsigned long long seed, float* X, int B, int T, int D, void* stream);
/* F.softmax_cross_entropy(class_weight=mask_pad_id) fwd+bwd+argmax (seq2seq.py:448,468) */
int ast_softmax_ce(float* logits_inout, int ld, const int* targets, int B, int V, float* row_loss, int* argmax,
                   void* stream);
/* C = alpha*op(A)op(B) + beta*C + bias ; which: 0 = fp32 SIMT, 1 = tcgen05 TF32 (NT only) */
int ast_gemm(int which, int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
             int ldb, float beta, float* C, int ldc, const float* bias, void* stream);
/* test hook: n (1..4) same-shape problems in one 2-CTA launch; problem g uses A + g*strideA, B + g*strideB, C + g*strideC;
 * split_k: 0 none, -1 automatic, > 0 count */
int ast_gemm_grouped(int n, int ta, int tb, int M, int N, int K, const float* A, long long strideA, int lda, const float* B,
                     long long strideB, int ldb, float* C, long long strideC, int ldc, int split_k, void* stream);
/* fp32-faithful tensor-core GEMM, C = A . B^T + bias (A: M x K, B: N x K, both k-contiguous): 3xTF32 on tcgen05
   (hi.hi + lo.hi + hi.lo with hi = rna_tf32(v), lo = rna_tf32(v - hi)).  Ahi/Alo, Bhi/Blo are scratch buffers of A's / B's
   size, filled by the call (ast_split_tf32).  Replaces the cuDNN convolution forward of CNN_1 (seq2seq.py:165) as an
   implicit GEMM over overlapping rows (lda < K). */
int ast_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream);
int ast_gemm3_nt(int M, int N, int K, const float* A, float* Ahi, float* Alo, long long a_floats, int lda, const float* B,
                 float* Bhi, float* Blo, long long b_floats, int ldb, float* C, int ldc, const float* bias, void* stream);
/* persistent LSTM recurrence over a sequence, one chain (kernel-level test hook) */
int ast_lstm_seq(int backward, float* G, const float* Wl, float* Hs, float* Cs, float* out_or_dout, int T, int B, int h,
                 const float* dh_fin, const float* dc_fin, int exact, void* stream);

/* diagnostics: device buffer (>= 512 uint64) that receives 32-bit clock stamps of steps 8..23 of the tcgen05 LSTM kernels:
   forward [0, 128) = 16 steps x 8 slots; backward [128, 384) = 16 steps x (8 slots of epilogue thread 0, 8 of thread 224),
   [384, 416) = 16 steps x 2 slots of the MMA issuer (tools/enc_step_probe.py decodes them); NULL switches the probe off */
int ast_lstm_probe(unsigned long long* dev_buf);
/* diagnostics: with ast_set_option(m, "stage_timing", 1), milliseconds between the stage marks (CNN / encoder / decoder ...) of the
   last forward_loss + backward on the caller's stream; synchronises; returns the number of intervals */
int ast_stage_times(ast_model* m, float* ms, char* names, int name_stride, int cap);
/* test hook: copy a named internal buffer ("raw0", "rnn_in", "G_00", "H_21", "d_enc", ...) */
int ast_debug_fetch(ast_model* m, const char* name, float* out, long long max_floats, long long* n_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AST_B200_H */

// Per-step decoder kernels (seq2seq.py:336-396, 468): batch-as-M "skinny" tensor-core GEMM with
// fused epilogues (bias / tanh / LSTM cell / tanh-backward / cell-backward), Luong attention forward+backward,
// fused softmax-cross-entropy forward+backward+argmax, embedding gather/scatter and the LSTM cell backward.
// These are the one-kernel-per-op versions used by the decode_step protocol, greedy / beam decoding and as the
// cross-check for the persistent decoder-sequence kernels in dec_seq.cu; the arithmetic lives in decoder_dev.cuh.
//
// The decoder is strictly sequential (input feeding + scheduled sampling), its weights (31.6 MB fp32) live in
// B200's 126 MB L2 across steps, and every GEMM has M = batch (16..32): the bound is L2 weight streaming +
// dependency latency, not tensor throughput.  mma.sync m16n8k8 TF32 (3-term split = fp32 accuracy) matches
// M = 16 exactly; a tcgen05 tile (M >= 64) would be >= 75 % padding.
#include "decoder_dev.cuh"

namespace ast {

template <int MT, bool EXACT>
__global__ void __launch_bounds__(SK_THREADS)
skinny_kernel(SkinnyArgs p) {
    __shared__ SkinnySmem sm;
    if (blockIdx.y > 0) {
        // more than 32 rows (beam search batched over utterances: rows = utterances x hypotheses): this CTA's 32-row block.
        // Rows are independent, so a block is the same problem with every per-row pointer moved down by r0 rows.
        const size_t r0 = (size_t)blockIdx.y * 32;
        p.B = min(32, p.B - (int)r0);
        p.X[0] += r0 * p.ldx[0];
        if (p.K[1] > 0) p.X[1] += r0 * p.ldx[1];
        if (p.Y) p.Y += r0 * p.ldy;
        if (p.Y2) p.Y2 += r0 * p.ldy2;
        if (p.add) p.add += r0 * p.ld_add;
        if (p.aux) p.aux += r0 * p.ld_aux;
        if (p.epi == EPI_LSTM) {
            const size_t Hh = (size_t)(p.N >> 2);
            p.c_prev += r0 * Hh; p.c_out += r0 * Hh; p.h_out += r0 * Hh; p.hd_out += r0 * p.ld_hd;
            p.drop_base += r0 * Hh;
        }
    } else if (p.B > 32) {
        p.B = 32;
    }
    skinny_tile<MT, EXACT>(p, blockIdx.x * SK_COLS, sm);
}

int skinny(cudaStream_t st, const SkinnyArgs& p, bool exact) {
    AST_CHECK(p.B >= 1, "skinny: batch %d unsupported", p.B);
    if (p.B > 32) {
        AST_CHECK(p.epi == EPI_NONE || p.epi == EPI_TANH || p.epi == EPI_LSTM, "skinny: more than 32 rows only for the forward epilogues");
        AST_CHECK(!p.sc_demb, "skinny: more than 32 rows not supported with the embedding scatter");
    }
    AST_CHECK(p.K[0] % 16 == 0 && p.K[1] % 16 == 0, "skinny: K (%d,%d) must be multiples of 16", p.K[0], p.K[1]);
    AST_CHECK(p.ldx[0] % 4 == 0 && p.ldw[0] % 4 == 0 && (p.K[1] == 0 || (p.ldx[1] % 4 == 0 && p.ldw[1] % 4 == 0)),
              "skinny: leading dims must be multiples of 4");
    if (p.epi == EPI_LSTM) AST_CHECK(p.N % 16 == 0, "skinny: LSTM epilogue needs N %% 16 == 0");
    const dim3 grid(cdiv(p.N, SK_COLS), cdiv(p.B, 32));
    if (p.B <= 16) {
        if (exact) skinny_kernel<1, true><<<grid, SK_THREADS, 0, st>>>(p);
        else skinny_kernel<1, false><<<grid, SK_THREADS, 0, st>>>(p);
    } else {
        if (exact) skinny_kernel<2, true><<<grid, SK_THREADS, 0, st>>>(p);
        else skinny_kernel<2, false><<<grid, SK_THREADS, 0, st>>>(p);
    }
    AST_LAUNCH_OK();
    return 0;
}

// =========================================================================================
// embedding gather + input feeding concat (seq2seq.py:365-372) with on-device scheduled sampling
// =========================================================================================
// word = use_true[step] ? y[b][step] : prev_argmax[b] ; x0[b] = [ E[word]*dropmask ; ht_prev[b] ]
__global__ void embed_concat_kernel(const float* __restrict__ emb, const int* __restrict__ y, int ldy_tok,
                                    const unsigned char* __restrict__ use_true, const int* __restrict__ prev_argmax,
                                    const int* __restrict__ forced_words, const float* __restrict__ ht_prev, int ld_ht,
                                    float* __restrict__ x0, int* __restrict__ words_used, int B, int E, int A, int V,
                                    int step, float drop, unsigned long long seed, unsigned drop_stream) {
    const int b = blockIdx.x;
    int word;
    if (forced_words) word = forced_words[b];
    else word = (use_true == nullptr || use_true[step] || prev_argmax == nullptr) ? y[(size_t)b * ldy_tok + step] : prev_argmax[b];
    word = min(max(word, 0), V - 1);
    if (threadIdx.x == 0 && words_used) words_used[b] = word;
    float* dst = x0 + (size_t)b * (E + A);
    for (int j = threadIdx.x; j < E; j += blockDim.x) {
        const float dm = dropout_scale(seed, drop_stream, (uint32_t)(((size_t)step * B + b) * E + j), drop);
        dst[j] = emb[(size_t)word * E + j] * dm;
    }
    for (int j = threadIdx.x; j < A; j += blockDim.x) dst[E + j] = ht_prev ? ht_prev[(size_t)b * ld_ht + j] : 0.f;
}

int embed_concat(cudaStream_t st, const float* emb, const int* y, int ldy_tok, const unsigned char* use_true,
                 const int* prev_argmax, const int* forced_words, const float* ht_prev, int ld_ht, float* x0,
                 int* words_used, int B, int E, int A, int V, int step, float drop, unsigned long long seed,
                 unsigned drop_stream) {
    embed_concat_kernel<<<B, 128, 0, st>>>(emb, y, ldy_tok, use_true, prev_argmax, forced_words, ht_prev, ld_ht, x0,
                                           words_used, B, E, A, V, step, drop, seed, drop_stream);
    AST_LAUNCH_OK();
    return 0;
}

// dEmb[word[b]][j] += dx0[b][j] * dropmask   (EmbedID backward: duplicates accumulate)
__global__ void embed_scatter_kernel(float* __restrict__ demb, const float* __restrict__ dx0, int ld_dx,
                                     const int* __restrict__ words, int B, int E, int step, float drop,
                                     unsigned long long seed, unsigned drop_stream) {
    const int b = blockIdx.x;
    const int word = words[b];
    for (int j = threadIdx.x; j < E; j += blockDim.x) {
        const float dm = dropout_scale(seed, drop_stream, (uint32_t)(((size_t)step * B + b) * E + j), drop);
        atomicAdd(&demb[(size_t)word * E + j], dx0[(size_t)b * ld_dx + j] * dm);
    }
}
int embed_scatter(cudaStream_t st, float* demb, const float* dx0, int ld_dx, const int* words, int B, int E, int step,
                  float drop, unsigned long long seed, unsigned drop_stream) {
    embed_scatter_kernel<<<B, 128, 0, st>>>(demb, dx0, ld_dx, words, B, E, step, drop, seed, drop_stream);
    AST_LAUNCH_OK();
    return 0;
}

// =========================================================================================
// attention (seq2seq.py:336-358), no length mask (the reference's is commented out, :344-347)
// =========================================================================================
__global__ void __launch_bounds__(256) attn_dot_kernel(const float* __restrict__ enc, long long enc_bs, const float* v, int ldv,
                                                       float* s, int Tp, int H) {
    attn_dot_block(enc, enc_bs, v, ldv, s, Tp, H, blockIdx.y, blockIdx.x);
}
int attn_dot(cudaStream_t st, const float* enc, long long enc_bs, const float* v, int ldv, float* s, int B, int Tp, int H) {
    AST_CHECK(H % 4 == 0 && ldv % 4 == 0, "attn_dot: H/ldv must be multiples of 4");
    dim3 grid(cdiv(Tp, 8), B);
    attn_dot_kernel<<<grid, 256, 0, st>>>(enc, enc_bs, v, ldv, s, Tp, H);
    AST_LAUNCH_OK();
    return 0;
}

__global__ void __launch_bounds__(256) attn_ctx_kernel(const float* __restrict__ enc, long long enc_bs, const float* s, float* alpha,
                                                       float* cv, int ld_cv, int Tp, int H) {
    extern __shared__ float dsm[];          // Tp alphas + 8*128 partials
    __shared__ float scratch[32];
    attn_ctx_block(enc, enc_bs, s, alpha, cv, ld_cv, Tp, H, blockIdx.y, blockIdx.x, dsm, dsm + ((Tp + 3) & ~3), scratch);
}
int attn_ctx(cudaStream_t st, const float* enc, long long enc_bs, const float* s, float* alpha, float* cv, int ld_cv,
             int B, int Tp, int H) {
    AST_CHECK(H % 4 == 0 && ld_cv % 4 == 0, "attn_ctx: H/ld must be multiples of 4");
    dim3 grid(cdiv(H, 128), B);
    attn_ctx_kernel<<<grid, 256, sizeof(float) * (Tp + 4 + 8 * 128), st>>>(enc, enc_bs, s, alpha, cv, ld_cv, Tp, H);
    AST_LAUNCH_OK();
    return 0;
}

// Grouped attention (beam search batched over utterances): row b attends over the encoder states of utterance b / rows_per_enc,
// which has its OWN length lens[g] <= Tp_ld (utterances of different lengths share one launch; scores / alpha rows are Tp_ld apart
// and the tail [lens[g], Tp_ld) of alpha is zeroed).  Same arithmetic per row as attn_dot / attn_ctx on that utterance alone.
__global__ void __launch_bounds__(256) attn_dot_grouped_kernel(const float* __restrict__ enc, long long enc_bs, int rows_per_enc,
                                                               const int* __restrict__ lens, const float* v, int ldv, float* s,
                                                               int Tp_ld, int H) {
    const int b = blockIdx.y, g = b / rows_per_enc;
    const int Tg = lens[g];
    if ((int)blockIdx.x * 8 >= Tg) return;
    attn_dot_block(enc + (size_t)g * enc_bs, 0, v + (size_t)b * ldv, ldv, s + (size_t)b * Tp_ld, Tg, H, 0, blockIdx.x);
}
__global__ void __launch_bounds__(256) attn_ctx_grouped_kernel(const float* __restrict__ enc, long long enc_bs, int rows_per_enc,
                                                               const int* __restrict__ lens, const float* s, float* alpha, float* cv,
                                                               int ld_cv, int Tp_ld, int H) {
    extern __shared__ float dsm[];
    __shared__ float scratch[32];
    const int b = blockIdx.y, g = b / rows_per_enc;
    const int Tg = lens[g];
    attn_ctx_block(enc + (size_t)g * enc_bs, 0, s + (size_t)b * Tp_ld, alpha + (size_t)b * Tp_ld, cv + (size_t)b * ld_cv, ld_cv, Tg, H, 0,
                   blockIdx.x, dsm, dsm + ((Tg + 3) & ~3), scratch);
    if (blockIdx.x == 0)
        for (int t = Tg + threadIdx.x; t < Tp_ld; t += blockDim.x) alpha[(size_t)b * Tp_ld + t] = 0.f;
}
int attn_grouped(cudaStream_t st, const float* enc, long long enc_bs, int rows_per_enc, const int* lens, const float* q, int ldq,
                 float* scores, float* alpha, float* cv, int ld_cv, int rows, int Tp_ld, int H) {
    AST_CHECK(H % 4 == 0 && ldq % 4 == 0 && ld_cv % 4 == 0, "attn_grouped: H/ld must be multiples of 4");
    attn_dot_grouped_kernel<<<dim3(cdiv(Tp_ld, 8), rows), 256, 0, st>>>(enc, enc_bs, rows_per_enc, lens, q, ldq, scores, Tp_ld, H);
    AST_LAUNCH_OK();
    attn_ctx_grouped_kernel<<<dim3(cdiv(H, 128), rows), 256, sizeof(float) * (Tp_ld + 4 + 8 * 128), st>>>(enc, enc_bs, rows_per_enc, lens,
                                                                                                      scores, alpha, cv, ld_cv, Tp_ld, H);
    AST_LAUNCH_OK();
    return 0;
}

__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ enc, float* d_enc, long long enc_bs, const float* alpha,
                                                       const float* dalpha, const float* dcv, int ld_dcv, const float* qv, int ld_q,
                                                       float* dq, int ld_dq, int Tp, int H) {
    extern __shared__ float dsm[];          // alpha[Tp], ds[Tp], 8*128 partials
    __shared__ float scratch[32];
    attn_bwd_block(enc, d_enc, enc_bs, alpha, dalpha, dcv, ld_dcv, qv, ld_q, dq, ld_dq, Tp, H, blockIdx.y, blockIdx.x, dsm,
                   dsm + ((2 * Tp + 3) & ~3), scratch);
}
int attn_bwd(cudaStream_t st, const float* enc, float* d_enc, long long enc_bs, const float* alpha, const float* dalpha,
             const float* dcv, int ld_dcv, const float* qv, int ld_q, float* dq, int ld_dq, int B, int Tp, int H) {
    AST_CHECK(H % 4 == 0 && ld_dcv % 4 == 0 && ld_q % 4 == 0, "attn_bwd: H/ld must be multiples of 4");
    dim3 grid(cdiv(H, 128), B);
    attn_bwd_kernel<<<grid, 256, sizeof(float) * (2 * Tp + 4 + 8 * 128), st>>>(enc, d_enc, enc_bs, alpha, dalpha, dcv, ld_dcv, qv, ld_q,
                                                                         dq, ld_dq, Tp, H);
    AST_LAUNCH_OK();
    return 0;
}

// =========================================================================================
// fused softmax cross-entropy forward + backward + argmax (seq2seq.py:448,468; Appendix A.7)
// =========================================================================================
// One CTA per batch row.  row_loss[b] = -w[t]*logp[t]/B ; z <- (softmax(z) - onehot(t)) * w[t]/B in place
// (cols [V,ldz) zeroed) ; argmax[b] = lowest index of the row maximum.  write_grad = 0 -> argmax only.
__global__ void __launch_bounds__(256) softmax_ce_kernel(float* __restrict__ z, int ldz, const int* __restrict__ y, int ldy_tok,
                                                         int step_next, float* __restrict__ row_loss, int* __restrict__ argmax_out,
                                                         int B, int V, int write_grad) {
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    const int b = blockIdx.x;
    const int t = write_grad ? y[(size_t)b * ldy_tok + step_next] : 0;
    const int mi = softmax_ce_row(z + (size_t)b * ldz, ldz, V, t, B, row_loss ? row_loss + b : nullptr, write_grad, scratch, iscratch);
    if (threadIdx.x == 0 && argmax_out) argmax_out[b] = mi;
}
int softmax_ce(cudaStream_t st, float* z, int ldz, const int* y, int ldy_tok, int step_next, float* row_loss,
               int* argmax_out, int B, int V, bool write_grad) {
    softmax_ce_kernel<<<B, 256, 0, st>>>(z, ldz, y, ldy_tok, step_next, row_loss, argmax_out, B, V, write_grad ? 1 : 0);
    AST_LAUNCH_OK();
    return 0;
}

// All decoder steps at once: CTA (s, b) handles logits row s*B + b against target y[b][s + 1].
__global__ void __launch_bounds__(256) softmax_ce_all_kernel(float* __restrict__ z, int ldz, const int* __restrict__ y, int ldy_tok,
                                                             float* __restrict__ row_loss, int* __restrict__ argmax_out, int B, int V) {
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    const int row = blockIdx.x, s = row / B, b = row - s * B;
    const int t = y[(size_t)b * ldy_tok + s + 1];
    const int mi = softmax_ce_row(z + (size_t)row * ldz, ldz, V, t, B, row_loss + row, 1, scratch, iscratch);
    if (threadIdx.x == 0 && argmax_out) argmax_out[row] = mi;
}
int softmax_ce_all(cudaStream_t st, float* z, int ldz, const int* y, int ldy_tok, float* row_loss, int* argmax_out, int S, int B, int V) {
    softmax_ce_all_kernel<<<S * B, 256, 0, st>>>(z, ldz, y, ldy_tok, row_loss, argmax_out, B, V);
    AST_LAUNCH_OK();
    return 0;
}

// loss = sum over (steps x B) row losses, fixed order, double accumulation, single CTA.
__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int n, float* __restrict__ loss) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += row_loss[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)sh[0];
}
int loss_reduce(cudaStream_t st, const float* row_loss, int n, float* loss) {
    loss_reduce_kernel<<<1, 256, 0, st>>>(row_loss, n, loss);
    AST_LAUNCH_OK();
    return 0;
}

// =========================================================================================
// LSTM cell backward (decoder, one step): dG in place over the saved activations
// =========================================================================================
__global__ void lstm_cell_bwd_kernel(float* __restrict__ act, const float* __restrict__ c, const float* __restrict__ c_prev,
                                     const float* __restrict__ d_out, int ld_dout, const float* __restrict__ dh_rec,
                                     int ld_dhrec, float* __restrict__ dc, int B, int H, int step_row0, float drop,
                                     unsigned long long seed, unsigned drop_stream) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx % H;
    const float dm = dropout_scale(seed, drop_stream, (uint32_t)((size_t)step_row0 * H + idx), drop);
    float dh = d_out[(size_t)b * ld_dout + j] * dm;
    if (dh_rec) dh += dh_rec[(size_t)b * ld_dhrec + j];
    const float4 a = *reinterpret_cast<const float4*>(act + (size_t)idx * 4);
    const float cc = c[idx], cp = c_prev[idx];
    const float tc = tanhf(cc);
    const float dct = dc[idx] + dh * a.w * (1.f - tc * tc);
    float4 dg;
    dg.x = dct * a.y * (1.f - a.x * a.x);
    dg.y = dct * a.x * a.y * (1.f - a.y);
    dg.z = dct * cp * a.z * (1.f - a.z);
    dg.w = dh * tc * a.w * (1.f - a.w);
    dc[idx] = dct * a.z;
    *reinterpret_cast<float4*>(act + (size_t)idx * 4) = dg;
}
int lstm_cell_bwd(cudaStream_t st, float* act, const float* c, const float* c_prev, const float* d_out, int ld_dout,
                  const float* dh_rec, int ld_dhrec, float* dc, int B, int H, int step_row0, float drop,
                  unsigned long long seed, unsigned drop_stream) {
    lstm_cell_bwd_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(act, c, c_prev, d_out, ld_dout, dh_rec, ld_dhrec, dc, B, H,
                                                            step_row0, drop, seed, drop_stream);
    AST_LAUNCH_OK();
    return 0;
}

// =========================================================================================
// pieces of the tensor-core decode step (beam search: rows = utterances x hypotheses, fp32-faithful 3xTF32 tcgen05 GEMMs)
// =========================================================================================
// out[r][0:w0] = src0[r][0:w0], out[r][w0:w0+w1] = src1[r][0:w1], written as the (hi, lo) TF32 split the 3xTF32 GEMM consumes
// (hi = rna_tf32(x), lo = rna_tf32(x - hi), same rounding as split_tf32).  One float4 per thread; widths are multiples of 4.
__global__ void split_concat2_kernel(const float* __restrict__ src0, int ld0, int w0, const float* __restrict__ src1, int ld1, int w1,
                                     float* __restrict__ hi, float* __restrict__ lo, int ldo, int R) {
    const int w4 = (w0 + w1) >> 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * w4; idx += gridDim.x * blockDim.x) {
        const int r = idx / w4, k = (idx - r * w4) << 2;
        const float4 v = k < w0 ? __ldcg(reinterpret_cast<const float4*>(src0 + (size_t)r * ld0 + k))
                                : __ldcg(reinterpret_cast<const float4*>(src1 + (size_t)r * ld1 + (k - w0)));
        float4 h, l;
        h.x = __uint_as_float(f2tf32(v.x)); h.y = __uint_as_float(f2tf32(v.y)); h.z = __uint_as_float(f2tf32(v.z)); h.w = __uint_as_float(f2tf32(v.w));
        l.x = __uint_as_float(f2tf32(v.x - h.x)); l.y = __uint_as_float(f2tf32(v.y - h.y));
        l.z = __uint_as_float(f2tf32(v.z - h.z)); l.w = __uint_as_float(f2tf32(v.w - h.w));
        *reinterpret_cast<float4*>(hi + (size_t)r * ldo + k) = h;
        *reinterpret_cast<float4*>(lo + (size_t)r * ldo + k) = l;
    }
}
int split_concat2(cudaStream_t st, const float* src0, int ld0, int w0, const float* src1, int ld1, int w1, float* hi, float* lo, int ldo, int R) {
    AST_CHECK(w0 % 4 == 0 && w1 % 4 == 0 && ld0 % 4 == 0 && (w1 == 0 || ld1 % 4 == 0) && ldo % 4 == 0, "split_concat2: widths / strides must be multiples of 4");
    const int n4 = R * ((w0 + w1) >> 2);
    split_concat2_kernel<<<std::max(1, std::min(cdiv(n4, 256), 148 * 8)), 256, 0, st>>>(src0, ld0, w0, src1, ld1, w1, hi, lo, ldo, R);
    AST_LAUNCH_OK();
    return 0;
}

// F.lstm on pre-activations that already include the bias (Appendix A.3, interleaved gates): act <- (a, i, f, o) in place,
// c = a i + f c_prev, h = o tanh(c); eval mode (no dropout): hd = h
__global__ void lstm_cell_rows_kernel(float* __restrict__ act, const float* __restrict__ c_prev, float* __restrict__ c_out,
                                      float* __restrict__ h_out, float* __restrict__ hd_out, int ld_hd, int R, int H) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * H) return;
    const int r = idx / H, j = idx - r * H;
    const float4 g = __ldcg(reinterpret_cast<const float4*>(act + ((size_t)r * H + j) * 4));
    const float ga = tanhf(g.x), gi = sigmoidf_(g.y), gf = sigmoidf_(g.z), go = sigmoidf_(g.w);
    const float c = ga * gi + gf * __ldcg(c_prev + (size_t)r * H + j);
    const float hv = go * tanhf(c);
    *reinterpret_cast<float4*>(act + ((size_t)r * H + j) * 4) = make_float4(ga, gi, gf, go);
    c_out[(size_t)r * H + j] = c;
    h_out[(size_t)r * H + j] = hv;
    hd_out[(size_t)r * ld_hd + j] = hv;
}
int lstm_cell_rows(cudaStream_t st, float* act, const float* c_prev, float* c_out, float* h_out, float* hd_out, int ld_hd, int R, int H) {
    lstm_cell_rows_kernel<<<cdiv(R * H, 256), 256, 0, st>>>(act, c_prev, c_out, h_out, hd_out, ld_hd, R, H);
    AST_LAUNCH_OK();
    return 0;
}

__global__ void tanh_rows_kernel(float* __restrict__ x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = tanhf(x[i]);
}
int tanh_rows(cudaStream_t st, float* x, size_t n) {
    tanh_rows_kernel<<<(int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 148 * 8)), 256, 0, st>>>(x, n);
    AST_LAUNCH_OK();
    return 0;
}

}  // namespace ast

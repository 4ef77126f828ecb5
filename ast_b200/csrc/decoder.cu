// Per-step decoder kernels (seq2seq.py:336-396, 468): batch-as-M "skinny" tensor-core GEMM with
// fused epilogues (bias / tanh / LSTM cell / tanh-backward), Luong attention forward+backward with
// warp-shuffle reductions, fused softmax-cross-entropy forward+backward+argmax, embedding
// gather/scatter and the LSTM cell backward.
//
// The decoder is strictly sequential (input feeding + scheduled sampling), its weights (31.6 MB
// fp32) live in B200's 126 MB L2 across steps, and every GEMM has M = batch (16..32): the bound is
// L2 weight streaming + launch latency, not tensor throughput.  mma.sync m16n8k8 TF32 (3-term
// split = fp32 accuracy) matches M = 16 exactly; a tcgen05 tile (M >= 64) would be >= 75 % padding.
#include "common.cuh"
#include "kernels.h"

namespace ast {

// =========================================================================================
// skinny GEMM:  Y[b][n] = epi( sum_seg sum_k X_seg[b][k] * W_seg[n][k]  + bias[n] )
// CTA = 8 warps = 16 output columns (two n-tiles), K split over the warps in chunks of 16,
// cross-warp reduction through smem, then the epilogue.
// =========================================================================================
constexpr int SK_THREADS = 256;
constexpr int SK_COLS = 16;

template <int MT, bool EXACT>
__global__ void __launch_bounds__(SK_THREADS)
skinny_kernel(SkinnyArgs p) {
    constexpr int MROWS = 16 * MT;
    __shared__ float red[8][MROWS][SK_COLS + 1];
    __shared__ float outv[MROWS][SK_COLS + 1];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
    const int n0 = blockIdx.x * SK_COLS;

    float acc[2][MT][4];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][mt][j] = 0.f;

    const int nch0 = p.K[0] >> 4, nch1 = p.K[1] >> 4;
    const int nn0 = n0 + g, nn1 = n0 + 8 + g;
#pragma unroll 2
    for (int c = w; c < nch0 + nch1; c += 8) {
        const int seg = c < nch0 ? 0 : 1;
        const int k = ((seg ? c - nch0 : c) << 4) + 4 * q;
        const float* __restrict__ W = p.W[seg];
        const float* __restrict__ X = p.X[seg];
        const int ldw = p.ldw[seg], ldx = p.ldx[seg];
        float4 wv0 = make_float4(0.f, 0.f, 0.f, 0.f), wv1 = wv0;
        if (nn0 < p.N) wv0 = *reinterpret_cast<const float4*>(W + (size_t)nn0 * ldw + k);
        if (nn1 < p.N) wv1 = *reinterpret_cast<const float4*>(W + (size_t)nn1 * ldw + k);
        float4 xv[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int row = mt * 16 + g + 8 * hf;
                xv[mt][hf] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < p.B) xv[mt][hf] = *reinterpret_cast<const float4*>(X + (size_t)row * ldx + k);
            }
        // two k-steps; lane q supplies physical k = 4q+{0,1} then 4q+{2,3} for both operands
        const float b0a[2] = {wv0.x, wv0.y}, b0b[2] = {wv0.z, wv0.w};
        const float b1a[2] = {wv1.x, wv1.y}, b1b[2] = {wv1.z, wv1.w};
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const float aa[4] = {xv[mt][0].x, xv[mt][1].x, xv[mt][0].y, xv[mt][1].y};
            const float ab[4] = {xv[mt][0].z, xv[mt][1].z, xv[mt][0].w, xv[mt][1].w};
            mma_f32<EXACT>(acc[0][mt], aa, b0a);
            mma_f32<EXACT>(acc[0][mt], ab, b0b);
            mma_f32<EXACT>(acc[1][mt], aa, b1a);
            mma_f32<EXACT>(acc[1][mt], ab, b1b);
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            red[w][mt * 16 + g][s * 8 + 2 * q] = acc[s][mt][0];
            red[w][mt * 16 + g][s * 8 + 2 * q + 1] = acc[s][mt][1];
            red[w][mt * 16 + g + 8][s * 8 + 2 * q] = acc[s][mt][2];
            red[w][mt * 16 + g + 8][s * 8 + 2 * q + 1] = acc[s][mt][3];
        }
    __syncthreads();
    for (int idx = tid; idx < MROWS * SK_COLS; idx += SK_THREADS) {
        const int row = idx >> 4, col = idx & 15;
        float v = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) v += red[ww][row][col];
        const int n = n0 + col;
        if (p.bias && n < p.N) v += p.bias[n];
        outv[row][col] = v;
    }
    __syncthreads();

    if (p.epi == EPI_LSTM) {
        // 16 columns = 4 hidden units x (a,i,f,o)
        for (int idx = tid; idx < MROWS * 4; idx += SK_THREADS) {
            const int row = idx >> 2, ul = idx & 3;
            const int unit = (n0 >> 2) + ul;
            if (row < p.B && 4 * unit < p.N) {
                const int Hh = p.N >> 2;
                const float ga = tanhf(outv[row][4 * ul]), gi = sigmoidf_(outv[row][4 * ul + 1]);
                const float gf = sigmoidf_(outv[row][4 * ul + 2]), go = sigmoidf_(outv[row][4 * ul + 3]);
                const float c = ga * gi + gf * p.c_prev[(size_t)row * Hh + unit];
                const float hv = go * tanhf(c);
                *reinterpret_cast<float4*>(p.Y + (size_t)row * p.ldy + 4 * unit) = make_float4(ga, gi, gf, go);
                p.c_out[(size_t)row * Hh + unit] = c;
                p.h_out[(size_t)row * Hh + unit] = hv;
                const float dm = dropout_scale(p.seed, p.drop_stream, (uint32_t)(p.drop_base + (size_t)row * Hh + unit), p.drop);
                p.hd_out[(size_t)row * p.ld_hd + unit] = hv * dm;
            }
        }
        return;
    }
    for (int idx = tid; idx < MROWS * SK_COLS; idx += SK_THREADS) {
        const int row = idx >> 4, col = idx & 15;
        const int n = n0 + col;
        if (row >= p.B || n >= p.N) continue;
        float v = outv[row][col];
        if (p.add) v += p.add[(size_t)row * p.ld_add + n];
        if (p.epi == EPI_TANH) v = tanhf(v);
        else if (p.epi == EPI_TANHBWD) { const float t = p.aux[(size_t)row * p.ld_aux + n]; v *= (1.f - t * t); }
        p.Y[(size_t)row * p.ldy + n] = v;
    }
}

int skinny(cudaStream_t st, const SkinnyArgs& p, bool exact) {
    AST_CHECK(p.B >= 1 && p.B <= 32, "skinny: batch %d unsupported (1..32)", p.B);
    AST_CHECK(p.K[0] % 16 == 0 && p.K[1] % 16 == 0, "skinny: K (%d,%d) must be multiples of 16", p.K[0], p.K[1]);
    AST_CHECK(p.ldx[0] % 4 == 0 && p.ldw[0] % 4 == 0 && (p.K[1] == 0 || (p.ldx[1] % 4 == 0 && p.ldw[1] % 4 == 0)),
              "skinny: leading dims must be multiples of 4");
    if (p.epi == EPI_LSTM) AST_CHECK(p.N % 16 == 0, "skinny: LSTM epilogue needs N %% 16 == 0");
    const int grid = cdiv(p.N, SK_COLS);
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
// s[b][t] = enc[eb][t][:] . v[b][:]      (eb = b, or 0 when one utterance is shared by all hyps)
__global__ void attn_dot_kernel(const float* __restrict__ enc, long long enc_bs, const float* __restrict__ v, int ldv,
                                float* __restrict__ s, int Tp, int H) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= Tp) return;
    const float* e = enc + (size_t)b * enc_bs + (size_t)t * H;
    const float* vv = v + (size_t)b * ldv;
    float acc = 0.f;
    for (int j = lane * 4; j < H; j += 128) {
        const float4 a = *reinterpret_cast<const float4*>(e + j);
        const float4 c = *reinterpret_cast<const float4*>(vv + j);
        acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s[(size_t)b * Tp + t] = acc;
}
int attn_dot(cudaStream_t st, const float* enc, long long enc_bs, const float* v, int ldv, float* s, int B, int Tp, int H) {
    AST_CHECK(H % 4 == 0 && ldv % 4 == 0, "attn_dot: H/ldv must be multiples of 4");
    dim3 grid(cdiv(Tp, 8), B);
    attn_dot_kernel<<<grid, 256, 0, st>>>(enc, enc_bs, v, ldv, s, Tp, H);
    AST_LAUNCH_OK();
    return 0;
}

// alpha = softmax_t(s) ; cv[b][j] = sum_t alpha[t] * enc[eb][t][j]
__global__ void attn_ctx_kernel(const float* __restrict__ enc, long long enc_bs, const float* __restrict__ s,
                                float* __restrict__ alpha, float* __restrict__ cv, int ld_cv, int Tp, int H) {
    extern __shared__ float sa[];          // Tp alphas
    __shared__ float scratch[32];
    const int b = blockIdx.y;
    const float* sb = s + (size_t)b * Tp;
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) mx = fmaxf(mx, sb[t]);
    mx = block_max(mx, scratch);
    float sum = 0.f;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) { const float e = expf(sb[t] - mx); sa[t] = e; sum += e; }
    sum = block_sum(sum, scratch);
    const float inv = 1.f / sum;
    __syncthreads();
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) {
        const float al = sa[t] * inv;
        sa[t] = al;
        if (blockIdx.x == 0) alpha[(size_t)b * Tp + t] = al;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H) return;
    const float* e = enc + (size_t)b * enc_bs + j;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int t = 0;
    for (; t + 3 < Tp; t += 4) {
        a0 = fmaf(sa[t], e[(size_t)t * H], a0);
        a1 = fmaf(sa[t + 1], e[(size_t)(t + 1) * H], a1);
        a2 = fmaf(sa[t + 2], e[(size_t)(t + 2) * H], a2);
        a3 = fmaf(sa[t + 3], e[(size_t)(t + 3) * H], a3);
    }
    for (; t < Tp; ++t) a0 = fmaf(sa[t], e[(size_t)t * H], a0);
    cv[(size_t)b * ld_cv + j] = (a0 + a1) + (a2 + a3);
}
int attn_ctx(cudaStream_t st, const float* enc, long long enc_bs, const float* s, float* alpha, float* cv, int ld_cv,
             int B, int Tp, int H) {
    dim3 grid(cdiv(H, 128), B);
    attn_ctx_kernel<<<grid, 128, sizeof(float) * Tp, st>>>(enc, enc_bs, s, alpha, cv, ld_cv, Tp, H);
    AST_LAUNCH_OK();
    return 0;
}

// backward: ds = alpha*(dalpha - sum alpha*dalpha); dq[j] = sum_t ds[t]*enc[t][j];
//           d_enc[t][j] += alpha[t]*dcv[j] + ds[t]*q[j]
__global__ void attn_bwd_kernel(const float* __restrict__ enc, float* __restrict__ d_enc, long long enc_bs,
                                const float* __restrict__ alpha, const float* __restrict__ dalpha,
                                const float* __restrict__ dcv, int ld_dcv, const float* __restrict__ qv, int ld_q,
                                float* __restrict__ dq, int ld_dq, int Tp, int H) {
    extern __shared__ float sm[];          // alpha[Tp], ds[Tp]
    __shared__ float scratch[32];
    float* sal = sm; float* sds = sm + Tp;
    const int b = blockIdx.y;
    float dot = 0.f;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) {
        const float al = alpha[(size_t)b * Tp + t], da = dalpha[(size_t)b * Tp + t];
        sal[t] = al; sds[t] = da; dot = fmaf(al, da, dot);
    }
    dot = block_sum(dot, scratch);
    __syncthreads();
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) sds[t] = sal[t] * (sds[t] - dot);
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H) return;
    const float dcvj = dcv[(size_t)b * ld_dcv + j], qj = qv[(size_t)b * ld_q + j];
    const float* e = enc + (size_t)b * enc_bs + j;
    float* de = d_enc + (size_t)b * enc_bs + j;
    float a0 = 0.f, a1 = 0.f;
    int t = 0;
    for (; t + 1 < Tp; t += 2) {
        const float e0 = e[(size_t)t * H], e1 = e[(size_t)(t + 1) * H];
        const float d0 = de[(size_t)t * H], d1 = de[(size_t)(t + 1) * H];
        a0 = fmaf(sds[t], e0, a0); a1 = fmaf(sds[t + 1], e1, a1);
        de[(size_t)t * H] = d0 + sal[t] * dcvj + sds[t] * qj;
        de[(size_t)(t + 1) * H] = d1 + sal[t + 1] * dcvj + sds[t + 1] * qj;
    }
    for (; t < Tp; ++t) {
        a0 = fmaf(sds[t], e[(size_t)t * H], a0);
        de[(size_t)t * H] += sal[t] * dcvj + sds[t] * qj;
    }
    dq[(size_t)b * ld_dq + j] = a0 + a1;
}
int attn_bwd(cudaStream_t st, const float* enc, float* d_enc, long long enc_bs, const float* alpha, const float* dalpha,
             const float* dcv, int ld_dcv, const float* qv, int ld_q, float* dq, int ld_dq, int B, int Tp, int H) {
    dim3 grid(cdiv(H, 128), B);
    attn_bwd_kernel<<<grid, 128, sizeof(float) * 2 * Tp, st>>>(enc, d_enc, enc_bs, alpha, dalpha, dcv, ld_dcv, qv, ld_q,
                                                              dq, ld_dq, Tp, H);
    AST_LAUNCH_OK();
    return 0;
}

// =========================================================================================
// fused softmax cross-entropy forward + backward + argmax (seq2seq.py:448,468; Appendix A.7)
// =========================================================================================
// One CTA per batch row.  row_loss[b] = -w[t]*logp[t]/B ; z <- (softmax(z) - onehot(t)) * w[t]/B in place
// (cols [V,ldz) zeroed) ; argmax[b] = lowest index of the row maximum.  target < 0 -> argmax only.
__global__ void softmax_ce_kernel(float* __restrict__ z, int ldz, const int* __restrict__ y, int ldy_tok, int step_next,
                                  float* __restrict__ row_loss, int* __restrict__ argmax_out, int B, int V,
                                  int write_grad) {
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    const int b = blockIdx.x;
    float* zr = z + (size_t)b * ldz;
    float mx = -INFINITY; int mi = 0x7fffffff;
    for (int n = threadIdx.x; n < V; n += blockDim.x) {
        const float v = zr[n];
        if (v > mx) { mx = v; mi = n; }
    }
    // (max, lowest index) reduction
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    if (lane == 0) { scratch[w] = mx; iscratch[w] = mi; }
    __syncthreads();
    mx = (lane < nw) ? scratch[lane] : -INFINITY;
    mi = (lane < nw) ? iscratch[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    if (threadIdx.x == 0 && argmax_out) argmax_out[b] = mi;
    if (!write_grad) return;
    float sum = 0.f;
    for (int n = threadIdx.x; n < V; n += blockDim.x) sum += expf(zr[n] - mx);
    sum = block_sum(sum, scratch);
    const float lse = mx + logf(sum);
    const int t = y[(size_t)b * ldy_tok + step_next];
    const float wt = (t == 0) ? 0.f : 1.f;                      // mask_pad_id: class weight 0 for PAD
    const float scale = wt / (float)B;
    if (threadIdx.x == 0) row_loss[b] = -scale * (zr[t] - lse);
    __syncthreads();
    for (int n = threadIdx.x; n < ldz; n += blockDim.x) {
        float gz = 0.f;
        if (n < V) gz = (expf(zr[n] - lse) - (n == t ? 1.f : 0.f)) * scale;
        zr[n] = gz;
    }
}
int softmax_ce(cudaStream_t st, float* z, int ldz, const int* y, int ldy_tok, int step_next, float* row_loss,
               int* argmax_out, int B, int V, bool write_grad) {
    softmax_ce_kernel<<<B, 256, 0, st>>>(z, ldz, y, ldy_tok, step_next, row_loss, argmax_out, B, V, write_grad ? 1 : 0);
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

}  // namespace ast

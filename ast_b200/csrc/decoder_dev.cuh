// Device-side building blocks of the decoder step (seq2seq.py:336-396, 448, 468), shared by the per-step
// kernels (decoder.cu: decode_step protocol, beam search) and the persistent decoder-sequence kernels
// (dec_seq.cu: a whole forward_loss decoder pass / its backward in ONE cooperative launch).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ast {

constexpr int SK_THREADS = 256;
constexpr int SK_COLS = 16;

struct SkinnySmem {
    float red[8][32][SK_COLS + 1];
    float outv[32][SK_COLS + 1];
};

// One 16-column group of  Y[b][n] = epi( sum_seg sum_k X_seg[b][k] * W_seg[n][k] + bias[n] )  by one CTA of 8 warps:
// K split over the warps in chunks of 16 (tensor cores: mma.sync m16n8k8 TF32, batch rows = M), cross-warp
// reduction through smem, fused epilogue.  Must be called by all 256 threads of the CTA.
template <int MT, bool EXACT>
__device__ __forceinline__ void skinny_tile(const SkinnyArgs& p, int n0, SkinnySmem& sm) {
    constexpr int MROWS = 16 * MT;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;

    float acc[2][MT][4];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][mt][j] = 0.f;

    const int nch0 = p.K[0] >> 4, nch1 = p.K[1] >> 4;
    const int nn0 = n0 + g, nn1 = n0 + 8 + g;
    // A warp's chunks are c = w, w + 8, ...; the loads of UB chunks are issued before the first mma of the batch (a
    // load -> mma loop serialises on L2 latency: a phase of the persistent kernels is a handful of such round trips).
    constexpr int UB = MT == 1 ? 3 : 2;
    for (int c0 = w; c0 < nch0 + nch1; c0 += 8 * UB) {
        float4 wv0[UB], wv1[UB], xv[UB][MT][2];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int c = c0 + 8 * u;
            wv0[u] = wv1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) xv[u][mt][0] = xv[u][mt][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nch0 + nch1) {
                const int seg = c < nch0 ? 0 : 1;
                const int k = ((seg ? c - nch0 : c) << 4) + 4 * q;
                const float* __restrict__ W = p.W[seg];
                const float* __restrict__ X = p.X[seg];
                const int ldw = p.ldw[seg], ldx = p.ldx[seg];
                if (nn0 < p.N) wv0[u] = __ldg(reinterpret_cast<const float4*>(W + (size_t)nn0 * ldw + k));
                if (nn1 < p.N) wv1[u] = __ldg(reinterpret_cast<const float4*>(W + (size_t)nn1 * ldw + k));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int row = mt * 16 + g + 8 * hf;
                        // activations may have been written earlier in the same (persistent) kernel: no read-only path
                        if (row < p.B) xv[u][mt][hf] = __ldcg(reinterpret_cast<const float4*>(X + (size_t)row * ldx + k));
                    }
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            // two k-steps; lane q supplies physical k = 4q+{0,1} then 4q+{2,3} for both operands (zero chunks add nothing)
            const float b0a[2] = {wv0[u].x, wv0[u].y}, b0b[2] = {wv0[u].z, wv0[u].w};
            const float b1a[2] = {wv1[u].x, wv1[u].y}, b1b[2] = {wv1[u].z, wv1[u].w};
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const float aa[4] = {xv[u][mt][0].x, xv[u][mt][1].x, xv[u][mt][0].y, xv[u][mt][1].y};
                const float ab[4] = {xv[u][mt][0].z, xv[u][mt][1].z, xv[u][mt][0].w, xv[u][mt][1].w};
                mma_f32<EXACT>(acc[0][mt], aa, b0a);
                mma_f32<EXACT>(acc[0][mt], ab, b0b);
                mma_f32<EXACT>(acc[1][mt], aa, b1a);
                mma_f32<EXACT>(acc[1][mt], ab, b1b);
            }
        }
    }
    __syncthreads();          // sm may still be read by the previous tile's epilogue
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            sm.red[w][mt * 16 + g][s * 8 + 2 * q] = acc[s][mt][0];
            sm.red[w][mt * 16 + g][s * 8 + 2 * q + 1] = acc[s][mt][1];
            sm.red[w][mt * 16 + g + 8][s * 8 + 2 * q] = acc[s][mt][2];
            sm.red[w][mt * 16 + g + 8][s * 8 + 2 * q + 1] = acc[s][mt][3];
        }
    __syncthreads();
    for (int idx = tid; idx < MROWS * SK_COLS; idx += SK_THREADS) {
        const int row = idx >> 4, col = idx & 15;
        float v = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) v += sm.red[ww][row][col];
        const int n = n0 + col;
        if (p.bias && n < p.N) v += __ldg(p.bias + n);
        sm.outv[row][col] = v;
    }
    __syncthreads();

    if (p.epi == EPI_LSTM) {
        // 16 columns = 4 hidden units x (a,i,f,o)
        for (int idx = tid; idx < MROWS * 4; idx += SK_THREADS) {
            const int row = idx >> 2, ul = idx & 3;
            const int unit = (n0 >> 2) + ul;
            if (row < p.B && 4 * unit < p.N) {
                const int Hh = p.N >> 2;
                const float ga = tanhf(sm.outv[row][4 * ul]), gi = sigmoidf_(sm.outv[row][4 * ul + 1]);
                const float gf = sigmoidf_(sm.outv[row][4 * ul + 2]), go = sigmoidf_(sm.outv[row][4 * ul + 3]);
                const float c = ga * gi + gf * __ldcg(p.c_prev + (size_t)row * Hh + unit);
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
        float v = sm.outv[row][col];
        if (p.add) v += __ldcg(p.add + (size_t)row * p.ld_add + n);
        if (p.epi == EPI_TANH) v = tanhf(v);
        else if (p.epi == EPI_TANHBWD) { const float t = __ldcg(p.aux + (size_t)row * p.ld_aux + n); v *= (1.f - t * t); }
        else if (p.epi == EPI_CELLBWD && n < p.cb_ncols) {
            // v = d(out) of hidden unit n of the layer below: fused LSTM cell backward -> dG in place over act
            const int Hh = p.cb_H;
            const size_t e = (size_t)row * Hh + n;
            const float dm = dropout_scale(p.seed, p.drop_stream, (uint32_t)(p.drop_base + e), p.drop);
            float dh = v * dm;
            if (p.cb_dh_rec) dh += __ldcg(p.cb_dh_rec + (size_t)row * p.cb_ld_dh_rec + n);
            const float4 a = __ldcg(reinterpret_cast<const float4*>(p.cb_act + e * 4));
            const float cc = __ldcg(p.cb_c + e), cp = __ldcg(p.cb_c_prev + e);
            const float tc = tanhf(cc);
            const float dct = __ldcg(p.cb_dc + e) + dh * a.w * (1.f - tc * tc);
            float4 dg;
            dg.x = dct * a.y * (1.f - a.x * a.x);
            dg.y = dct * a.x * a.y * (1.f - a.y);
            dg.z = dct * cp * a.z * (1.f - a.z);
            dg.w = dh * tc * a.w * (1.f - a.w);
            p.cb_dc[e] = dct * a.z;
            *reinterpret_cast<float4*>(p.cb_act + e * 4) = dg;
        }
        if (p.sc_demb && n < p.sc_E) {      // EmbedID backward: scatter-add (duplicates accumulate)
            const float dm = dropout_scale(p.seed, p.drop_stream, (uint32_t)(p.drop_base + (size_t)row * p.sc_E + n), p.drop);
            atomicAdd(p.sc_demb + (size_t)__ldg(p.sc_words + row) * p.sc_E + n, v * dm);
        }
        if (p.Y) p.Y[(size_t)row * p.ldy + n] = v;
        if (p.Y2) p.Y2[(size_t)row * p.ldy2 + n] = v;
    }
}

// ---- attention pieces, one CTA each ------------------------------------------------------------------------
// scores for 8 consecutive t (one per warp): s[b][t] = enc[b][t][:] . v[b][:]
__device__ __forceinline__ void attn_dot_block(const float* __restrict__ enc, long long enc_bs, const float* v, int ldv,
                                               float* s, int Tp, int H, int b, int tblock) {
    const int t = tblock * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= Tp) return;
    const float* e = enc + (size_t)b * enc_bs + (size_t)t * H;
    const float* vv = v + (size_t)b * ldv;
    float acc = 0.f;
    for (int j = lane * 4; j < H; j += 128) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(e + j));
        const float4 c = __ldcg(reinterpret_cast<const float4*>(vv + j));
        acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s[(size_t)b * Tp + t] = acc;
}

// alpha = softmax_t(s[b]) ; cv[b][j0..j0+128) = sum_t alpha[t] * enc[b][t][j].  256 threads: 8 warps split T',
// each lane owns 4 consecutive columns (float4) of the 128-column slice; partials reduced through smem.
// sa: Tp floats, part: 8*128 floats, scratch: 32 floats.
__device__ __forceinline__ void attn_ctx_block(const float* __restrict__ enc, long long enc_bs, const float* s, float* alpha,
                                               float* cv, int ld_cv, int Tp, int H, int b, int jblock, float* sa, float* part,
                                               float* scratch) {
    const float* sb = s + (size_t)b * Tp;
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) mx = fmaxf(mx, __ldcg(sb + t));
    mx = block_max(mx, scratch);
    float sum = 0.f;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) { const float e = expf(__ldcg(sb + t) - mx); sa[t] = e; sum += e; }
    sum = block_sum(sum, scratch);
    const float inv = 1.f / sum;
    __syncthreads();
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) {
        const float al = sa[t] * inv;
        sa[t] = al;
        if (jblock == 0) alpha[(size_t)b * Tp + t] = al;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = jblock * 128 + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < H) {
        const float* e = enc + (size_t)b * enc_bs + j;
        for (int t = w; t < Tp; t += 8) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(e + (size_t)t * H));
            const float a = sa[t];
            acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
        }
    }
    *reinterpret_cast<float4*>(part + w * 128 + lane * 4) = acc;
    __syncthreads();
    if (threadIdx.x < 128) {
        float r = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) r += part[ww * 128 + threadIdx.x];
        const int jj = jblock * 128 + threadIdx.x;
        if (jj < H) cv[(size_t)b * ld_cv + jj] = r;
    }
    __syncthreads();
}

// backward: ds = alpha*(dalpha - sum alpha*dalpha); dq[j] = sum_t ds[t]*enc[t][j]; d_enc[t][j] += alpha[t]*dcv[j] + ds[t]*q[j]
// sm: 2*Tp floats (alpha, ds), part: 8*128 floats.
__device__ __forceinline__ void attn_bwd_block(const float* __restrict__ enc, float* d_enc, long long enc_bs, const float* alpha,
                                               const float* dalpha, const float* dcv, int ld_dcv, const float* qv, int ld_q,
                                               float* dq, int ld_dq, int Tp, int H, int b, int jblock, float* sm, float* part,
                                               float* scratch) {
    float* sal = sm; float* sds = sm + Tp;
    float dot = 0.f;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) {
        const float al = __ldcg(alpha + (size_t)b * Tp + t), da = __ldcg(dalpha + (size_t)b * Tp + t);
        sal[t] = al; sds[t] = da; dot = fmaf(al, da, dot);
    }
    dot = block_sum(dot, scratch);
    __syncthreads();
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) sds[t] = sal[t] * (sds[t] - dot);
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = jblock * 128 + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < H) {
        const float4 dcvj = __ldcg(reinterpret_cast<const float4*>(dcv + (size_t)b * ld_dcv + j));
        const float4 qj = __ldcg(reinterpret_cast<const float4*>(qv + (size_t)b * ld_q + j));
        const float* e = enc + (size_t)b * enc_bs + j;
        float* de = d_enc + (size_t)b * enc_bs + j;
        for (int t = w; t < Tp; t += 8) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(e + (size_t)t * H));
            float4 d = __ldcg(reinterpret_cast<const float4*>(de + (size_t)t * H));
            const float a = sal[t], dsv = sds[t];
            acc.x = fmaf(dsv, x.x, acc.x); acc.y = fmaf(dsv, x.y, acc.y); acc.z = fmaf(dsv, x.z, acc.z); acc.w = fmaf(dsv, x.w, acc.w);
            d.x += a * dcvj.x + dsv * qj.x; d.y += a * dcvj.y + dsv * qj.y; d.z += a * dcvj.z + dsv * qj.z; d.w += a * dcvj.w + dsv * qj.w;
            *reinterpret_cast<float4*>(de + (size_t)t * H) = d;
        }
    }
    *reinterpret_cast<float4*>(part + w * 128 + lane * 4) = acc;
    __syncthreads();
    if (threadIdx.x < 128) {
        float r = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) r += part[ww * 128 + threadIdx.x];
        const int jj = jblock * 128 + threadIdx.x;
        if (jj < H) dq[(size_t)b * ld_dq + jj] = r;
    }
    __syncthreads();
}

// fused softmax-CE forward + backward + argmax for one row (one CTA, 256 threads).  See softmax_ce_kernel.
__device__ __forceinline__ int softmax_ce_row(float* zr, int ldz, int V, int target, int B, float* row_loss, int write_grad,
                                              float* scratch, int* iscratch) {
    float mx = -INFINITY; int mi = 0x7fffffff;
    for (int n = threadIdx.x; n < V; n += blockDim.x) {
        const float v = __ldcg(zr + n);
        if (v > mx) { mx = v; mi = n; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    __syncthreads();
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
    if (!write_grad) return mi;
    float sum = 0.f;
    for (int n = threadIdx.x; n < V; n += blockDim.x) sum += expf(__ldcg(zr + n) - mx);
    sum = block_sum(sum, scratch);
    const float lse = mx + logf(sum);
    const float wt = (target == 0) ? 0.f : 1.f;                      // mask_pad_id: class weight 0 for PAD
    const float scale = wt / (float)B;
    if (threadIdx.x == 0) *row_loss = -scale * (__ldcg(zr + target) - lse);
    __syncthreads();
    for (int n = threadIdx.x; n < ldz; n += blockDim.x) {
        float gz = 0.f;
        if (n < V) gz = (expf(__ldcg(zr + n) - lse) - (n == target ? 1.f : 0.f)) * scale;
        zr[n] = gz;
    }
    return mi;
}

}  // namespace ast

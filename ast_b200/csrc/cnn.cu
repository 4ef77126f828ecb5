// CNN front-end pieces (seq2seq.py:158-180): im2col for CNN_0 (K = kh*kw = 117, 1 % of FLOPs),
// train-mode BatchNorm statistics / apply(+ReLU, + relayout) / backward, the col2im gather for
// CNN_1's data gradient and the weight (un)permutes.  The convolutions themselves are GEMMs
// (gemm_simt.cu / gemm_tc.cu); CNN_1 reads its input as an overlapping-rows matrix so no im2col
// buffer is ever materialised for the 20 %-of-FLOPs layer.
//
// HBM layouts (channels-last, time-major rows):
//   raw0 / a0 : [b][f][t1][C0]          a0p : [b][f][S0][C0]  (S0 = 2*ceil((T1+8)/2), 4 zero rows
//                                              in front, zeros behind: CNN_1's padding in memory)
//   raw1      : [b][f][Rs][C1]  with Rs = S0/2 "virtual rows" per (b,f) segment, rows >= T' junk
//   rnn_in    : [t'][b][c*F'+f] (the reference's feature order, seq2seq.py:177-179)
#include "common.cuh"
#include "kernels.h"

namespace ast {

// ---- im2col for CNN_0 -------------------------------------------------------------------
// cols[(b,f,t1)][kt*kw+kd] = X[b][sh*t1 - ph + kt][sw*f + kd]  (zero outside [0,T)), ld = ldc.
__global__ void im2col0_kernel(const float* __restrict__ X, float* __restrict__ cols, int B, int T,
                               int D, int Fp, int T1, int kh, int kw, int sh, int sw, int ph, int ldc) {
    const int K = kh * kw;
    const size_t total = (size_t)B * Fp * T1 * ldc;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % ldc);
        const size_t row = i / ldc;
        const int t1 = (int)(row % T1);
        const int f = (int)((row / T1) % Fp);
        const int b = (int)(row / ((size_t)T1 * Fp));
        float v = 0.f;
        if (k < K) {
            const int kt = k / kw, kd = k % kw;
            const int t = sh * t1 - ph + kt, d = sw * f + kd;
            if (t >= 0 && t < T && d < D) v = X[((size_t)b * T + t) * D + d];
        }
        cols[i] = v;
    }
}

int im2col0(cudaStream_t st, const float* X, float* cols, int B, int T, int D, int Fp, int T1, int kh,
            int kw, int sh, int sw, int ph, int ldc) {
    const size_t total = (size_t)B * Fp * T1 * ldc;
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    im2col0_kernel<<<grid, 256, 0, st>>>(X, cols, B, T, D, Fp, T1, kh, kw, sh, sw, ph, ldc);
    AST_LAUNCH_OK();
    return 0;
}

// ---- BN statistics: per-channel sum / sum-of-squares over valid rows (double accumulation) ---
// x: rows x C (ld = C).  Row r is valid iff (r % seg_rows) < seg_valid.  stats[0..C) += sum,
// stats[C..2C) += sumsq.  blockDim = (32, 8): x = channel lane, y = row lane.
__global__ void bn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int rows, int C,
                                int seg_rows, int seg_valid, int rows_per_block) {
    __shared__ double ssum[8][33], ssq[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    double s = 0.0, q = 0.0;
    if (c < C) {
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            if ((r % seg_rows) < seg_valid) {
                const float v = x[(size_t)r * C + c];
                s += v; q += (double)v * v;
            }
        }
    }
    ssum[threadIdx.y][threadIdx.x] = s;
    ssq[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { s += ssum[j][threadIdx.x]; q += ssq[j][threadIdx.x]; }
        atomicAdd(&stats[c], s);
        atomicAdd(&stats[C + c], q);
    }
}

// mean / invstd from the sums, plus Chainer's running-stat update (Appendix A.2):
// avg_mean = .9*avg_mean + .1*mean ; avg_var = .9*avg_var + .1*var*m/(m-1).
__global__ void bn_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ avg_mean,
                                   float* __restrict__ avg_var, int C, double m, float eps, float decay,
                                   int update_running) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = stats[c] / m;
    double var = stats[C + c] / m - mu * mu;
    if (var < 0) var = 0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (update_running) {
        const double adj = m / fmax(m - 1.0, 1.0);
        avg_mean[c] = decay * avg_mean[c] + (1.f - decay) * (float)mu;
        avg_var[c] = decay * avg_var[c] + (1.f - decay) * (float)(var * adj);
    }
}

// eval-mode: mean = avg_mean, invstd = 1/sqrt(avg_var + eps)
__global__ void bn_eval_prepare_kernel(const float* __restrict__ avg_mean, const float* __restrict__ avg_var,
                                       float* __restrict__ mean, float* __restrict__ invstd, int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = avg_mean[c];
    invstd[c] = 1.f / sqrtf(avg_var[c] + eps);
}

int bn_stats(cudaStream_t st, const float* x, double* stats, int rows, int C, int seg_rows, int seg_valid) {
    AST_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st));
    const int rpb = std::max(64, cdiv(rows, 148 * 4 / std::max(1, cdiv(C, 32))));
    dim3 grid(cdiv(C, 32), cdiv(rows, rpb)), block(32, 8);
    bn_stats_kernel<<<grid, block, 0, st>>>(x, stats, rows, C, seg_rows, seg_valid, rpb);
    AST_LAUNCH_OK();
    return 0;
}

int bn_finalize(cudaStream_t st, const double* stats, float* mean, float* invstd, float* avg_mean,
                float* avg_var, int C, double m, float eps, float decay, bool update_running) {
    bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats, mean, invstd, avg_mean, avg_var, C, m, eps, decay,
                                                     update_running ? 1 : 0);
    AST_LAUNCH_OK();
    return 0;
}

int bn_eval_prepare(cudaStream_t st, const float* avg_mean, const float* avg_var, float* mean, float* invstd,
                    int C, float eps) {
    bn_eval_prepare_kernel<<<cdiv(C, 128), 128, 0, st>>>(avg_mean, avg_var, mean, invstd, C, eps);
    AST_LAUNCH_OK();
    return 0;
}

// ---- BN apply + ReLU for layer 0: raw0 [seg][T1][C] -> a0p [seg][S0][C] with `pad` zero rows in front
__global__ void bn_relu_pad_kernel(const float* __restrict__ raw, float* __restrict__ out,
                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   int nseg, int T1, int S0, int pad, int C) {
    const int C4 = C >> 2;
    const size_t total = (size_t)nseg * S0 * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const size_t row = i / C4;
        const int s = (int)(row % S0);
        const size_t seg = row / S0;
        const int t1 = s - pad;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t1 >= 0 && t1 < T1) {
            const float4 x = *reinterpret_cast<const float4*>(raw + ((seg * T1 + t1) * (size_t)C) + c4 * 4);
            const float4 mu = *reinterpret_cast<const float4*>(mean + c4 * 4);
            const float4 is = *reinterpret_cast<const float4*>(invstd + c4 * 4);
            const float4 g = *reinterpret_cast<const float4*>(gamma + c4 * 4);
            const float4 b = *reinterpret_cast<const float4*>(beta + c4 * 4);
            o.x = fmaxf(g.x * ((x.x - mu.x) * is.x) + b.x, 0.f);
            o.y = fmaxf(g.y * ((x.y - mu.y) * is.y) + b.y, 0.f);
            o.z = fmaxf(g.z * ((x.z - mu.z) * is.z) + b.z, 0.f);
            o.w = fmaxf(g.w * ((x.w - mu.w) * is.w) + b.w, 0.f);
        }
        *reinterpret_cast<float4*>(out + row * (size_t)C + c4 * 4) = o;
    }
}

int bn_relu_pad(cudaStream_t st, const float* raw, float* out, const float* mean, const float* invstd,
                const float* gamma, const float* beta, int nseg, int T1, int S0, int pad, int C) {
    AST_CHECK(C % 4 == 0, "bn_relu_pad: C %% 4 != 0");
    const size_t total = (size_t)nseg * S0 * (C / 4);
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    bn_relu_pad_kernel<<<grid, 256, 0, st>>>(raw, out, mean, invstd, gamma, beta, nseg, T1, S0, pad, C);
    AST_LAUNCH_OK();
    return 0;
}

// ---- BN apply + ReLU + relayout for the last CNN layer -------------------------------------
// raw1 [b][f][Rs][C] (rows t' < Tp valid) -> rnn_in[t'][b][c*Fp+f] and (optionally) the step-ordered
// copy for the reverse stack rnn_rev[i] = rnn_in[(-i) mod Tp]  (seq2seq.py:219 `X[-i]`).
// One CTA per (t', b): reads the F' channel rows of that frame (each C contiguous floats), writes the whole R = C*F' feature
// row of rnn_in / rnn_rev contiguously through a shared-memory transpose.  (The element-per-thread version wrote with a
// stride of F' floats between lanes: ncu showed 2.0 TB/s, 31 % of the HBM roofline, for a pure relayout.)
__global__ void __launch_bounds__(128) bn_relu_to_rnn_kernel(const float* __restrict__ raw, float* __restrict__ rnn_in,
                                      float* __restrict__ rnn_rev, const float* __restrict__ mean,
                                      const float* __restrict__ invstd, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int B, int Fp, int Rs, int Tp, int C) {
    extern __shared__ float row[];          // R floats
    const int t = blockIdx.x / B, b = blockIdx.x - t * B;
    const int R = C * Fp;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float g = gamma[c], mu = mean[c], is = invstd[c], bt = beta[c];
        for (int f = 0; f < Fp; ++f)        // same expression as the backward mask (g * xhat + beta > 0)
            row[c * Fp + f] = fmaxf(g * ((raw[(((size_t)b * Fp + f) * Rs + t) * C + c] - mu) * is) + bt, 0.f);
    }
    __syncthreads();
    float4* o1 = reinterpret_cast<float4*>(rnn_in + ((size_t)t * B + b) * R);
    float4* o2 = rnn_rev ? reinterpret_cast<float4*>(rnn_rev + ((size_t)((Tp - t) % Tp) * B + b) * R) : nullptr;   // step i with (-i) mod Tp == t
    for (int i = threadIdx.x; i < R / 4; i += blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(row)[i];
        o1[i] = v;
        if (o2) o2[i] = v;
    }
}

int bn_relu_to_rnn(cudaStream_t st, const float* raw, float* rnn_in, float* rnn_rev, const float* mean,
                   const float* invstd, const float* gamma, const float* beta, int B, int Fp, int Rs, int Tp,
                   int C) {
    AST_CHECK((C * Fp) % 4 == 0, "bn_relu_to_rnn: C*F' must be a multiple of 4");
    bn_relu_to_rnn_kernel<<<Tp * B, 128, sizeof(float) * C * Fp, st>>>(raw, rnn_in, rnn_rev, mean, invstd, gamma, beta, B, Fp, Rs, Tp, C);
    AST_LAUNCH_OK();
    return 0;
}

// ---- BN + ReLU backward --------------------------------------------------------------------
// Generic over an index functor giving, for (valid row r, channel c), where dy lives.
// Pass 1: dbeta[c] = sum dy*mask, dgamma[c] = sum dy*mask*xhat   (double atomics into stats[2C])
// Pass 2: dx = gamma*invstd*(dy*mask - dbeta/m - xhat*dgamma/m)  (junk rows get 0)
struct DyFromRnn {          // last CNN layer: dy = d_rnn_in[t'][b][c*Fp+f] (+ d_rnn_rev[(Tp-t')%Tp][b][..])
    const float* d_in; const float* d_rev; int B, Fp, Rs, Tp, C;
    __device__ __forceinline__ float operator()(int r, int c) const {
        const int t = r % Rs; const int sf = r / Rs; const int f = sf % Fp; const int b = sf / Fp;
        const int R = C * Fp;
        float v = d_in[((size_t)t * B + b) * R + c * Fp + f];
        if (d_rev) v += d_rev[((size_t)((Tp - t) % Tp) * B + b) * R + c * Fp + f];
        return v;
    }
};
struct DyFromPadded {       // layer 0: dy = da0p[seg][pad + t1][c]
    const float* d; int T1, S0, pad, C;
    __device__ __forceinline__ float operator()(int r, int c) const {
        const int t1 = r % T1; const int seg = r / T1;
        return d[((size_t)seg * S0 + pad + t1) * C + c];
    }
};

template <class DY>
__global__ void bn_bwd_reduce_kernel(DY dyf, const float* __restrict__ raw, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, double* __restrict__ stats, int rows, int C,
                                     int seg_rows, int seg_valid, int rows_per_block) {
    __shared__ double sb[8][33], sg[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    double db = 0.0, dg = 0.0;
    if (c < C) {
        const float mu = mean[c], is = invstd[c], g = gamma[c], bt = beta[c];
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            if ((r % seg_rows) < seg_valid) {
                const float xh = (raw[(size_t)r * C + c] - mu) * is;
                if (g * xh + bt > 0.f) {
                    const float dy = dyf(r, c);
                    db += dy; dg += (double)dy * xh;
                }
            }
        }
    }
    sb[threadIdx.y][threadIdx.x] = db;
    sg[threadIdx.y][threadIdx.x] = dg;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { db += sb[j][threadIdx.x]; dg += sg[j][threadIdx.x]; }
        atomicAdd(&stats[c], db);
        atomicAdd(&stats[C + c], dg);
    }
}

template <class DY>
__global__ void bn_bwd_apply_kernel(DY dyf, const float* __restrict__ raw, float* __restrict__ dx,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const double* __restrict__ stats, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int rows, int C, int seg_rows, int seg_valid,
                                    double m) {
    const size_t total = (size_t)rows * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int r = (int)(i / C);
        float o = 0.f;
        const float db = (float)stats[c], dg = (float)stats[C + c];
        if ((r % seg_rows) < seg_valid) {
            const float is = invstd[c], g = gamma[c];
            const float xh = (raw[i] - mean[c]) * is;
            const float dy = (g * xh + beta[c] > 0.f) ? dyf(r, c) : 0.f;
            o = g * is * (dy - db / (float)m - xh * (dg / (float)m));
        }
        dx[i] = o;
        if (r == 0) { dgamma[c] = dg; dbeta[c] = db; }
    }
}

template <class DY>
static int bn_bwd_impl(cudaStream_t st, DY dyf, const float* raw, float* dx, const float* mean,
                       const float* invstd, const float* gamma, const float* beta, double* stats,
                       float* dgamma, float* dbeta, int rows, int C, int seg_rows, int seg_valid, double m) {
    AST_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st));
    const int rpb = std::max(64, cdiv(rows, 148 * 4 / std::max(1, cdiv(C, 32))));
    dim3 grid(cdiv(C, 32), cdiv(rows, rpb)), block(32, 8);
    bn_bwd_reduce_kernel<DY><<<grid, block, 0, st>>>(dyf, raw, mean, invstd, gamma, beta, stats, rows, C,
                                                      seg_rows, seg_valid, rpb);
    AST_LAUNCH_OK();
    const size_t total = (size_t)rows * C;
    const int g2 = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    bn_bwd_apply_kernel<DY><<<g2, 256, 0, st>>>(dyf, raw, dx, mean, invstd, gamma, beta, stats, dgamma, dbeta,
                                                 rows, C, seg_rows, seg_valid, m);
    AST_LAUNCH_OK();
    return 0;
}

// ---- last CNN layer: dy arrives in the RNN feature layout  d_rnn_in[t'][b][c*F'+f] (+ d_rnn_rev[(T'-t')%T'][b][..]) ----
// Same two passes as bn_bwd_impl, organised per (t', b) feature row so that every global access is contiguous: the row(s)
// of dy are staged in shared memory (float4 loads), then lane c walks f.
__global__ void __launch_bounds__(256) bn_bwd_rnn_reduce_kernel(const float* __restrict__ d_in, const float* __restrict__ d_rev,
                                         const float* __restrict__ raw, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, double* __restrict__ stats, int B, int Fp, int Rs,
                                         int Tp, int C, int rows_per_block, double* __restrict__ partials) {
    extern __shared__ float dyrow[];        // R floats
    const int R = C * Fp, TB = Tp * B;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(TB, r0 + rows_per_block);
    constexpr int MAXC = 4;                 // channels per thread (C <= 1024)
    double db[MAXC], dg[MAXC];
    float mu[MAXC], is[MAXC], g[MAXC], bt[MAXC];
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
        const int c = threadIdx.x + k * 256;
        db[k] = dg[k] = 0.0;
        if (c < C) { mu[k] = mean[c]; is[k] = invstd[c]; g[k] = gamma[c]; bt[k] = beta[c]; }
    }
    for (int r = r0; r < r1; ++r) {
        const int t = r / B, b = r - t * B;
        const float4* s1 = reinterpret_cast<const float4*>(d_in + (size_t)r * R);
        const float4* s2 = d_rev ? reinterpret_cast<const float4*>(d_rev + ((size_t)((Tp - t) % Tp) * B + b) * R) : nullptr;
        __syncthreads();
        for (int i = threadIdx.x; i < R / 4; i += 256) {
            float4 v = s1[i];
            if (s2) { const float4 w = s2[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
            reinterpret_cast<float4*>(dyrow)[i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = threadIdx.x + k * 256;
            if (c < C)
                for (int f = 0; f < Fp; ++f) {
                    const float xh = (raw[(((size_t)b * Fp + f) * Rs + t) * C + c] - mu[k]) * is[k];
                    if (g[k] * xh + bt[k] > 0.f) {
                        const float dy = dyrow[c * Fp + f];
                        db[k] += dy; dg[k] += (double)dy * xh;
                    }
                }
        }
    }
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
        const int c = threadIdx.x + k * 256;
        if (c < C) {
            // 569 blocks x 2C double atomics onto 2C addresses serialised in L2 (most of this kernel's 83 us): per-block partial sums
            // (coalesced stores) and one small summation kernel instead
            if (partials) { partials[(size_t)blockIdx.x * 2 * C + c] = db[k]; partials[(size_t)blockIdx.x * 2 * C + C + c] = dg[k]; }
            else { atomicAdd(&stats[c], db[k]); atomicAdd(&stats[C + c], dg[k]); }
        }
    }
}
// stats[i] = sum over blocks of partials[b][i]: one CTA per 32 columns, 32 x 32 threads (thread (y, x) sums blocks y, y+32, ... of
// column x: coalesced 256-byte rows), then a shared-memory tree over y.  (A single thread per column walking all ~570 blocks was a
// 81 us serial chain of dependent L2 loads.)
__global__ void __launch_bounds__(1024) bn_partials_sum_kernel(const double* __restrict__ partials, int nblocks, int n, double* __restrict__ stats) {
    __shared__ double red[32][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + x;
    double s = 0.0;
    if (i < n)
        for (int b = y; b < nblocks; b += 32) s += partials[(size_t)b * n + i];
    red[y][x] = s;
    __syncthreads();
    for (int h = 16; h > 0; h >>= 1) {
        if (y < h) red[y][x] += red[y + h][x];
        __syncthreads();
    }
    if (y == 0 && i < n) stats[i] = red[0][x];
}

// one CTA per (b, t in [0, Rs)): rows t >= T' are the junk rows of the padded segment and get zeros
__global__ void __launch_bounds__(128) bn_bwd_rnn_apply_kernel(const float* __restrict__ d_in, const float* __restrict__ d_rev,
                                        const float* __restrict__ raw, float* __restrict__ dx, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const double* __restrict__ stats,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int Fp, int Rs, int Tp,
                                        int C, float inv_m) {
    extern __shared__ float dyrow[];
    const int b = blockIdx.x / Rs, t = blockIdx.x - b * Rs;
    const int R = C * Fp;
    if (t >= Tp) {
        for (int f = 0; f < Fp; ++f)
            for (int c = threadIdx.x; c < C; c += blockDim.x) dx[(((size_t)b * Fp + f) * Rs + t) * C + c] = 0.f;
        return;
    }
    const float4* s1 = reinterpret_cast<const float4*>(d_in + ((size_t)t * B + b) * R);
    const float4* s2 = d_rev ? reinterpret_cast<const float4*>(d_rev + ((size_t)((Tp - t) % Tp) * B + b) * R) : nullptr;
    for (int i = threadIdx.x; i < R / 4; i += blockDim.x) {
        float4 v = s1[i];
        if (s2) { const float4 w = s2[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        reinterpret_cast<float4*>(dyrow)[i] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float db = (float)stats[c], dg = (float)stats[C + c];
        const float is = invstd[c], g = gamma[c], mu = mean[c], bt = beta[c];
        for (int f = 0; f < Fp; ++f) {
            const size_t i = (((size_t)b * Fp + f) * Rs + t) * C + c;
            const float xh = (raw[i] - mu) * is;
            const float dy = (g * xh + bt > 0.f) ? dyrow[c * Fp + f] : 0.f;
            dx[i] = g * is * (dy - db * inv_m - xh * (dg * inv_m));
        }
        if (blockIdx.x == 0) { dgamma[c] = dg; dbeta[c] = db; }
    }
}

int bn_bwd_from_rnn(cudaStream_t st, const float* d_in, const float* d_rev, const float* raw, float* dx,
                    const float* mean, const float* invstd, const float* gamma, const float* beta,
                    double* stats, float* dgamma, float* dbeta, int B, int Fp, int Rs, int Tp, int C, double* partials,
                    int partial_blocks) {
    AST_CHECK((C * Fp) % 4 == 0 && C <= 1024, "bn_bwd_from_rnn: need C*F' %% 4 == 0 and C <= 1024");
    const int TB = Tp * B;
    const int rpb = std::max(1, cdiv(TB, 148 * 4));
    const int nblk = cdiv(TB, rpb);
    if (partials && nblk > partial_blocks) partials = nullptr;
    if (!partials) AST_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st));
    const size_t smem = sizeof(float) * C * Fp;
    bn_bwd_rnn_reduce_kernel<<<nblk, 256, smem, st>>>(d_in, d_rev, raw, mean, invstd, gamma, beta, stats, B, Fp, Rs, Tp, C, rpb, partials);
    AST_LAUNCH_OK();
    if (partials) {
        bn_partials_sum_kernel<<<cdiv(2 * C, 32), 1024, 0, st>>>(partials, nblk, 2 * C, stats);
        AST_LAUNCH_OK();
    }
    bn_bwd_rnn_apply_kernel<<<B * Rs, 128, smem, st>>>(d_in, d_rev, raw, dx, mean, invstd, gamma, beta, stats, dgamma, dbeta, B, Fp,
                                                       Rs, Tp, C, (float)(1.0 / ((double)B * Fp * Tp)));
    AST_LAUNCH_OK();
    return 0;
}

int bn_bwd_from_padded(cudaStream_t st, const float* da0p, const float* raw, float* dx, const float* mean,
                       const float* invstd, const float* gamma, const float* beta, double* stats,
                       float* dgamma, float* dbeta, int nseg, int T1, int S0, int pad, int C) {
    DyFromPadded f{da0p, T1, S0, pad, C};
    return bn_bwd_impl(st, f, raw, dx, mean, invstd, gamma, beta, stats, dgamma, dbeta, nseg * T1, C, T1, T1,
                       (double)nseg * T1);
}

// ---- col2im gather for CNN_1's data gradient ------------------------------------------------
// dA: virtual rows [seg][Rs][kh*C0] (row r covers padded input rows sh*r .. sh*r+kh-1).
// da0p[seg][s][ci] = sum_{kt : (s-kt) % sh == 0, r=(s-kt)/sh in [0,Tp)} dA[seg][r][kt*C0+ci]
__global__ void col2im1_kernel(const float* __restrict__ dA, float* __restrict__ da0p, int nseg, int S0,
                               int Rs, int Tp, int C0, int kh, int sh) {
    const int C4 = C0 >> 2;
    const size_t total = (size_t)nseg * S0 * C4;
    const int ldA = kh * C0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const size_t row = i / C4;
        const int s = (int)(row % S0);
        const size_t seg = row / S0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kt = s % sh; kt < kh; kt += sh) {
            const int r = (s - kt) / sh;
            if (s - kt >= 0 && r < Tp) {
                const float4 v = *reinterpret_cast<const float4*>(dA + (seg * Rs + r) * (size_t)ldA + kt * C0 + c4 * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        *reinterpret_cast<float4*>(da0p + row * (size_t)C0 + c4 * 4) = acc;
    }
}

int col2im1(cudaStream_t st, const float* dA, float* da0p, int nseg, int S0, int Rs, int Tp, int C0, int kh, int sh) {
    const size_t total = (size_t)nseg * S0 * (C0 / 4);
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    col2im1_kernel<<<grid, 256, 0, st>>>(dA, da0p, nseg, S0, Rs, Tp, C0, kh, sh);
    AST_LAUNCH_OK();
    return 0;
}

// ---- weight permutes: W1 (co, ci, kt) <-> W1p (co, kt, ci) ------------------------------------
__global__ void permute_w1_kernel(const float* __restrict__ src, float* __restrict__ dst, int Co, int Ci, int Kt,
                                  int to_p) {
    const int total = Co * Ci * Kt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        // i indexes the (co, kt, ci) layout
        const int ci = i % Ci; const int kt = (i / Ci) % Kt; const int co = i / (Ci * Kt);
        const int j = (co * Ci + ci) * Kt + kt;       // (co, ci, kt)
        if (to_p) dst[i] = src[j]; else dst[j] = src[i];
    }
}
// Weights of CNN_1's data gradient as a TRANSPOSED convolution over time (stride 2): the padded-input rows of parity p receive
// the taps kt = p, p+2, ... from the output rows r = j - u, u = (kt - p) / 2.  Reading d(raw1) as an overlapping-rows matrix
// (row j = rows j-U .. j of d(raw1), U = ntaps_p - 1) makes the whole gradient two plain GEMMs straight into da0p - no
// (M1 x 9*C0) im2col-gradient buffer, no col2im pass.  Wt_p[(u' * Co + co) * Ci + ci] = W1p[co][kt = 2 (U - u') + p][ci].
__global__ void build_w1t_kernel(const float* __restrict__ W1p, float* __restrict__ Wt, int Co, int Ci, int Kt, int p, int ntaps) {
    const int total = ntaps * Co * Ci;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Ci; const int co = (i / Ci) % Co; const int up = i / (Ci * Co);
        const int kt = 2 * (ntaps - 1 - up) + p;
        Wt[i] = W1p[((size_t)co * Kt + kt) * Ci + ci];
    }
}
int build_w1t(cudaStream_t st, const float* W1p, float* Wt, int Co, int Ci, int Kt, int p) {
    const int ntaps = (Kt - p + 1) / 2;
    if (ntaps <= 0) return 0;
    build_w1t_kernel<<<std::min(cdiv(ntaps * Co * Ci, 256), 148 * 8), 256, 0, st>>>(W1p, Wt, Co, Ci, Kt, p, ntaps);
    AST_LAUNCH_OK();
    return 0;
}

int permute_w1(cudaStream_t st, const float* src, float* dst, int Co, int Ci, int Kt, bool to_p) {
    permute_w1_kernel<<<std::min(cdiv(Co * Ci * Kt, 256), 148 * 8), 256, 0, st>>>(src, dst, Co, Ci, Kt, to_p ? 1 : 0);
    AST_LAUNCH_OK();
    return 0;
}

}  // namespace ast

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
__global__ void bn_partials_sum_kernel(const double* __restrict__ partials, int nblocks, int n, double* __restrict__ stats);

// float4 channel vectors: lane x of a (32, 8) block owns channels 4 (32 bx + x) .. + 3 (a warp reads 512 contiguous bytes per row),
// row lane y strides the block's rows.  fp32 accumulation per thread (a few dozen rows), the 8 row lanes are added in double, one
// partial sum per block (bn_partials_sum adds them: deterministic, no double atomics; the scalar version ran at 1.2 TB/s).
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ partials, int rows, int C,
                                int seg_rows, int seg_valid, int rows_per_block) {
    __shared__ double ssum[8][32][4], ssq[8][32][4];
    const int c = 4 * (blockIdx.x * 32 + threadIdx.x);
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
#pragma unroll 4
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            if ((r % seg_rows) < seg_valid) {
                const float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * C + c);
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
                q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { ssum[threadIdx.y][threadIdx.x][k] = s[k]; ssq[threadIdx.y][threadIdx.x][k] = q[k]; }
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double* out = partials + (size_t)blockIdx.y * 2 * C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { a += ssum[j][threadIdx.x][k]; b += ssq[j][threadIdx.x][k]; }
            out[c + k] = a; out[C + c + k] = b;
        }
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

int bn_stats(cudaStream_t st, const float* x, double* stats, int rows, int C, int seg_rows, int seg_valid, double* partials,
             int partial_blocks) {
    AST_CHECK(C % 4 == 0 && partials != nullptr && partial_blocks >= 1, "bn_stats: C %% 4 != 0 or no buffer for the partial sums");
    const int gx = cdiv(C, 128);
    int rpb = std::max(32, cdiv(rows, std::max(1, 148 * 4 / gx)));
    rpb = std::max(rpb, cdiv(rows, partial_blocks));
    const int gy = cdiv(rows, rpb);
    dim3 grid(gx, gy), block(32, 8);
    bn_stats_kernel<<<grid, block, 0, st>>>(x, partials, rows, C, seg_rows, seg_valid, rpb);
    AST_LAUNCH_OK();
    bn_partials_sum_kernel<<<cdiv(2 * C, 32), 1024, 0, st>>>(partials, gy, 2 * C, stats);
    AST_LAUNCH_OK();
    return 0;
}

// bn_partials_sum + bn_finalize in one launch (forward, train mode): block = 16 channels, thread (y, x): x < 16 sums the partial
// SUMS of channel c0 + x, x >= 16 the partial SUMS OF SQUARES of channel c0 + x - 16 over blocks y, y + 32, ...; tree over y; the
// first 16 threads then finish mean / invstd / running statistics exactly as bn_finalize_kernel does.
__global__ void __launch_bounds__(1024) bn_sum_finalize_kernel(const double* __restrict__ partials, int nblocks, int C, double* __restrict__ stats,
                                                               float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ avg_mean,
                                                               float* __restrict__ avg_var, double m, float eps, float decay, int update_running) {
    __shared__ double red[32][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int c = blockIdx.x * 16 + (x & 15);
    const int col = (x < 16 ? 0 : C) + c;
    double s = 0.0;
    if (c < C)
        for (int b = y; b < nblocks; b += 32) s += partials[(size_t)b * 2 * C + col];
    red[y][x] = s;
    __syncthreads();
    for (int hh = 16; hh > 0; hh >>= 1) {
        if (y < hh) red[y][x] += red[y + hh][x];
        __syncthreads();
    }
    if (y == 0 && c < C) stats[col] = red[0][x];
    if (y == 0 && x < 16 && c < C) {
        const double mu = red[0][x] / m;
        double var = red[0][x + 16] / m - mu * mu;
        if (var < 0) var = 0;
        mean[c] = (float)mu;
        invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (update_running) {
            const double adj = m / fmax(m - 1.0, 1.0);
            avg_mean[c] = decay * avg_mean[c] + (1.f - decay) * (float)mu;
            avg_var[c] = decay * avg_var[c] + (1.f - decay) * (float)(var * adj);
        }
    }
}

int bn_stats_finalize(cudaStream_t st, const float* x, double* stats, int rows, int C, int seg_rows, int seg_valid, double* partials,
                      int partial_blocks, float* mean, float* invstd, float* avg_mean, float* avg_var, double m, float eps, float decay,
                      bool update_running) {
    AST_CHECK(C % 4 == 0 && partials != nullptr && partial_blocks >= 1, "bn_stats_finalize: C %% 4 != 0 or no buffer for the partial sums");
    const int gx = cdiv(C, 128);
    int rpb = std::max(32, cdiv(rows, std::max(1, 148 * 4 / gx)));
    rpb = std::max(rpb, cdiv(rows, partial_blocks));
    const int gy = cdiv(rows, rpb);
    dim3 grid(gx, gy), block(32, 8);
    bn_stats_kernel<<<grid, block, 0, st>>>(x, partials, rows, C, seg_rows, seg_valid, rpb);
    AST_LAUNCH_OK();
    bn_sum_finalize_kernel<<<cdiv(C, 16), 1024, 0, st>>>(partials, gy, C, stats, mean, invstd, avg_mean, avg_var, m, eps, decay, update_running ? 1 : 0);
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
    __device__ __forceinline__ float4 vec4(int r, int c) const {        // c % 4 == 0
        const int t1 = r % T1; const int seg = r / T1;
        return *reinterpret_cast<const float4*>(d + ((size_t)seg * S0 + pad + t1) * C + c);
    }
};

// Both passes walk rows of C contiguous floats with float4 channel vectors: lane x of a (32, 8) block owns channels
// 4 (32 bx + x) .. + 3, row lane y strides the block's rows, so a warp reads 512 contiguous bytes per row.  The first pass
// accumulates in fp32 per thread (a few dozen terms), reduces the 8 row lanes in double and writes per-block partial sums
// (bn_partials_sum adds them up; double atomics onto 2C addresses serialised in L2: 47 us for 32 MB at C = 128).
template <class DY>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(DY dyf, const float* __restrict__ raw, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, double* __restrict__ partials, int rows, int C,
                                     int seg_rows, int seg_valid, int rows_per_block) {
    __shared__ double sb[8][32][4], sg[8][32][4];
    const int c = 4 * (blockIdx.x * 32 + threadIdx.x);
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    float db[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
#pragma unroll 4
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            if ((r % seg_rows) < seg_valid) {
                const float4 x = *reinterpret_cast<const float4*>(raw + (size_t)r * C + c);
                const float4 dy = dyf.vec4(r, c);
                const float xh[4] = {(x.x - mu.x) * is.x, (x.y - mu.y) * is.y, (x.z - mu.z) * is.z, (x.w - mu.w) * is.w};
                const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w}, dd[4] = {dy.x, dy.y, dy.z, dy.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (gg[k] * xh[k] + bb[k] > 0.f) { db[k] += dd[k]; dg[k] += dd[k] * xh[k]; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { sb[threadIdx.y][threadIdx.x][k] = db[k]; sg[threadIdx.y][threadIdx.x][k] = dg[k]; }
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double* out = partials + (size_t)blockIdx.y * 2 * C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double b = 0.0, gq = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { b += sb[j][threadIdx.x][k]; gq += sg[j][threadIdx.x][k]; }
            out[c + k] = b; out[C + c + k] = gq;
        }
    }
}

template <class DY>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(DY dyf, const float* __restrict__ raw, float* __restrict__ dx,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const double* __restrict__ stats, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int rows, int C, int seg_rows, int seg_valid,
                                    double m) {
    const int C4 = C >> 2;
    const size_t total = (size_t)rows * C4;
    const float inv_m = (float)(1.0 / m);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = 4 * (int)(i % C4);
        const int r = (int)(i / C4);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        const float db[4] = {(float)stats[c], (float)stats[c + 1], (float)stats[c + 2], (float)stats[c + 3]};
        const float dg[4] = {(float)stats[C + c], (float)stats[C + c + 1], (float)stats[C + c + 2], (float)stats[C + c + 3]};
        if ((r % seg_rows) < seg_valid) {
            const float4 x = *reinterpret_cast<const float4*>(raw + (size_t)r * C + c);
            const float4 dy = dyf.vec4(r, c);
            const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
            const float4 g = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
            const float xv[4] = {x.x, x.y, x.z, x.w}, dd[4] = {dy.x, dy.y, dy.z, dy.w}, mm[4] = {mu.x, mu.y, mu.z, mu.w};
            const float ii[4] = {is.x, is.y, is.z, is.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w};
            float ov[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xh = (xv[k] - mm[k]) * ii[k];
                const float d = (gg[k] * xh + bb[k] > 0.f) ? dd[k] : 0.f;
                ov[k] = gg[k] * ii[k] * (d - db[k] * inv_m - xh * (dg[k] * inv_m));
            }
            o = make_float4(ov[0], ov[1], ov[2], ov[3]);
        }
        *reinterpret_cast<float4*>(dx + (size_t)r * C + c) = o;
        if (r == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { dgamma[c + k] = dg[k]; dbeta[c + k] = db[k]; }
        }
    }
}

template <class DY>
static int bn_bwd_impl(cudaStream_t st, DY dyf, const float* raw, float* dx, const float* mean,
                       const float* invstd, const float* gamma, const float* beta, double* stats,
                       float* dgamma, float* dbeta, int rows, int C, int seg_rows, int seg_valid, double m,
                       double* partials, int partial_blocks) {
    AST_CHECK(C % 4 == 0, "bn_bwd: C %% 4 != 0");
    const int gx = cdiv(C, 128);
    int rpb = std::max(64, cdiv(rows, std::max(1, 148 * 4 / gx)));
    rpb = std::max(rpb, cdiv(rows, partial_blocks));
    const int gy = cdiv(rows, rpb);
    dim3 grid(gx, gy), block(32, 8);
    bn_bwd_reduce_kernel<DY><<<grid, block, 0, st>>>(dyf, raw, mean, invstd, gamma, beta, partials, rows, C,
                                                      seg_rows, seg_valid, rpb);
    AST_LAUNCH_OK();
    bn_partials_sum_kernel<<<cdiv(2 * C, 32), 1024, 0, st>>>(partials, gy, 2 * C, stats);
    AST_LAUNCH_OK();
    const size_t total = (size_t)rows * (C / 4);
    const int g2 = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    bn_bwd_apply_kernel<DY><<<g2, 256, 0, st>>>(dyf, raw, dx, mean, invstd, gamma, beta, stats, dgamma, dbeta,
                                                 rows, C, seg_rows, seg_valid, m);
    AST_LAUNCH_OK();
    return 0;
}

// ---- last CNN layer: dy arrives in the RNN feature layout  d_rnn_in[t'][b][c*F'+f] (+ d_rnn_rev[(T'-t')%T'][b][..]) ----
// Same two passes as bn_bwd_impl, organised per (t', b) feature row so that every global access is contiguous: the row(s)
// of dy are staged in shared memory (float4 loads), then lane c walks f.
// Thread x of a (R/4, 2) block owns four consecutive elements j = c F' + f of the R-float feature row: dy (and the reverse
// stack's) is one 128-bit load per row, perfectly coalesced; the four raw values are a gather inside three 128-byte lines per
// warp (channel c = j / F', plane f = j % F').  No shared-memory staging and no barrier inside the row loop (the staged
// version ran at 1.1 TB/s, 17 % of the HBM roofline: two barriers per row, one row in flight per block).  fp32 accumulation
// per thread (a few dozen terms); the F' planes and the 2 row lanes of a channel are added in double; per-block partial sums,
// bn_partials_sum adds them up (deterministic).
template <int FP>
__global__ void __launch_bounds__(768) bn_bwd_rnn_reduce_kernel(const float* __restrict__ d_in, const float* __restrict__ d_rev,
                                         const float* __restrict__ raw, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, int B, int Fp_rt, int Rs,
                                         int Tp, int C, int rows_per_block, double* __restrict__ partials) {
    extern __shared__ float sm_bn[];        // params [4][C] | sums [2 (db, dg)][2 (row lane)][R]
    const int Fp = FP > 0 ? FP : Fp_rt;
    const int R = C * Fp, TB = Tp * B;
    float* prm = sm_bn; float* sums = sm_bn + 4 * C;
    const int x = threadIdx.x, y = threadIdx.y, nx = blockDim.x;
    for (int c = y * nx + x; c < C; c += 2 * nx) { prm[c] = mean[c]; prm[C + c] = invstd[c]; prm[2 * C + c] = gamma[c]; prm[3 * C + c] = beta[c]; }
    __syncthreads();
    const int r0 = blockIdx.x * rows_per_block, r1 = min(TB, r0 + rows_per_block);
    constexpr int MAXK = 2;                 // R <= 4 * blockDim.x * MAXK
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        const int j0 = 4 * (x + k * nx);
        if (j0 >= R) break;
        int off[4]; float mu[4], is[4], g[4], bt[4], db[4], dg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = j0 + i, c = j / Fp, f = j - c * Fp;
            off[i] = f * Rs * C + c;
            mu[i] = prm[c]; is[i] = prm[C + c]; g[i] = prm[2 * C + c]; bt[i] = prm[3 * C + c];
            db[i] = dg[i] = 0.f;
        }
#pragma unroll 2
        for (int r = r0 + y; r < r1; r += 2) {
            const int t = r / B, b = r - t * B;
            float4 dy = *reinterpret_cast<const float4*>(d_in + (size_t)r * R + j0);
            if (d_rev) {
                const float4 w = *reinterpret_cast<const float4*>(d_rev + ((size_t)((Tp - t) % Tp) * B + b) * R + j0);
                dy.x += w.x; dy.y += w.y; dy.z += w.z; dy.w += w.w;
            }
            const float* rw = raw + ((size_t)b * Fp * Rs + t) * C;
            const float dd[4] = {dy.x, dy.y, dy.z, dy.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xh = (rw[off[i]] - mu[i]) * is[i];
                if (g[i] * xh + bt[i] > 0.f) { db[i] += dd[i]; dg[i] += dd[i] * xh; }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { sums[y * R + j0 + i] = db[i]; sums[(2 + y) * R + j0 + i] = dg[i]; }
    }
    __syncthreads();
    double* out = partials + (size_t)blockIdx.x * 2 * C;
    for (int c = y * nx + x; c < C; c += 2 * nx) {
        double b2 = 0.0, g2 = 0.0;
        for (int f = 0; f < Fp; ++f) {
            b2 += (double)sums[c * Fp + f] + (double)sums[R + c * Fp + f];
            g2 += (double)sums[2 * R + c * Fp + f] + (double)sums[3 * R + c * Fp + f];
        }
        out[c] = b2; out[C + c] = g2;
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
    AST_CHECK(partials != nullptr && partial_blocks >= 1, "bn_bwd_from_rnn: no buffer for the per-block partial sums");
    const int R = C * Fp;
    const int nx = std::min(384, R / 4);
    AST_CHECK(R <= 4 * nx * 2, "bn_bwd_from_rnn: feature row of %d floats too long", R);
    int rpb = 2 * std::max(1, cdiv(TB, 2 * 148 * 2));
    rpb = std::max(rpb, cdiv(TB, partial_blocks));
    const int nblk = cdiv(TB, rpb);
    const size_t smem_r = sizeof(float) * (4 * (size_t)C + 4 * (size_t)R);
    if (Fp == 3) bn_bwd_rnn_reduce_kernel<3><<<nblk, dim3(nx, 2), smem_r, st>>>(d_in, d_rev, raw, mean, invstd, gamma, beta, B, Fp, Rs, Tp, C, rpb, partials);
    else if (Fp == 1) bn_bwd_rnn_reduce_kernel<1><<<nblk, dim3(nx, 2), smem_r, st>>>(d_in, d_rev, raw, mean, invstd, gamma, beta, B, Fp, Rs, Tp, C, rpb, partials);
    else bn_bwd_rnn_reduce_kernel<0><<<nblk, dim3(nx, 2), smem_r, st>>>(d_in, d_rev, raw, mean, invstd, gamma, beta, B, Fp, Rs, Tp, C, rpb, partials);
    AST_LAUNCH_OK();
    bn_partials_sum_kernel<<<cdiv(2 * C, 32), 1024, 0, st>>>(partials, nblk, 2 * C, stats);
    AST_LAUNCH_OK();
    bn_bwd_rnn_apply_kernel<<<B * Rs, 128, sizeof(float) * C * Fp, st>>>(d_in, d_rev, raw, dx, mean, invstd, gamma, beta, stats, dgamma, dbeta, B, Fp,
                                                                       Rs, Tp, C, (float)(1.0 / ((double)B * Fp * Tp)));
    AST_LAUNCH_OK();
    return 0;
}

int bn_bwd_from_padded(cudaStream_t st, const float* da0p, const float* raw, float* dx, const float* mean,
                       const float* invstd, const float* gamma, const float* beta, double* stats,
                       float* dgamma, float* dbeta, int nseg, int T1, int S0, int pad, int C, double* partials,
                       int partial_blocks) {
    DyFromPadded f{da0p, T1, S0, pad, C};
    return bn_bwd_impl(st, f, raw, dx, mean, invstd, gamma, beta, stats, dgamma, dbeta, nseg * T1, C, T1, T1,
                       (double)nseg * T1, partials, partial_blocks);
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

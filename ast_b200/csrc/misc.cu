// Memory-bound helpers: fused multi-tensor optimizer (nn.py:81-119 + Appendix A.10), utterance
// pack + CMVN (+frame-drop, +noise) (dataloader.py:83-108,156; seq2seq.py:297-305; Kaldi apply-cmvn),
// transposes, column sums, small copies.  All coalesced / float4 where the layout allows; grids are
// sized in multiples of the 148 SMs.
#include "common.cuh"
#include "kernels.h"

namespace ast {

// ---- optimizer -----------------------------------------------------------------------------
// pass 1: sum over the flat buffer of (gscale*g + wd*p)^2  -> norm_sq[0] (double)
__global__ void opt_sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ p, size_t n, float gscale,
                                  float wd, double* __restrict__ norm_sq) {
    __shared__ double sh[8];
    double s = 0.0;
    const size_t n4 = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        const float4 pv = reinterpret_cast<const float4*>(p)[i];
        const float a = gscale * gv.x + wd * pv.x, b = gscale * gv.y + wd * pv.y;
        const float c = gscale * gv.z + wd * pv.z, d = gscale * gv.w + wd * pv.w;
        s += (double)a * a + (double)b * b + (double)c * c + (double)d * d;
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        atomicAdd(norm_sq, t);
    }
}

// pass 2: WeightDecay -> GradientClipping(global norm) -> AMSGrad (eps outside the sqrt)
__global__ void opt_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                   float* __restrict__ v, float* __restrict__ vhat, size_t n, float gscale, float wd,
                                   float clip, const double* __restrict__ norm_sq, float alpha_t, float beta1,
                                   float beta2, float eps, FrozenRanges fr, float noise_sigma, unsigned long long noise_seed) {
    const double nrm = sqrt(norm_sq[0]);
    float rate = 1.f;
    if (clip > 0.f && nrm > 0.0) { const double r = (double)clip / nrm; if (r < 1.0) rate = (float)r; }
    const size_t n4 = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        bool frozen = false;
        for (int k = 0; k < fr.n; ++k) frozen |= (i * 4 >= fr.begin[k] && i * 4 < fr.end[k]);
        if (frozen) continue;
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i], hv = reinterpret_cast<float4*>(vhat)[i];
#define AST_ADAM1(c, k)                                                    \
        {                                                                 \
            float gg = (gscale * gv.c + wd * pv.c) * rate;                \
            if (noise_sigma > 0.f) gg += noise_sigma * rng_normal(noise_seed, 48, (uint32_t)(4 * i + k));      \
            mv.c += (1.f - beta1) * (gg - mv.c);                          \
            vv.c += (1.f - beta2) * (gg * gg - vv.c);                     \
            hv.c = fmaxf(hv.c, vv.c);                                     \
            pv.c -= alpha_t * mv.c / (sqrtf(hv.c) + eps);                 \
        }
        AST_ADAM1(x, 0) AST_ADAM1(y, 1) AST_ADAM1(z, 2) AST_ADAM1(w, 3)
#undef AST_ADAM1
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
        reinterpret_cast<float4*>(vhat)[i] = hv;
    }
}

int opt_sqnorm(cudaStream_t st, const float* g, const float* p, size_t n, float gscale, float wd, double* norm_sq) {
    AST_CHECK(n % 4 == 0, "opt_sqnorm: flat size must be a multiple of 4");
    AST_CUDA_OK(cudaMemsetAsync(norm_sq, 0, sizeof(double), st));
    opt_sqnorm_kernel<<<148 * 4, 256, 0, st>>>(g, p, n, gscale, wd, norm_sq);
    AST_LAUNCH_OK();
    return 0;
}
int opt_amsgrad(cudaStream_t st, float* p, const float* g, float* m, float* v, float* vhat, size_t n, float gscale,
                float wd, float clip, const double* norm_sq, float alpha_t, float beta1, float beta2, float eps,
                const FrozenRanges& fr, float noise_sigma, unsigned long long noise_seed) {
    AST_CHECK(n % 4 == 0, "opt_amsgrad: flat size must be a multiple of 4");
    opt_amsgrad_kernel<<<148 * 4, 256, 0, st>>>(p, g, m, v, vhat, n, gscale, wd, clip, norm_sq, alpha_t, beta1, beta2, eps, fr, noise_sigma, noise_seed);
    AST_LAUNCH_OK();
    return 0;
}

// optimizers.SGD (nn.py:91-93) behind the same hooks: WeightDecay -> GradientClipping(global norm) -> GradientNoise -> p -= lr * g
__global__ void opt_sgd_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float gscale, float wd, float clip,
                               const double* __restrict__ norm_sq, float lr, FrozenRanges fr, float noise_sigma, unsigned long long noise_seed) {
    const double nrm = sqrt(norm_sq[0]);
    float rate = 1.f;
    if (clip > 0.f && nrm > 0.0) { const double r = (double)clip / nrm; if (r < 1.0) rate = (float)r; }
    const size_t n4 = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        bool frozen = false;
        for (int k = 0; k < fr.n; ++k) frozen |= (i * 4 >= fr.begin[k] && i * 4 < fr.end[k]);
        if (frozen) continue;
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float gg[4] = {(gscale * gv.x + wd * pv.x) * rate, (gscale * gv.y + wd * pv.y) * rate, (gscale * gv.z + wd * pv.z) * rate, (gscale * gv.w + wd * pv.w) * rate};
        if (noise_sigma > 0.f)
#pragma unroll
            for (int c = 0; c < 4; ++c) gg[c] += noise_sigma * rng_normal(noise_seed, 48, (uint32_t)(4 * i + c));
        pv.x -= lr * gg[0]; pv.y -= lr * gg[1]; pv.z -= lr * gg[2]; pv.w -= lr * gg[3];
        reinterpret_cast<float4*>(p)[i] = pv;
    }
}
int opt_sgd(cudaStream_t st, float* p, const float* g, size_t n, float gscale, float wd, float clip, const double* norm_sq, float lr,
            const FrozenRanges& fr, float noise_sigma, unsigned long long noise_seed) {
    AST_CHECK(n % 4 == 0, "opt_sgd: flat size must be a multiple of 4");
    opt_sgd_kernel<<<148 * 4, 256, 0, st>>>(p, g, n, gscale, wd, clip, norm_sq, lr, fr, noise_sigma, noise_seed);
    AST_LAUNCH_OK();
    return 0;
}

// x *= a over a flat float buffer (data-parallel tail batches: a rank's gradient is weighted by its share of the global batch
// before the all-reduce; a == 0 writes exact zeros whatever x held)
__global__ void scale_kernel(float* __restrict__ x, float a, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = a == 0.f ? 0.f : x[i] * a;
}
int scale_inplace(cudaStream_t st, float* x, float a, size_t n) {
    if (n == 0) return 0;
    scale_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(x, a, n);
    AST_LAUNCH_OK();
    return 0;
}

// ---- pack + CMVN (+ frame drop + multiplicative noise) ---------------------------------------
// raw: concatenated utterances (sum_len x D); X[b][t][d] = t < len_b ? (raw*scale[b][d]+offset[b][d]) * keep * noise : 0
// Box-Muller normal from the counter RNG when noise_sigma > 0 and no explicit noise tensor is given.
__global__ void pack_cmvn_kernel(const float* __restrict__ raw, const long long* __restrict__ row_off,
                                 const int* __restrict__ lens, const float* __restrict__ scale,
                                 const float* __restrict__ offset, const unsigned char* __restrict__ keep,
                                 const float* __restrict__ noise, float noise_sigma, unsigned long long seed,
                                 float* __restrict__ X, int B, int T, int D) {
    const size_t total = (size_t)B * T * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const size_t bt = i / D;
        const int t = (int)(bt % T);
        const int b = (int)(bt / T);
        float v = 0.f;
        if (t < lens[b] && (keep == nullptr || keep[bt])) {
            v = raw[(size_t)(row_off[b] + t) * D + d];
            if (scale) v = v * scale[b * D + d] + offset[b * D + d];
            if (noise) v *= noise[i];
            else if (noise_sigma > 0.f) {
                const float u1 = fmaxf(rng_uniform(seed, 0x51u, (uint32_t)i), 1e-7f);
                const float u2 = rng_uniform(seed, 0x52u, (uint32_t)i);
                v *= 1.f + noise_sigma * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
            }
        }
        X[i] = v;
    }
}
// float4 variant (D % 4 == 0, 16-byte aligned buffers): one thread = 4 consecutive features of one frame, so the raw read,
// the scale/offset reads and the padded write are all 16-byte accesses; frame index arithmetic once per 4 elements.
__global__ void pack_cmvn4_kernel(const float* __restrict__ raw, const long long* __restrict__ row_off,
                                  const int* __restrict__ lens, const float* __restrict__ scale,
                                  const float* __restrict__ offset, const unsigned char* __restrict__ keep,
                                  const float* __restrict__ noise, float noise_sigma, unsigned long long seed,
                                  float* __restrict__ X, int B, int T, int D) {
    const int D4 = D >> 2;
    const size_t total4 = (size_t)B * T * D4;
    for (size_t i4 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i4 < total4; i4 += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(i4 % D4) * 4;
        const size_t bt = i4 / D4;
        const int t = (int)(bt % T);
        const int b = (int)(bt / T);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < lens[b] && (keep == nullptr || keep[bt])) {
            v = *reinterpret_cast<const float4*>(raw + (size_t)(row_off[b] + t) * D + d);
            if (scale) {
                const float4 sc = *reinterpret_cast<const float4*>(scale + (size_t)b * D + d);
                const float4 of = *reinterpret_cast<const float4*>(offset + (size_t)b * D + d);
                v.x = v.x * sc.x + of.x; v.y = v.y * sc.y + of.y; v.z = v.z * sc.z + of.z; v.w = v.w * sc.w + of.w;
            }
            const size_t i = bt * D + d;
            float f[4] = {1.f, 1.f, 1.f, 1.f};
            if (noise) { const float4 n = *reinterpret_cast<const float4*>(noise + i); f[0] = n.x; f[1] = n.y; f[2] = n.z; f[3] = n.w; }
            else if (noise_sigma > 0.f) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float u1 = fmaxf(rng_uniform(seed, 0x51u, (uint32_t)(i + j)), 1e-7f);
                    const float u2 = rng_uniform(seed, 0x52u, (uint32_t)(i + j));
                    f[j] = 1.f + noise_sigma * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
                }
            }
            v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
        }
        *reinterpret_cast<float4*>(X + bt * D + d) = v;
    }
}
int pack_cmvn(cudaStream_t st, const float* raw, const long long* row_off, const int* lens, const float* scale,
              const float* offset, const unsigned char* keep, const float* noise, float noise_sigma,
              unsigned long long seed, float* X, int B, int T, int D) {
    const size_t total = (size_t)B * T * D;
    if (total == 0) return 0;
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (D % 4 == 0 && al16(raw) && al16(X) && al16(scale) && al16(offset) && al16(noise)) {
        const int grid4 = (int)std::min<size_t>((total / 4 + 255) / 256, 148 * 16);
        pack_cmvn4_kernel<<<grid4, 256, 0, st>>>(raw, row_off, lens, scale, offset, keep, noise, noise_sigma, seed, X, B, T, D);
        AST_LAUNCH_OK();
        return 0;
    }
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    pack_cmvn_kernel<<<grid, 256, 0, st>>>(raw, row_off, lens, scale, offset, keep, noise, noise_sigma, seed, X, B, T, D);
    AST_LAUNCH_OK();
    return 0;
}

// X *= noise (explicit tensor) or X *= N(1, sigma)   (seq2seq.py:297-305 on an already packed batch)
__global__ void mul_noise_kernel(const float* __restrict__ X, float* __restrict__ Y, const float* __restrict__ noise,
                                 float sigma, unsigned long long seed, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float f;
        if (noise) f = noise[i];
        else {
            const float u1 = fmaxf(rng_uniform(seed, 0x51u, (uint32_t)i), 1e-7f);
            const float u2 = rng_uniform(seed, 0x52u, (uint32_t)i);
            f = 1.f + sigma * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
        }
        Y[i] = X[i] * f;
    }
}
int mul_noise(cudaStream_t st, const float* X, float* Y, const float* noise, float sigma, unsigned long long seed, size_t n) {
    if (n == 0) return 0;
    mul_noise_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(X, Y, noise, sigma, seed, n);
    AST_LAUNCH_OK();
    return 0;
}

// ---- transpose: dst (C x R, ld_dst) = src (R x C, ld_src)^T ------------------------------------
__global__ void transpose_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (c < C && r < R) dst[(size_t)c * ld_dst + r] = tile[threadIdx.x][j];
    }
}
int transpose(cudaStream_t st, const float* src, int ld_src, float* dst, int ld_dst, int R, int C) {
    dim3 grid(cdiv(C, 32), cdiv(R, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(src, ld_src, dst, ld_dst, R, C);
    AST_LAUNCH_OK();
    return 0;
}

// ---- column sums: out[c] (+)= sum_r x[r][c]  (bias gradients) ---------------------------------
__global__ void colsum_kernel(const float* __restrict__ x, int ld, float* __restrict__ out, int rows, int C, int rows_per_block) {
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    if (c < C) for (int r = r0 + threadIdx.y; r < r1; r += 8) s += x[(size_t)r * ld + c];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int j = 1; j < 8; ++j) s += sh[j][threadIdx.x];
        atomicAdd(&out[c], s);
    }
}
// float4 columns: lane x of a (32, 8) block owns columns 4 (32 bx + x) .. + 3, row lane y strides the block's rows with four
// independent 128-bit loads in flight (the scalar version above had one dependent load chain per thread: 1.1 TB/s on the
// 21 MB encoder gate-gradient matrices, and six of them sit on the tail of the step)
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ x, int ld, float* __restrict__ out, int rows, int C, int rows_per_block) {
    __shared__ float4 sh[8][32];
    const int c = 4 * (blockIdx.x * 32 + threadIdx.x);
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        const float* xp = x + c;
        int r = r0 + threadIdx.y;
        for (; r + 24 < r1; r += 32) {
            const float4 a = *reinterpret_cast<const float4*>(xp + (size_t)r * ld), b = *reinterpret_cast<const float4*>(xp + (size_t)(r + 8) * ld);
            const float4 d = *reinterpret_cast<const float4*>(xp + (size_t)(r + 16) * ld), e = *reinterpret_cast<const float4*>(xp + (size_t)(r + 24) * ld);
            s.x += (a.x + b.x) + (d.x + e.x); s.y += (a.y + b.y) + (d.y + e.y); s.z += (a.z + b.z) + (d.z + e.z); s.w += (a.w + b.w) + (d.w + e.w);
        }
        for (; r < r1; r += 8) {
            const float4 a = *reinterpret_cast<const float4*>(xp + (size_t)r * ld);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { const float4 t = sh[j][threadIdx.x]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        atomicAdd(&out[c], s.x);
        if (c + 1 < C) atomicAdd(&out[c + 1], s.y);
        if (c + 2 < C) atomicAdd(&out[c + 2], s.z);
        if (c + 3 < C) atomicAdd(&out[c + 3], s.w);
    }
}
int colsum(cudaStream_t st, const float* x, int ld, float* out, int rows, int C, bool accumulate) {
    if (!accumulate) AST_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
    if (rows <= 0) return 0;
    if (ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((C + 3) / 4) * 4 <= ld) {
        const int gx = cdiv(C, 128);
        const int rpb = std::max(32, cdiv(rows, std::max(1, 148 * 4 / gx)));
        dim3 grid(gx, cdiv(rows, rpb)), block(32, 8);
        colsum4_kernel<<<grid, block, 0, st>>>(x, ld, out, rows, C, rpb);
        AST_LAUNCH_OK();
        return 0;
    }
    const int rpb = std::max(32, cdiv(rows, std::max(1, 148 * 2 / std::max(1, cdiv(C, 32)))));
    dim3 grid(cdiv(C, 32), cdiv(rows, rpb)), block(32, 8);
    colsum_kernel<<<grid, block, 0, st>>>(x, ld, out, rows, C, rpb);
    AST_LAUNCH_OK();
    return 0;
}

// ---- strided 2-D copy: dst[r][c] = src[r][c] ---------------------------------------------------
__global__ void copy2d_kernel(const float* __restrict__ src, long long ld_src, float* __restrict__ dst, long long ld_dst,
                              int R, int C) {
    const size_t total = (size_t)R * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const size_t r = i / C;
        dst[r * ld_dst + c] = src[r * ld_src + c];
    }
}
int copy2d(cudaStream_t st, const float* src, long long ld_src, float* dst, long long ld_dst, int R, int C) {
    const size_t total = (size_t)R * C;
    if (total == 0) return 0;
    copy2d_kernel<<<(int)std::min<size_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(src, ld_src, dst, ld_dst, R, C);
    AST_LAUNCH_OK();
    return 0;
}

// Zero up to 16 float ranges in ONE launch (the encoder's 12 initial-state slots: 12 memsets were 40 us of launch gaps in front of
// the layer-0 projection).  Counts must be multiples of 4 floats, pointers 16-byte aligned.
__global__ void __launch_bounds__(256) zero_multi_kernel(ZeroBatch b) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* p = reinterpret_cast<float4*>(b.ptr[blockIdx.y]);
    const size_t n4 = b.count[blockIdx.y] / 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
}
int zero_multi(cudaStream_t st, const ZeroBatch& b) {
    if (b.n == 0) return 0;
    size_t mx = 0;
    for (int i = 0; i < b.n; ++i) mx = std::max(mx, b.count[i] / 4);
    if (mx == 0) return 0;
    dim3 grid((unsigned)std::min<size_t>((mx + 255) / 256, 32), b.n);
    zero_multi_kernel<<<grid, 256, 0, st>>>(b);
    AST_LAUNCH_OK();
    return 0;
}

// Several strided 2-D copies in ONE launch (decoder initial state = 12 slices of the encoder finals, seq2seq.py:318-334).
__global__ void copy2d_multi_kernel(Copy2DBatch b) {
    const Copy2DJob j = b.job[blockIdx.y];
    const size_t total = (size_t)j.R * j.C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % j.C); const size_t r = i / j.C;
        j.dst[r * j.ld_dst + c] = j.src[r * j.ld_src + c];
    }
}
int copy2d_multi(cudaStream_t st, const Copy2DBatch& b) {
    if (b.n == 0) return 0;
    size_t mx = 0;
    for (int i = 0; i < b.n; ++i) mx = std::max(mx, (size_t)b.job[i].R * b.job[i].C);
    if (mx == 0) return 0;
    dim3 grid((unsigned)std::min<size_t>((mx + 255) / 256, 64), b.n);
    copy2d_multi_kernel<<<grid, 256, 0, st>>>(b);
    AST_LAUNCH_OK();
    return 0;
}

// dst[i] += src[i]
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
int add_inplace(cudaStream_t st, float* dst, const float* src, size_t n) {
    if (n == 0) return 0;
    add_inplace_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(dst, src, n);
    AST_LAUNCH_OK();
    return 0;
}

// greedy decode bookkeeping (seq2seq.py:510-521): preds[step][b] = argmax[b]; seen_eos |= ...;
// done_step = first step after which every row has emitted EOS.
__global__ void greedy_track_kernel(const int* __restrict__ argmax, int* __restrict__ preds, int* __restrict__ seen,
                                    int* __restrict__ done_step, int B, int step, int eos) {
    __shared__ int all_seen;
    if (threadIdx.x == 0) all_seen = 1;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int wd = argmax[b];
        preds[(size_t)step * B + b] = wd;
        int s = seen[b] | (wd == eos ? 1 : 0);
        seen[b] = s;
        if (!s) all_seen = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0 && all_seen && done_step[0] < 0) done_step[0] = step;
}
int greedy_track(cudaStream_t st, const int* argmax, int* preds, int* seen, int* done_step, int B, int step, int eos) {
    greedy_track_kernel<<<1, 64, 0, st>>>(argmax, preds, seen, done_step, B, step, eos);
    AST_LAUNCH_OK();
    return 0;
}

}  // namespace ast

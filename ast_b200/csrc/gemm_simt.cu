// FP32 SIMT GEMM:  C[m,n] = alpha * sum_k opA(m,k) * opB(k,n) + beta * C[m,n] + bias[n]
//
// The exact-fp32 dense path (decode / "identical hypotheses" mode) and the bring-up / cross-check
// path for the tcgen05 TF32 kernel in gemm_tc.cu.  Row-major everywhere:
//   opA(m,k) = transA ? A[k*lda + m] : A[m*lda + k]
//   opB(k,n) = transB ? B[n*ldb + k] : B[k*ldb + n]
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered smem.
#include "common.cuh"
#include "kernels.h"

namespace ast {

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;

// Load a (rows x GBK) operand tile into smem laid out [k][row] (k-major) so the inner product
// loop reads conflict-free float4 along `row`.
//   KCONTIG = true : element (r,k) at base[r*ld + k]   (k contiguous in memory)
//   KCONTIG = false: element (r,k) at base[k*ld + r]   (r contiguous in memory)
template <bool KCONTIG, bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, int ld, int r0, int k0,
                                          int R, int K, float (*dst)[GBM + 4]) {
    const int tid = threadIdx.x;
    if (KCONTIG) {
        // 128 rows x 16 k = 512 float4; thread handles 2: row = idx/4, kq = idx%4
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * GTHREADS;
            const int r = idx >> 2, kq = (idx & 3) * 4;
            const int gr = r0 + r, gk = k0 + kq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gr < R) {
                const float* p = base + (size_t)gr * ld + gk;
                if (VEC && gk + 3 < K) {
                    const float4 t = *reinterpret_cast<const float4*>(p);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (gk + j < K) v[j] = p[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[kq + j][r] = v[j];
        }
    } else {
        // 16 k x 128 rows = 512 float4 along rows; k = idx/32, rq = idx%32
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * GTHREADS;
            const int k = idx >> 5, rq = (idx & 31) * 4;
            const int gk = k0 + k, gr = r0 + rq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gk < K) {
                const float* p = base + (size_t)gk * ld + gr;
                if (VEC && gr + 3 < R) {
                    const float4 t = *reinterpret_cast<const float4*>(p);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (gr + j < R) v[j] = p[j];
                }
            }
            *reinterpret_cast<float4*>(&dst[k][rq]) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(GTHREADS)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
             const float* __restrict__ B, int ldb, float beta, float* __restrict__ C, int ldc,
             const float* __restrict__ bias) {
    __shared__ __align__(16) float As[2][GBK][GBM + 4];
    __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 8x8 each

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (K + GBK - 1) / GBK;
    load_tile<!TA, VEC>(A, lda, m0, 0, M, K, As[0]);
    load_tile<TB, VEC>(B, ldb, n0, 0, N, K, Bs[0]);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile<!TA, VEC>(A, lda, m0, (kt + 1) * GBK, M, K, As[cur ^ 1]);
            load_tile<TB, VEC>(B, ldb, n0, (kt + 1) * GBK, N, K, Bs[cur ^ 1]);
        }
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            // rows ty*4..+3 and 64+ty*4..+3 ; cols tx*4..+3 and 64+tx*4..+3
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = alpha * acc[i][j];
            if (bias) v += bias[n];
            float* c = C + (size_t)m * ldc + n;
            if (beta != 0.f) v += beta * (*c);
            *c = v;
        }
    }
}

int sgemm_simt(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha, const float* A,
               int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias) {
    if (M <= 0 || N <= 0) return 0;
    AST_CHECK(K > 0, "sgemm_simt: K must be > 0 (got %d)", K);
    const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
    dim3 grid(cdiv(N, GBN), cdiv(M, GBM));
#define AST_SGEMM(TA, TB, V) \
    sgemm_kernel<TA, TB, V><<<grid, GTHREADS, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias)
    if (vec) {
        if (!ta && !tb) AST_SGEMM(false, false, true);
        else if (!ta && tb) AST_SGEMM(false, true, true);
        else if (ta && !tb) AST_SGEMM(true, false, true);
        else AST_SGEMM(true, true, true);
    } else {
        if (!ta && !tb) AST_SGEMM(false, false, false);
        else if (!ta && tb) AST_SGEMM(false, true, false);
        else if (ta && !tb) AST_SGEMM(true, false, false);
        else AST_SGEMM(true, true, false);
    }
#undef AST_SGEMM
    AST_LAUNCH_OK();
    return 0;
}

}  // namespace ast

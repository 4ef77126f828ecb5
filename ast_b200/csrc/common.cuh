// Shared device/host helpers for the ast_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ast {

// ---- error plumbing (never throw across the C ABI) ------------------------------------
void set_last_error(const char* fmt, ...);
const char* get_last_error();
// true when a tool that SERIALISES kernel execution is attached (Nsight Compute / compute-sanitizer found in /proc/self/maps, or
// AST_NO_COOP / AST_NO_PERSIST set): kernels that wait for one another (persistent encoder wavefront) would deadlock, and ncu
// rejects cooperative + cluster launches, so both fall back to their per-chunk / non-cooperative launch structure.
bool kernels_are_serialised();
extern unsigned long long g_kernel_launches;   // every kernel this library enqueues (bench.py gpu_launches)

#define AST_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ast::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                                cudaGetErrorString(_e));                                    \
            return -2;                                                                      \
        }                                                                                   \
    } while (0)

#define AST_LAUNCH_OK()                                                                     \
    do {                                                                                    \
        ++ast::g_kernel_launches;                                                           \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            ast::set_last_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,           \
                                cudaGetErrorString(_e));                                    \
            return -2;                                                                      \
        }                                                                                   \
    } while (0)

#define AST_CHECK(cond, ...)                                                                \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ast::set_last_error(__VA_ARGS__);                                               \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

#define AST_TRY(expr)                                                                       \
    do {                                                                                    \
        int _r = (expr);                                                                    \
        if (_r != 0) return _r;                                                             \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __CUDACC__
// ---- warp helpers ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum over blockDim.x threads (blockDim.x multiple of 32, <= 1024).
// `scratch` must hold 32 floats.  All threads receive the result.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// ---- accurate-enough activations (fp32 parity with numpy: use the libm-grade versions) --
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- TF32 tensor-core primitives (mma.sync m16n8k8) --------------------------------------
// Used for the batch-as-M (M = 16) recurrent / decoder GEMMs where a tcgen05 tile (M >= 64)
// would be >= 75 % padding.  The "3x" split (hi*hi + lo*hi + hi*lo) recovers ~fp32 accuracy.
__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
// D(16x8) += A(16x8,row) * B(8x8,col).  Fragment layout (g = lane>>2, q = lane&3):
//   a0=A[g][q] a1=A[g+8][q] a2=A[g][q+4] a3=A[g+8][q+4];  b0=B[q][g] b1=B[q+4][g];
//   c0=C[g][2q] c1=C[g][2q+1] c2=C[g+8][2q] c3=C[g+8][2q+1].
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <bool EXACT>
__device__ __forceinline__ void mma_f32(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
    uint32_t ah[4], bh[2];
    if (EXACT) {
        uint32_t al[4], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
        mma_tf32(c, al, bh);
        mma_tf32(c, ah, bl);
        mma_tf32(c, ah, bh);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ah[i] = f2tf32(a[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) bh[i] = f2tf32(b[i]);
        mma_tf32(c, ah, bh);
    }
}

// ---- bounded spin on a flag another LAUNCH publishes (persistent-wavefront hand-offs) -----------------
// CUDA gives no forward-progress guarantee between separate launches: if the producer never becomes resident (another tenant
// holds the SMs, a stream got serialised behind the waiter) an unbounded spin hangs the GPU until the watchdog / the operator
// kills the process.  After SPIN_TIMEOUT_NS of %globaltimer the waiter traps instead: the launch fails loudly
// (cudaErrorLaunchFailure at the next synchronisation) and the host reports it.  The clock is read every 1024 polls only.
constexpr unsigned long long SPIN_TIMEOUT_NS = 4000000000ULL;
__device__ __forceinline__ void spin_until_ge(const unsigned* flag, unsigned target) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= target) return;
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned polls = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((++polls & 1023u) == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > SPIN_TIMEOUT_NS) __trap();
        }
    } while (v < target);
}

// ---- counter-based RNG for dropout / noise (stateless: mask is re-derivable in backward) ---
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint32_t stream, uint32_t idx) {
    uint32_t a = hash_u32(static_cast<uint32_t>(seed) ^ (stream * 0x9E3779B9U));
    uint32_t b = hash_u32(static_cast<uint32_t>(seed >> 32) + idx * 0x85EBCA6BU + a);
    return hash_u32(a ^ b ^ idx);
}
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint32_t stream, uint32_t idx) {
    return (rng_u32(seed, stream, idx) >> 8) * (1.0f / 16777216.0f);   // [0,1)
}
// N(0, 1) from two counter-RNG draws (Box-Muller); GradientNoise hook
__device__ __forceinline__ float rng_normal(uint64_t seed, uint32_t stream, uint32_t idx) {
    const float u1 = ((rng_u32(seed, stream, idx) >> 8) + 1) * (1.0f / 16777216.0f);          // (0, 1]
    const float u2 = rng_uniform(seed, stream + 1, idx);
    return sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853071795865f * u2);
}
// scaled keep-mask of F.dropout: 0 with prob ratio, else 1/(1-ratio)
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t stream, uint32_t idx, float ratio) {
    if (ratio <= 0.f) return 1.f;
    return rng_uniform(seed, stream, idx) < ratio ? 0.f : 1.f / (1.f - ratio);
}
#endif  // __CUDACC__

}  // namespace ast

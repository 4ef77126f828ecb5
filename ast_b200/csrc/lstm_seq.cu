// Persistent LSTM recurrence over a whole sequence (seq2seq.py:192-225, chainer L.LSTM / F.lstm).
//
// One thread-block CLUSTER of 8 CTAs owns one (layer, direction) chain for all T steps: the lateral
// weight W_h (4h x h fp32, 1 MB at h = 256) is loaded into the cluster's shared memory ONCE
// (128 KB per CTA) and stays resident; per step the cluster does the (B x h)·(h x 4h) recurrent
// GEMM on tensor cores (mma.sync m16n8k8 TF32 with the 3-term split for fp32 accuracy -- batch is
// the M dimension, and M = 16 is exactly one mma tile, so a tcgen05 tile with M >= 64 would be
// >= 75 % padding), fuses the gate non-linearities, the cell update and the dropout mask, and
// all-gathers the new h through distributed shared memory followed by one cluster barrier.
// The input projections X·W_x^T + b for all timesteps are a single batched GEMM done beforehand.
//
// Chainer's interleaved gate layout (row 4j+k of W, k = a,i,f,o) means a contiguous block of 4U
// rows is exactly U hidden units with all four gates, so each CTA owns U = h/8 units outright.
//
// Backward runs the same structure in reverse: dG_t (pre-activation gate gradients) is produced
// elementwise, written in place over the saved activations, multiplied by the resident W_h slice and
// reduce-scattered across the cluster.  Weight gradients are batched GEMMs over the saved dG.
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace ast {

constexpr int NC = 8;            // CTAs per cluster
constexpr int LTHREADS = 256;

template <int MT, bool EXACT>
__global__ void __launch_bounds__(LTHREADS, 1)
lstm_seq_fwd_kernel(LstmChains ch, int T, int B, int h, float drop, unsigned long long seed) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const LstmChain a = ch.c[blockIdx.x / NC];
    const int U = h / NC;                 // hidden units owned by this CTA
    const int ldw = h + 4;                // padded smem row stride (conflict-free fragment loads)
    const int H4 = 4 * h;
    constexpr int MROWS = 16 * MT;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                     // [4U][ldw], rows permuted into (a,i)/(f,o) n-tiles per warp
    float* hb = Ws + (size_t)4 * U * ldw; // [2][MROWS][ldw]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;

    // resident weight slice: smem row p -> (warp p/16, tile (p%16)/8, col p%8) -> unit 4*warp + col/2,
    // gate 2*tile + col%2 ; source row 4*(U*rank + unit) + gate of the (4h x h) lateral matrix.
    const int h4 = h >> 2;
    for (int idx = tid; idx < 4 * U * h4; idx += LTHREADS) {
        const int p = idx / h4, k4 = idx % h4;
        const int wp = p >> 4, s = (p >> 3) & 1, j = p & 7;
        const int unit = 4 * wp + (j >> 1), gate = 2 * s + (j & 1);
        const float4 v = *reinterpret_cast<const float4*>(a.Wl + (size_t)(4 * (U * rank + unit) + gate) * h + k4 * 4);
        *reinterpret_cast<float4*>(Ws + (size_t)p * ldw + k4 * 4) = v;
    }
    // h_{-1} from slot 0 of Hs into buffer 0 (rows >= B stay zero for the whole run)
    for (int idx = tid; idx < 2 * MROWS * ldw; idx += LTHREADS) hb[idx] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < B * h4; idx += LTHREADS) {
        const int m = idx / h4, k4 = idx % h4;
        *reinterpret_cast<float4*>(hb + (size_t)m * ldw + k4 * 4) =
            *reinterpret_cast<const float4*>(a.Hs + (size_t)m * h + k4 * 4);
    }
    const bool wact = w < (U >> 2);       // warps beyond U/4 only take part in the barriers
    const int ju = U * rank + 4 * w + q;  // this thread's hidden unit
    float creg[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int row = mt * 16 + g + 8 * hf;
            creg[mt][hf] = (wact && row < B) ? a.Cs[(size_t)row * h + ju] : 0.f;
        }
    __syncthreads();
    cluster.sync();

    const int ksteps = h >> 3;
    for (int i = 0; i < T; ++i) {
        const float* hc = hb + (size_t)(i & 1) * MROWS * ldw;
        float* hn = hb + (size_t)((i + 1) & 1) * MROWS * ldw;
        float acc[2][MT][4];
        if (wact) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int row = mt * 16 + g + 8 * hf;
                    float4 gx = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < B) gx = *reinterpret_cast<const float4*>(a.G + ((size_t)i * B + row) * H4 + 4 * ju);
                    acc[0][mt][2 * hf] = gx.x; acc[0][mt][2 * hf + 1] = gx.y;
                    acc[1][mt][2 * hf] = gx.z; acc[1][mt][2 * hf + 1] = gx.w;
                }
            const float* w0 = Ws + (size_t)(w * 16 + g) * ldw + q;
            const float* w1 = w0 + (size_t)8 * ldw;
#pragma unroll 4
            for (int ks = 0; ks < ksteps; ++ks) {
                const int k0 = ks * 8;
                const float b0[2] = {w0[k0], w0[k0 + 4]};
                const float b1[2] = {w1[k0], w1[k0 + 4]};
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const float* hr = hc + (size_t)(mt * 16 + g) * ldw + k0 + q;
                    const float af[4] = {hr[0], hr[(size_t)8 * ldw], hr[4], hr[(size_t)8 * ldw + 4]};
                    mma_f32<EXACT>(acc[0][mt], af, b0);
                    mma_f32<EXACT>(acc[1][mt], af, b1);
                }
            }
            // gates -> cell -> h ; write saved activations, states and the all-gather
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int row = mt * 16 + g + 8 * hf;
                    const float ga = tanhf(acc[0][mt][2 * hf]);
                    const float gi = sigmoidf_(acc[0][mt][2 * hf + 1]);
                    const float gf = sigmoidf_(acc[1][mt][2 * hf]);
                    const float go = sigmoidf_(acc[1][mt][2 * hf + 1]);
                    const float c = ga * gi + gf * creg[mt][hf];
                    float hv = go * tanhf(c);
                    if (row < B) {
                        creg[mt][hf] = c;
                        const size_t r = (size_t)i * B + row;
                        *reinterpret_cast<float4*>(a.G + r * H4 + 4 * ju) = make_float4(ga, gi, gf, go);
                        a.Cs[(r + B) * h + ju] = c;
                        a.Hs[(r + B) * h + ju] = hv;
                        const float dm = dropout_scale(seed, a.drop_stream, (uint32_t)(r * h + ju), drop);
                        a.out[(long long)i * a.out_si + (long long)row * a.out_sb + ju] = hv * dm;
                    } else {
                        hv = 0.f;
                    }
                    // quad-gather 4 consecutive units, then each lane pushes the float4 to 2 CTAs
                    const int qb = lane & ~3;
                    float4 v4;
                    v4.x = __shfl_sync(0xffffffffu, hv, qb + 0);
                    v4.y = __shfl_sync(0xffffffffu, hv, qb + 1);
                    v4.z = __shfl_sync(0xffffffffu, hv, qb + 2);
                    v4.w = __shfl_sync(0xffffffffu, hv, qb + 3);
                    float* dst_local = hn + (size_t)row * ldw + U * rank + 4 * w;
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        float* dst = cluster.map_shared_rank(dst_local, 2 * q + d);
                        *reinterpret_cast<float4*>(dst) = v4;
                    }
                }
        }
        cluster.sync();
    }
}

template <int MT, bool EXACT>
__global__ void __launch_bounds__(LTHREADS, 1)
lstm_seq_bwd_kernel(LstmChains ch, int T, int B, int h, float drop, unsigned long long seed) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const LstmChain a = ch.c[blockIdx.x / NC];
    const int U = h / NC;
    const int K4 = 4 * U;                 // gate rows owned by this CTA (contraction length)
    const int ldw = h + 8;                // bank = 8q + g
    const int ldg = K4 + 4;               // bank = 4g + q
    const int H4 = 4 * h;
    constexpr int MROWS = 16 * MT;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                               // [K4][ldw]  natural row order
    float* dgs = Ws + (size_t)K4 * ldw;             // [MROWS][ldg]
    float* red = dgs + (size_t)MROWS * ldg;         // [2][NC][MROWS][U]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
    const int h4 = h >> 2;
    for (int idx = tid; idx < K4 * h4; idx += LTHREADS) {
        const int p = idx / h4, k4 = idx % h4;
        *reinterpret_cast<float4*>(Ws + (size_t)p * ldw + k4 * 4) =
            *reinterpret_cast<const float4*>(a.Wl + (size_t)(K4 * rank + p) * h + k4 * 4);
    }
    for (int idx = tid; idx < MROWS * ldg; idx += LTHREADS) dgs[idx] = 0.f;
    for (int idx = tid; idx < 2 * NC * MROWS * U; idx += LTHREADS) red[idx] = 0.f;

    // elementwise ownership: pair e -> (m, ul) with ul fastest (coalesced)
    const int npairs = MROWS * U;
    constexpr int MAXE = MT * 2;          // MROWS*U/256 <= 16*MT*32/256
    float dc[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int idx = tid + e * LTHREADS;
        const int m = idx / U, ul = idx % U;
        dc[e] = (idx < npairs && m < B && a.dc_fin) ? a.dc_fin[(size_t)m * a.ld_dc_fin + U * rank + ul] : 0.f;
    }
    __syncthreads();
    cluster.sync();

    const int ntile_per_warp = (h >> 3) / (LTHREADS / 32);   // n-tiles (8 cols) per warp
    for (int i = T - 1; i >= 0; --i) {
        const int buf = i & 1;
        const float* rprev = red + (size_t)((i + 1) & 1) * NC * MROWS * U;
        // 1. dG_t for the owned units
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * LTHREADS;
            if (idx < npairs) {
                const int m = idx / U, ul = idx % U;
                const int ju = U * rank + ul;
                float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < B) {
                    float dh;
                    if (i == T - 1) {
                        dh = a.dh_fin ? a.dh_fin[(size_t)m * a.ld_dh_fin + ju] : 0.f;
                    } else {
                        dh = 0.f;
#pragma unroll
                        for (int s = 0; s < NC; ++s) dh += rprev[((size_t)s * MROWS + m) * U + ul];
                    }
                    const size_t r = (size_t)i * B + m;
                    const float dm = dropout_scale(seed, a.drop_stream, (uint32_t)(r * h + ju), drop);
                    dh += a.dout[(long long)i * a.out_si + (long long)m * a.out_sb + ju] * dm;
                    const float4 act = *reinterpret_cast<const float4*>(a.G + r * H4 + 4 * ju);
                    const float c = a.Cs[(r + B) * h + ju], cp = a.Cs[r * h + ju];
                    const float tc = tanhf(c);
                    const float dct = dc[e] + dh * act.w * (1.f - tc * tc);
                    dg.x = dct * act.y * (1.f - act.x * act.x);
                    dg.y = dct * act.x * act.y * (1.f - act.y);
                    dg.z = dct * cp * act.z * (1.f - act.z);
                    dg.w = dh * tc * act.w * (1.f - act.w);
                    dc[e] = dct * act.z;
                    *reinterpret_cast<float4*>(a.G + r * H4 + 4 * ju) = dg;
                }
                *reinterpret_cast<float4*>(dgs + (size_t)m * ldg + 4 * ul) = dg;
            }
        }
        __syncthreads();
        if (i > 0 || a.dh0) {
            // 2. partial dh_{t-1}[m][n] = sum_p dG[m][p] * W[p][n] over this CTA's K4 gate rows
            for (int nt = 0; nt < ntile_per_warp; ++nt) {
                const int n0 = (w * ntile_per_warp + nt) * 8;
                float acc[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[mt][j] = 0.f;
#pragma unroll 4
                for (int ks = 0; ks < (K4 >> 3); ++ks) {
                    const int k0 = ks * 8;
                    const float bf[2] = {Ws[(size_t)(k0 + q) * ldw + n0 + g], Ws[(size_t)(k0 + q + 4) * ldw + n0 + g]};
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const float* dr = dgs + (size_t)(mt * 16 + g) * ldg + k0 + q;
                        const float af[4] = {dr[0], dr[(size_t)8 * ldg], dr[4], dr[(size_t)8 * ldg + 4]};
                        mma_f32<EXACT>(acc[mt], af, bf);
                    }
                }
                // 3. reduce-scatter: columns n0+2q, n0+2q+1 belong to CTA (n/U)
                const int n = n0 + 2 * q;
                const int owner = n / U, ul = n % U;
                float* base = cluster.map_shared_rank(red, owner) + ((size_t)(buf * NC + rank) * MROWS) * U + ul;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    *reinterpret_cast<float2*>(base + (size_t)(mt * 16 + g) * U) = make_float2(acc[mt][0], acc[mt][1]);
                    *reinterpret_cast<float2*>(base + (size_t)(mt * 16 + g + 8) * U) = make_float2(acc[mt][2], acc[mt][3]);
                }
            }
        }
        cluster.sync();
    }
    // gradients w.r.t. the initial state (slot 0), when requested
    if (a.dh0 || a.dc0) {
        const float* r0 = red;            // buffer of step i = 0
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * LTHREADS;
            if (idx < npairs) {
                const int m = idx / U, ul = idx % U;
                if (m < B) {
                    const int ju = U * rank + ul;
                    if (a.dh0) {
                        float s = 0.f;
                        for (int sidx = 0; sidx < NC; ++sidx) s += r0[((size_t)sidx * MROWS + m) * U + ul];
                        a.dh0[(size_t)m * h + ju] = s;
                    }
                    if (a.dc0) a.dc0[(size_t)m * h + ju] = dc[e];
                }
            }
        }
    }
}

static size_t fwd_smem(int h, int MT) { return sizeof(float) * ((size_t)4 * (h / NC) * (h + 4) + (size_t)2 * 16 * MT * (h + 4)); }
static size_t bwd_smem(int h, int MT) {
    const int U = h / NC;
    return sizeof(float) * ((size_t)4 * U * (h + 8) + (size_t)16 * MT * (4 * U + 4) + (size_t)2 * NC * 16 * MT * U);
}

template <class KernT>
static int launch_cluster(KernT kern, cudaStream_t st, int nchains, size_t smem, const LstmChains& ch, int T, int B,
                          int h, float drop, unsigned long long seed) {
    AST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nchains * NC);
    cfg.blockDim = dim3(LTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AST_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ch, T, B, h, drop, seed));
    ++g_kernel_launches;
    return 0;
}

static int check_shape(const char* who, int nchains, int T, int B, int h) {
    AST_CHECK(nchains >= 1 && nchains <= AST_MAX_CHAINS, "%s: nchains %d out of range", who, nchains);
    AST_CHECK(h % 64 == 0 && h >= 64 && h <= 256, "%s: per-direction hidden size %d unsupported (need multiple of 64, <= 256)", who, h);
    AST_CHECK(B >= 1 && B <= 32, "%s: batch %d unsupported by the persistent recurrence (1..32); split the batch", who, B);
    AST_CHECK(T >= 1, "%s: T must be >= 1", who);
    return 0;
}

int lstm_seq_fwd(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop,
                 unsigned long long seed, bool exact) {
    AST_TRY(check_shape("lstm_seq_fwd", nchains, T, B, h));
    const int MT = B <= 16 ? 1 : 2;
    const size_t smem = fwd_smem(h, MT);
    if (MT == 1) return exact ? launch_cluster(lstm_seq_fwd_kernel<1, true>, st, nchains, smem, ch, T, B, h, drop, seed)
                              : launch_cluster(lstm_seq_fwd_kernel<1, false>, st, nchains, smem, ch, T, B, h, drop, seed);
    return exact ? launch_cluster(lstm_seq_fwd_kernel<2, true>, st, nchains, smem, ch, T, B, h, drop, seed)
                 : launch_cluster(lstm_seq_fwd_kernel<2, false>, st, nchains, smem, ch, T, B, h, drop, seed);
}

int lstm_seq_bwd(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop,
                 unsigned long long seed, bool exact) {
    AST_TRY(check_shape("lstm_seq_bwd", nchains, T, B, h));
    const int MT = B <= 16 ? 1 : 2;
    const size_t smem = bwd_smem(h, MT);
    if (MT == 1) return exact ? launch_cluster(lstm_seq_bwd_kernel<1, true>, st, nchains, smem, ch, T, B, h, drop, seed)
                              : launch_cluster(lstm_seq_bwd_kernel<1, false>, st, nchains, smem, ch, T, B, h, drop, seed);
    return exact ? launch_cluster(lstm_seq_bwd_kernel<2, true>, st, nchains, smem, ch, T, B, h, drop, seed)
                 : launch_cluster(lstm_seq_bwd_kernel<2, false>, st, nchains, smem, ch, T, B, h, drop, seed);
}

}  // namespace ast

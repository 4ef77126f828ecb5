// Persistent LSTM recurrence over a whole sequence (seq2seq.py:192-225, chainer L.LSTM / F.lstm).
//
// One thread-block CLUSTER of 8 CTAs owns one chain = (layer, direction, 16-row batch slice) for all T
// steps: the lateral weight W_h (4h x h fp32, 1 MB at h = 256) is loaded into the cluster's shared memory
// ONCE (128 KB per CTA) and stays resident; per step the cluster does the (16 x h)·(h x 4h) recurrent GEMM
// on tensor cores (mma.sync m16n8k8 TF32; the 3-term split gives fp32 accuracy in "exact" mode -- batch is
// the M dimension, and M = 16 is exactly one mma tile, so a tcgen05 tile with M >= 64 would be >= 75 %
// padding), fuses the gate non-linearities, the cell update and the dropout mask, and all-gathers the new h
// through distributed shared memory.
//
// The step time is a pure dependency chain, so the kernel is built around latency (cycle counts measured on
// B200 with clock64, tools/lstm_timing.cu):
//  * dependent mma.sync latency is ~90 cycles -> 4 independent accumulator chains over k (8 instead of 32 long);
//  * the exchange uses st.async + mbarrier complete_tx: the producer fires 16-byte remote stores that signal the
//    CONSUMER's mbarrier when they land, so nobody waits for store acknowledgements or a cluster-wide barrier
//    (barrier.cluster.arrive.release alone cost ~1250 cycles per step); buffers are double-buffered and the
//    data flow itself guarantees a buffer is free before it is overwritten;
//  * global bookkeeping stores (saved activations / states) and the prefetch of the next step's input
//    projection are issued after the sends, off the critical path;
//  * in TF32 mode the gate non-linearities use ex2/rcp based forms (abs. error ~1e-7).
// The input projections X·W_x^T + b for all timesteps are a single batched tcgen05 GEMM done beforehand.
//
// Chainer's interleaved gate layout (row 4j+k of W, k = a,i,f,o) means a contiguous block of 4U rows is
// exactly U hidden units with all four gates, so each CTA owns U = h/8 units outright.
//
// Backward runs the same structure in reverse: dG_t (pre-activation gate gradients) is produced
// elementwise, written in place over the saved activations, multiplied by the resident W_h slice and
// reduce-scattered across the cluster (same st.async / mbarrier scheme).  Weight gradients are batched GEMMs
// over the saved dG.
#include "cluster_dev.cuh"
#include "kernels.h"

namespace ast {

constexpr int NC = 8;            // CTAs per cluster
constexpr int LTHREADS = 256;
constexpr int MROWS = 16;        // batch rows per chain (one mma M tile)

// ---- activations -------------------------------------------------------------------------------------
template <bool EXACT> __device__ __forceinline__ float act_sigmoid(float x) {
    if (EXACT) return 1.f / (1.f + expf(-x));
    return __fdividef(1.f, 1.f + __expf(-x));
}
template <bool EXACT> __device__ __forceinline__ float act_tanh(float x) {
    if (EXACT) return tanhf(x);
    return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f;
}

template <bool EXACT>
struct Frag {            // an mma operand register, optionally with its low-order TF32 term
    uint32_t hi, lo;
    __device__ __forceinline__ void set(float x) {
        hi = f2tf32(x);
        if (EXACT) lo = f2tf32(x - __uint_as_float(hi));
    }
};
template <bool EXACT>
__device__ __forceinline__ void mma_frag(float (&c)[4], const Frag<EXACT> (&a)[4], const Frag<EXACT> (&b)[2]) {
    const uint32_t ah[4] = {a[0].hi, a[1].hi, a[2].hi, a[3].hi};
    const uint32_t bh[2] = {b[0].hi, b[1].hi};
    if (EXACT) {
        const uint32_t al[4] = {a[0].lo, a[1].lo, a[2].lo, a[3].lo};
        const uint32_t bl[2] = {b[0].lo, b[1].lo};
        mma_tf32(c, al, bh);
        mma_tf32(c, ah, bl);
    }
    mma_tf32(c, ah, bh);
}

template <bool EXACT>
__global__ void __launch_bounds__(LTHREADS, 1)
lstm_seq_fwd_kernel(LstmChains ch, int T, int B, int h, float drop, unsigned long long seed) {
    const int rank = (int)cluster_rank();
    const LstmChain a = ch.c[blockIdx.x / NC];
    const int U = h / NC;                 // hidden units owned by this CTA
    const int ldw = h + 4;                // padded smem row stride (conflict-free fragment loads)
    const int H4 = 4 * h;
    const int nb = a.nb, b0 = a.b0;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                     // [4U][ldw], rows permuted into (a,i)/(f,o) n-tiles per warp
    float* hb = Ws + (size_t)4 * U * ldw; // [2][MROWS][ldw]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(hb + (size_t)2 * MROWS * ldw);   // [2], one per h buffer

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
    // h_{-1} from slot 0 of Hs into buffer 0 (rows >= nb stay zero for the whole run)
    for (int idx = tid; idx < 2 * MROWS * ldw; idx += LTHREADS) hb[idx] = 0.f;
    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int idx = tid; idx < nb * h4; idx += LTHREADS) {
        const int m = idx / h4, k4 = idx % h4;
        *reinterpret_cast<float4*>(hb + (size_t)m * ldw + k4 * 4) =
            *reinterpret_cast<const float4*>(a.Hs + (size_t)(b0 + m) * h + k4 * 4);
    }
    const bool wact = w < (U >> 2);       // warps beyond U/4 have no units: they only join the final barrier
    const int ju = U * rank + 4 * w + q;  // this thread's hidden unit
    float creg[2];
    float4 gx[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int row = g + 8 * hf;
        const bool v = wact && row < nb;
        creg[hf] = v ? a.Cs[(size_t)(b0 + row) * h + ju] : 0.f;
        gx[hf] = v ? *reinterpret_cast<const float4*>(a.G + (size_t)(b0 + row) * H4 + 4 * ju) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    cluster_sync_all();                   // every CTA's buffers and mbarriers are initialised before any remote store

    const uint32_t tx_bytes = (uint32_t)(MROWS * h * sizeof(float));   // what each CTA receives per step
    const int ksteps = h >> 3;
    if (wact) {
        for (int i = 0; i < T; ++i) {
            const int cur = i & 1, nxt = cur ^ 1;
            const float* hc = hb + (size_t)cur * MROWS * ldw;
            if (tid == 0 && i + 1 < T) mbar_expect_tx(&mbar[nxt], tx_bytes);     // arm the buffer this step fills
            if (i > 0) mbar_wait(&mbar[cur], ((i - 1) >> 1) & 1);                // all 8 slices of h_{i-1} have landed

            float acc[2][4][4];           // [tile: (a,i) | (f,o)][independent k-chain][mma C fragment]
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[t2][pp][j] = 0.f;
            const float* w0 = Ws + (size_t)(w * 16 + g) * ldw + q;
            const float* w1 = w0 + (size_t)8 * ldw;
            const float* hr = hc + (size_t)g * ldw + q;
            for (int ks = 0; ks < ksteps; ks += 4) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const int k0 = (ks + pp) * 8;
                    Frag<EXACT> fa[4], fb0[2], fb1[2];
                    fa[0].set(hr[k0]); fa[1].set(hr[(size_t)8 * ldw + k0]); fa[2].set(hr[k0 + 4]); fa[3].set(hr[(size_t)8 * ldw + k0 + 4]);
                    fb0[0].set(w0[k0]); fb0[1].set(w0[k0 + 4]);
                    fb1[0].set(w1[k0]); fb1[1].set(w1[k0 + 4]);
                    mma_frag<EXACT>(acc[0][pp], fa, fb0);
                    mma_frag<EXACT>(acc[1][pp], fa, fb1);
                }
            }
            float pre[2][4];
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
                for (int j = 0; j < 4; ++j) pre[t2][j] = (acc[t2][0][j] + acc[t2][1][j]) + (acc[t2][2][j] + acc[t2][3][j]);
            // gates -> cell -> h ; the all-gather goes out first (critical path)
            float4 actv[2]; float cv[2], hv[2];
            const uint32_t hn_local = saddr(hb + (size_t)nxt * MROWS * ldw);
            const uint32_t bar_local = saddr(&mbar[nxt]);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int row = g + 8 * hf;
                const float4 x = gx[hf];
                const float ga = act_tanh<EXACT>(pre[0][2 * hf] + x.x);
                const float gi = act_sigmoid<EXACT>(pre[0][2 * hf + 1] + x.y);
                const float gf = act_sigmoid<EXACT>(pre[1][2 * hf] + x.z);
                const float go = act_sigmoid<EXACT>(pre[1][2 * hf + 1] + x.w);
                const float c = ga * gi + gf * creg[hf];
                const float hval = (row < nb) ? go * act_tanh<EXACT>(c) : 0.f;
                actv[hf] = make_float4(ga, gi, gf, go); cv[hf] = c; hv[hf] = hval;
                if (row < nb) creg[hf] = c;
                if (i + 1 < T) {
                    // quad-gather 4 consecutive units, then each lane fires the 16-byte slice at 2 of the 8 CTAs
                    const int qb = lane & ~3;
                    float4 v4;
                    v4.x = __shfl_sync(0xffffffffu, hval, qb + 0);
                    v4.y = __shfl_sync(0xffffffffu, hval, qb + 1);
                    v4.z = __shfl_sync(0xffffffffu, hval, qb + 2);
                    v4.w = __shfl_sync(0xffffffffu, hval, qb + 3);
                    const uint32_t dst = hn_local + (uint32_t)(((size_t)row * ldw + U * rank + 4 * w) * sizeof(float));
#pragma unroll
                    for (int d = 0; d < 2; ++d) st_async_v4(mapa(dst, 2 * q + d), v4, mapa(bar_local, 2 * q + d));
                }
            }
            // bookkeeping off the critical path: saved activations / states, layer output, next input projection
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int row = g + 8 * hf;
                if (row < nb) {
                    const size_t r = (size_t)i * B + b0 + row;
                    *reinterpret_cast<float4*>(a.G + r * H4 + 4 * ju) = actv[hf];
                    a.Cs[(r + B) * h + ju] = cv[hf];
                    a.Hs[(r + B) * h + ju] = hv[hf];
                    const float dm = dropout_scale(seed, a.drop_stream, (uint32_t)(r * h + ju) + a.drop_off, drop);
                    a.out[(long long)i * a.out_si + (long long)(b0 + row) * a.out_sb + ju] = hv[hf] * dm;
                    if (i + 1 < T) gx[hf] = *reinterpret_cast<const float4*>(a.G + (r + B) * H4 + 4 * ju);   // prefetch
                }
            }
        }
    }
    cluster_sync_all();                   // nobody exits while a peer could still address its shared memory
}

template <bool EXACT>
__global__ void __launch_bounds__(LTHREADS, 1)
lstm_seq_bwd_kernel(LstmChains ch, int T, int B, int h, float drop, unsigned long long seed) {
    const int rank = (int)cluster_rank();
    const LstmChain a = ch.c[blockIdx.x / NC];
    const int U = h / NC;
    const int K4 = 4 * U;                 // gate rows owned by this CTA (contraction length)
    const int ldw = h + 8;                // bank = 8q + g
    const int ldg = K4 + 4;               // bank = 4g + q
    const int H4 = 4 * h;
    const int nb = a.nb, b0 = a.b0;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                               // [K4][ldw]  natural row order
    float* dgs = Ws + (size_t)K4 * ldw;             // [MROWS][ldg]
    float* red = dgs + (size_t)MROWS * ldg;         // [2][NC][MROWS][U]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + (size_t)2 * NC * MROWS * U);   // [2]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
    const int h4 = h >> 2;
    for (int idx = tid; idx < K4 * h4; idx += LTHREADS) {
        const int p = idx / h4, k4 = idx % h4;
        *reinterpret_cast<float4*>(Ws + (size_t)p * ldw + k4 * 4) =
            *reinterpret_cast<const float4*>(a.Wl + (size_t)(K4 * rank + p) * h + k4 * 4);
    }
    for (int idx = tid; idx < MROWS * ldg; idx += LTHREADS) dgs[idx] = 0.f;
    for (int idx = tid; idx < 2 * NC * MROWS * U; idx += LTHREADS) red[idx] = 0.f;
    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // elementwise ownership: pair e -> (m, ul) with ul fastest (coalesced); MROWS*U <= 512 -> <= 2 per thread
    const int npairs = MROWS * U;
    constexpr int MAXE = 2;
    float dc[MAXE];
    float4 p_act[MAXE]; float p_c[MAXE], p_cp[MAXE], p_dout[MAXE];       // prefetched inputs of the current step
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int idx = tid + e * LTHREADS;
        const int m = idx / U, ul = idx % U, ju = U * rank + ul;
        const bool v = idx < npairs && m < nb;
        dc[e] = (v && a.dc_fin) ? a.dc_fin[(size_t)(b0 + m) * a.ld_dc_fin + ju] : 0.f;
        if (v) {
            const size_t r = (size_t)(T - 1) * B + b0 + m;
            p_act[e] = *reinterpret_cast<const float4*>(a.G + r * H4 + 4 * ju);
            p_c[e] = a.Cs[(r + B) * h + ju]; p_cp[e] = a.Cs[r * h + ju];
            p_dout[e] = a.dout[(long long)(T - 1) * a.out_si + (long long)(b0 + m) * a.out_sb + ju];
        } else { p_act[e] = make_float4(0.f, 0.f, 0.f, 0.f); p_c[e] = p_cp[e] = p_dout[e] = 0.f; }
    }
    __syncthreads();
    cluster_sync_all();

    const uint32_t tx_bytes = (uint32_t)(NC * MROWS * U * sizeof(float));
    const int ntpw = (h >> 3) / (LTHREADS / 32);   // n-tiles (8 output columns) per warp: h/64
    for (int i = T - 1; i >= 0; --i) {
        const int buf = i & 1;
        const bool send = (i > 0) || (a.dh0 != nullptr);
        const float* rprev = red + (size_t)(buf ^ 1) * NC * MROWS * U;
        if (tid == 0 && send) mbar_expect_tx(&mbar[buf], tx_bytes);
        if (i < T - 1) mbar_wait(&mbar[buf ^ 1], ((T - 2 - i) >> 1) & 1);      // partial dh of step i+1 from all 8 CTAs
        __syncthreads();                  // every warp is done reading dgs of the previous step
        // 1. dG_t for the owned units
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * LTHREADS;
            if (idx < npairs) {
                const int m = idx / U, ul = idx % U;
                const int ju = U * rank + ul;
                float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < nb) {
                    float dh;
                    if (i == T - 1) {
                        dh = a.dh_fin ? a.dh_fin[(size_t)(b0 + m) * a.ld_dh_fin + ju] : 0.f;
                    } else {
                        dh = 0.f;
#pragma unroll
                        for (int s = 0; s < NC; ++s) dh += rprev[((size_t)s * MROWS + m) * U + ul];
                    }
                    const size_t r = (size_t)i * B + b0 + m;
                    const float dm = dropout_scale(seed, a.drop_stream, (uint32_t)(r * h + ju) + a.drop_off, drop);
                    dh += p_dout[e] * dm;
                    const float4 act = p_act[e];
                    const float c = p_c[e], cp = p_cp[e];
                    const float tc = act_tanh<EXACT>(c);
                    const float dct = dc[e] + dh * act.w * (1.f - tc * tc);
                    dg.x = dct * act.y * (1.f - act.x * act.x);
                    dg.y = dct * act.x * act.y * (1.f - act.y);
                    dg.z = dct * cp * act.z * (1.f - act.z);
                    dg.w = dh * tc * act.w * (1.f - act.w);
                    dc[e] = dct * act.z;
                    p_act[e] = dg;        // written back to HBM after the sends
                }
                *reinterpret_cast<float4*>(dgs + (size_t)m * ldg + 4 * ul) = dg;
            }
        }
        __syncthreads();
        if (send) {
            // 2. partial dh_{t-1}[m][n] = sum_p dG[m][p] * W[p][n] over this CTA's K4 gate rows, two n-tiles x two
            //    independent k-chains at a time; 3. reduce-scatter to the owner of columns n (CTA n / U)
            const float* dr = dgs + (size_t)g * ldg + q;
            const uint32_t red_local = saddr(red + ((size_t)(buf * NC + rank) * MROWS) * U);
            const uint32_t bar_local = saddr(&mbar[buf]);
            for (int nt0 = 0; nt0 < ntpw; nt0 += 2) {
                const int n0a = (w * ntpw + nt0) * 8, n0b = n0a + 8;
                const bool has_b = nt0 + 1 < ntpw;
                float acc[2][2][4];
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[t2][pp][j] = 0.f;
                for (int ks = 0; ks < (K4 >> 3); ks += 2) {
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const int k0 = (ks + pp) * 8;
                        Frag<EXACT> fa[4], fb[2];
                        fa[0].set(dr[k0]); fa[1].set(dr[(size_t)8 * ldg + k0]); fa[2].set(dr[k0 + 4]); fa[3].set(dr[(size_t)8 * ldg + k0 + 4]);
                        fb[0].set(Ws[(size_t)(k0 + q) * ldw + n0a + g]); fb[1].set(Ws[(size_t)(k0 + q + 4) * ldw + n0a + g]);
                        mma_frag<EXACT>(acc[0][pp], fa, fb);
                        if (has_b) {
                            fb[0].set(Ws[(size_t)(k0 + q) * ldw + n0b + g]); fb[1].set(Ws[(size_t)(k0 + q + 4) * ldw + n0b + g]);
                            mma_frag<EXACT>(acc[1][pp], fa, fb);
                        }
                    }
                }
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    if (t2 == 1 && !has_b) break;
                    const int n = (t2 ? n0b : n0a) + 2 * q;
                    const int owner = n / U, ul = n % U;
                    const uint32_t rb = mapa(bar_local, owner);
                    const uint32_t base = mapa(red_local + (uint32_t)(ul * sizeof(float)), owner);
                    st_async_v2(base + (uint32_t)((size_t)g * U * sizeof(float)),
                                make_float2(acc[t2][0][0] + acc[t2][1][0], acc[t2][0][1] + acc[t2][1][1]), rb);
                    st_async_v2(base + (uint32_t)((size_t)(g + 8) * U * sizeof(float)),
                                make_float2(acc[t2][0][2] + acc[t2][1][2], acc[t2][0][3] + acc[t2][1][3]), rb);
                }
            }
        }
        // 4. off the critical path: write dG_t in place, prefetch step i-1
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * LTHREADS;
            const int m = idx / U, ul = idx % U, ju = U * rank + ul;
            if (idx < npairs && m < nb) {
                const size_t r = (size_t)i * B + b0 + m;
                *reinterpret_cast<float4*>(a.G + r * H4 + 4 * ju) = p_act[e];
                if (i > 0) {
                    const size_t rp = r - B;
                    p_act[e] = *reinterpret_cast<const float4*>(a.G + rp * H4 + 4 * ju);
                    p_c[e] = p_cp[e]; p_cp[e] = a.Cs[rp * h + ju];
                    p_dout[e] = a.dout[(long long)(i - 1) * a.out_si + (long long)(b0 + m) * a.out_sb + ju];
                }
            }
        }
    }
    // gradients w.r.t. the initial state (slot 0), when requested
    if (a.dh0 || a.dc0) {
        if (a.dh0) mbar_wait(&mbar[0], ((T - 1) >> 1) & 1);
        const float* r0 = red;            // buffer of step i = 0
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * LTHREADS;
            if (idx < npairs) {
                const int m = idx / U, ul = idx % U;
                if (m < nb) {
                    const int ju = U * rank + ul;
                    if (a.dh0) {
                        float s = 0.f;
                        for (int sidx = 0; sidx < NC; ++sidx) s += r0[((size_t)sidx * MROWS + m) * U + ul];
                        a.dh0[(size_t)(b0 + m) * h + ju] = s;
                    }
                    if (a.dc0) a.dc0[(size_t)(b0 + m) * h + ju] = dc[e];
                }
            }
        }
    }
    cluster_sync_all();
}

static size_t fwd_smem(int h) { return sizeof(float) * ((size_t)4 * (h / NC) * (h + 4) + (size_t)2 * MROWS * (h + 4)) + 16; }
static size_t bwd_smem(int h) {
    const int U = h / NC;
    return sizeof(float) * ((size_t)4 * U * (h + 8) + (size_t)MROWS * (4 * U + 4) + (size_t)2 * NC * MROWS * U) + 16;
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

// Split every logical chain (covering rows [b0, b0+nb)) into 16-row chains: batch rows are independent, so a
// larger batch simply uses more clusters at the same per-step latency.
static int expand_chains(const char* who, const LstmChains& in, int nchains, int T, int B, int h, LstmChains& out, int& nout) {
    AST_CHECK(h % 64 == 0 && h >= 64 && h <= 256, "%s: per-direction hidden size %d unsupported (need multiple of 64, <= 256)", who, h);
    AST_CHECK(T >= 1 && B >= 1, "%s: T and B must be >= 1", who);
    nout = 0;
    for (int c = 0; c < nchains; ++c) {
        const int nb_total = in.c[c].nb > 0 ? in.c[c].nb : B;
        for (int r0 = 0; r0 < nb_total; r0 += MROWS) {
            AST_CHECK(nout < AST_MAX_CHAINS, "%s: too many 16-row chains (batch %d x %d chains > %d clusters per launch)", who, B, nchains, AST_MAX_CHAINS);
            out.c[nout] = in.c[c];
            out.c[nout].b0 = in.c[c].b0 + r0;
            out.c[nout].nb = std::min(MROWS, nb_total - r0);
            ++nout;
        }
    }
    return 0;
}

int lstm_seq_fwd(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop,
                 unsigned long long seed, bool exact) {
    LstmChains ex{}; int n = 0;
    AST_TRY(expand_chains("lstm_seq_fwd", ch, nchains, T, B, h, ex, n));
    if (!exact && h == 256) return lstm_seq_fwd_tc(st, ex, n, T, B, drop, seed);     // tcgen05 path (TF32 training mode)
    const size_t smem = fwd_smem(h);
    return exact ? launch_cluster(lstm_seq_fwd_kernel<true>, st, n, smem, ex, T, B, h, drop, seed)
                 : launch_cluster(lstm_seq_fwd_kernel<false>, st, n, smem, ex, T, B, h, drop, seed);
}

bool lstm_seq_gated_supported(int h, bool exact) { return !exact && h == 256; }
int lstm_seq_fwd_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop, unsigned long long seed,
                       const LstmGate& gate, int* ncta) {
    AST_CHECK(lstm_seq_gated_supported(h, false), "lstm_seq_fwd_gated: only the tcgen05 recurrence (h == 256) can be gated");
    LstmChains ex{}; int n = 0;
    AST_TRY(expand_chains("lstm_seq_fwd_gated", ch, nchains, T, B, h, ex, n));
    if (ncta) *ncta = n * 8;
    return lstm_seq_fwd_tc_gated(st, ex, n, T, B, drop, seed, gate);
}
int lstm_seq_bwd_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop, unsigned long long seed,
                       const LstmGate& gate, int* ncta) {
    AST_CHECK(lstm_seq_gated_supported(h, false), "lstm_seq_bwd_gated: only the tcgen05 recurrence (h == 256) can be gated");
    LstmChains ex{}; int n = 0;
    AST_TRY(expand_chains("lstm_seq_bwd_gated", ch, nchains, T, B, h, ex, n));
    if (ncta) *ncta = n * 8;
    return lstm_seq_bwd_tc_gated(st, ex, n, T, B, drop, seed, gate);
}

int lstm_seq_bwd(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, int h, float drop,
                 unsigned long long seed, bool exact) {
    LstmChains ex{}; int n = 0;
    AST_TRY(expand_chains("lstm_seq_bwd", ch, nchains, T, B, h, ex, n));
    bool wants_init_grads = false;
    for (int c = 0; c < n; ++c) wants_init_grads |= (ex.c[c].dh0 != nullptr) || (ex.c[c].dc0 != nullptr);
    (void)wants_init_grads;
    if (!exact && h == 256) return lstm_seq_bwd_tc(st, ex, n, T, B, drop, seed);
    const size_t smem = bwd_smem(h);
    return exact ? launch_cluster(lstm_seq_bwd_kernel<true>, st, n, smem, ex, T, B, h, drop, seed)
                 : launch_cluster(lstm_seq_bwd_kernel<false>, st, n, smem, ex, T, B, h, drop, seed);
}

}  // namespace ast

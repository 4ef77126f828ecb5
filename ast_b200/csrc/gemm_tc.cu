// TF32 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma.kind::tf32 (one elected thread issues) -> fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
//   C[m,n] (+)= sum_k opA(m,k) * opB(k,n) (+ bias[n])          row-major fp32 everywhere
//   opA(m,k) = ta ? A[k*lda + m] : A[m*lda + k]     (ta = 0: K-major,  ta = 1: M-major operand)
//   opB(k,n) = tb ? B[n*ldb + k] : B[k*ldb + n]     (tb = 1: K-major,  tb = 0: N-major operand)
//
// fp32 data is consumed directly as TF32 (the tensor core reads the top 19 bits), so no conversion pass
// touches HBM.  Both operand majors are expressed through the UMMA instruction descriptor (a_major /
// b_major), which lets the backward data-gradient (NN) and weight-gradient (TN) contractions read the
// forward tensors in place -- no transposed copies.  Weight gradients have K = T'*B (thousands) and a
// small output, so they run split-K over blockIdx.z with fp32 red.global.add into a zeroed C.
//
// S3 variant ("3xTF32", fp32-faithful): C = hi(x).hi(y) + lo(x).hi(y) + hi(x).lo(y) with hi = rna_tf32(v),
// lo = rna_tf32(v - hi) precomputed by split_tf32(); a k-block stages four tiles and issues three MMAs per k-step.  Used for the forward convolutions that feed train-mode BatchNorm.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue
// (TMEM lane quadrant = warp_id % 4).  Tile 128 x 128 x 32 (one 128-byte swizzle row of fp32 per k-block),
// 5-stage mbarrier ring (32 KB per stage).
#include <cuda.h>
#include <map>
#include <tuple>
#include "common.cuh"
#include "kernels.h"

namespace ast {

constexpr int TBM = 128, TBN = 128, TBK = 32, TSTAGES = 5;
constexpr int TC_THREADS = 192;
constexpr uint32_t STAGE_A_BYTES = TBM * TBK * 4, STAGE_B_BYTES = TBN * TBK * 4;
constexpr uint32_t TC_STG_BYTES = 4 * 32 * 33 * 4;      // epilogue staging (one 32x33 tile per epilogue warp), outside the ring
constexpr uint32_t TC_SMEM = TSTAGES * (STAGE_A_BYTES + STAGE_B_BYTES) + TC_STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start address, leading /
// stride byte offsets (all >> 4), version 1, layout type.
//   K-major fp32 : SWIZZLE_128B (2): rows of 128 B (32 k), 16-byte chunks XOR-ed with (row % 8); SBO = 8 rows.
//   MN-major fp32: SWIZZLE_128B_BASE32B (1) is the only layout the tensor core accepts for 32-bit MN-major
//                  operands: rows of 128 B (32 mn) per k, 32-byte chunks XOR-ed with (k-row % 4); SBO = 4 k-rows,
//                  LBO = stride between 32-wide MN chunks.  TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;           // version = 1 (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}

struct TcParams {
    int M, N, K;
    float* C; int ldc;
    const float* bias;
    float beta;
    int atomic;          // split-K: red.add into C
    int kb_per_split;    // k-blocks per split
    int splits;          // number of K splits (part of the linearised tile space)
    // optional gating (persistent encoder wavefront; kernels.h::TcGate): the A rows of chunk q = row / gate_rows exist once
    // gate_wait[q] >= gate_target; every epilogue warp adds 1 to gate_done[m-tile] after its stores of a tile are visible
    const unsigned* gate_wait; unsigned gate_target; int gate_B, gate_chunk, gate_T, gate_rev; unsigned* gate_done;
};

// TA: A operand is M-major (A stored K x M).  NB: B operand is N-major (B stored K x N).
// PERSISTENT: gridDim.x CTAs walk the linearised (split, m-tile, n-tile) space with stride gridDim.x (n fastest, so
// concurrently running CTAs share the A row-tile in L2).  The smem ring runs across tile boundaries, and the fp32
// accumulator is double-buffered in TMEM: while the epilogue warps drain tile i (tcgen05.ld -> smem transpose -> coalesced
// stores), the MMA warp is already accumulating tile i+1 in the other buffer.  With K of only 256..1536 the prologue and
// epilogue were ~1/3 of a tile's time in the one-tile-per-CTA version.
template <bool TA, bool NB, bool S3>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapAl, const __grid_constant__ CUtensorMap mapBl, TcParams p) {
    constexpr int NST = S3 ? 3 : TSTAGES;
    constexpr uint32_t ACC_COLS = S3 ? 2 * TBN : TBN;        // S3: main + correction accumulator
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;             // double-buffered
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + NST * STAGE_A_BYTES;
    uint8_t* sAl = smem + NST * (STAGE_A_BYTES + STAGE_B_BYTES);          // S3 only: low parts
    uint8_t* sBl = sAl + NST * STAGE_A_BYTES;
    float* stg_base = reinterpret_cast<float*>(smem + (S3 ? 2 : 1) * NST * (STAGE_A_BYTES + STAGE_B_BYTES));   // 4 x 32 x 33 floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg_base) + TC_STG_BYTES);
    uint64_t* full = bars;                 // [NST]
    uint64_t* empty = bars + NST;          // [NST]
    uint64_t* tmem_full = bars + 2 * NST;  // [2]
    uint64_t* tmem_empty = bars + 2 * NST + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb_total = (p.K + TBK - 1) / TBK;
    const int tiles_n = (p.N + TBN - 1) / TBN, tiles_m = (p.M + TBM - 1) / TBM;
    const int tiles_mn = tiles_m * tiles_n;
    const int ntiles = tiles_mn * p.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int z = t / tiles_mn, mn = t - z * tiles_mn;
                const int m0 = (p.gate_rev ? tiles_m - 1 - mn / tiles_n : mn / tiles_n) * TBM, n0 = (mn % tiles_n) * TBN;
                const int kb0 = z * p.kb_per_split;
                const int nkb = min(p.kb_per_split, nkb_total - kb0);
                if (p.gate_wait) {       // the producer kernel is still running: spin until it has published this tile's last row
                    // rows are (step, batch) pairs; the producer counts chunks of gate_chunk steps in ITS processing order
                    // (forward in time, or - gate_rev - backward from step gate_T - 1)
                    const int q = p.gate_rev ? (p.gate_T - 1 - m0 / p.gate_B) / p.gate_chunk
                                             : ((min(m0 + TBM, p.M) - 1) / p.gate_B) / p.gate_chunk;
                    spin_until_ge(p.gate_wait + q, p.gate_target);
                    asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy writes (other SMs) -> this thread's TMA reads
                }
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (S3 ? 2 : 1) * (STAGE_A_BYTES + STAGE_B_BYTES));
                    const int k0 = (kb0 + i) * TBK;
                    uint8_t* a = sA + s * STAGE_A_BYTES;
                    uint8_t* b = sB + s * STAGE_B_BYTES;
                    if (!TA) tma_load_2d(&mapA, &full[s], a, k0, m0);                     // box {32 k, 128 m}
                    else
#pragma unroll
                        for (int j = 0; j < TBM / 32; ++j) tma_load_2d(&mapA, &full[s], a + j * (TBK * 128), m0 + 32 * j, k0);   // box {32 m, 32 k}
                    if (!NB) tma_load_2d(&mapB, &full[s], b, k0, n0);
                    else
#pragma unroll
                        for (int j = 0; j < TBN / 32; ++j) tma_load_2d(&mapB, &full[s], b + j * (TBK * 128), n0 + 32 * j, k0);
                    if (S3) {          // K-major operands only (host enforces)
                        tma_load_2d(&mapAl, &full[s], sAl + s * STAGE_A_BYTES, k0, m0);
                        tma_load_2d(&mapBl, &full[s], sBl + s * STAGE_B_BYTES, k0, n0);
                    }
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, majors, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TA ? 1u : 0u) << 15) | ((NB ? 1u : 0u) << 16) |
                                   ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            // Base descriptors are built once; per MMA only the 14-bit start-address field advances (an add of an
            // immediate): building descriptors inside the loop cost ~66 cycles per MMA in this single dependent thread.
            // K-major: advance 32 B inside the 128 B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: each k-step is the next 8-row group (1024 B); LBO = stride between 32-wide MN chunks.
            const uint64_t a_base = TA ? make_smem_desc(smem_u32(sA), TBK * 128, 512, 1) : make_smem_desc(smem_u32(sA), 16, 1024, 2);
            const uint64_t b_base = NB ? make_smem_desc(smem_u32(sB), TBK * 128, 512, 1) : make_smem_desc(smem_u32(sB), 16, 1024, 2);
            constexpr uint32_t a_kstep = (TA ? 1024 : 32) >> 4, b_kstep = (NB ? 1024 : 32) >> 4;
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int z = t / tiles_mn;
                const int nkb = min(p.kb_per_split, nkb_total - z * p.kb_per_split);
                const int buf = it & 1;
                const uint32_t acc = tmem_base + buf * ACC_COLS;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);        // the epilogue has drained this buffer (first use: free)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t ad0 = a_base + (uint64_t)(s * (STAGE_A_BYTES >> 4));
                    const uint64_t bd0 = b_base + (uint64_t)(s * (STAGE_B_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < TBK / 8; ++k) {
                        umma_tf32(acc, ad0 + (uint64_t)(k * a_kstep), bd0 + (uint64_t)(k * b_kstep), idesc, (i > 0 || k > 0) ? 1u : 0u);
                        if (S3) {
                            const uint64_t al0 = ad0 + (uint64_t)((NST * (STAGE_A_BYTES + STAGE_B_BYTES)) >> 4);
                            const uint64_t bl0 = bd0 + (uint64_t)((NST * (STAGE_A_BYTES + STAGE_B_BYTES)) >> 4);
                            // the correction terms get their own accumulator (columns TBN..2*TBN-1): the tensor core's fp32
                            // accumulate truncates, and its error scales with the accumulator's magnitude - the big hi.hi sum
                            // sees a third of the adds, the 2^-11-sized correction sum contributes nothing measurable
                            umma_tf32(acc + TBN, al0 + (uint64_t)(k * a_kstep), bd0 + (uint64_t)(k * b_kstep), idesc, (i > 0 || k > 0) ? 1u : 0u);
                            umma_tf32(acc + TBN, ad0 + (uint64_t)(k * a_kstep), bl0 + (uint64_t)(k * b_kstep), idesc, 1u);
                        }
                    }
                    umma_commit(&empty[s]);          // frees the smem slot when these MMAs retire
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[buf]);        // accumulator of this tile complete
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> smem transpose -> global =====
        const int wq = warp & 3;                 // TMEM lane quadrant this warp may access
        // Each thread holds one accumulator ROW (32 columns per tcgen05.ld); a per-warp 32x33 smem transpose turns the
        // stores into full 128-byte row segments.
        float* stg = stg_base + wq * (32 * 33);
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int m0 = (p.gate_rev ? tiles_m - 1 - mn / tiles_n : mn / tiles_n) * TBM, n0 = (mn % tiles_n) * TBN;
            const int nkb = min(p.kb_per_split, nkb_total - z * p.kb_per_split);
            const int buf = it & 1;
            const uint32_t acc = tmem_base + buf * ACC_COLS + ((uint32_t)(wq * 32) << 16);
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < TBN / 32; ++c) {
                float v[32];
                tmem_ld32(acc + c * 32, v);
                if (S3) {
                    float v2[32];
                    tmem_ld32(acc + TBN + c * 32, v2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += v2[j];
                }
                if (c == TBN / 32 - 1) {         // last read of this buffer: hand it back to the MMA warp before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
                __syncwarp();
                const int n = n0 + c * 32 + lane;
                if (n < p.N && nkb > 0) {
                    const float bv = (p.bias && z == 0) ? p.bias[n] : 0.f;
                    const int mrow0 = m0 + wq * 32;
                    const int rmax = min(32, p.M - mrow0);
                    float* cp = p.C + (size_t)mrow0 * p.ldc + n;
                    for (int r = 0; r < rmax; ++r, cp += p.ldc) {
                        const float x = stg[r * 33 + lane] + bv;
                        // beta == 1 goes through the same fire-and-forget reduction as split-K: a load -> add -> store per row is a
                        // chain of 32 dependent L2 round trips per column chunk (250 us for a 4-tile CTA, measured)
                        if (p.atomic || p.beta == 1.f) atomicAdd(cp, x);
                        else *cp = (p.beta != 0.f) ? x + p.beta * (*cp) : x;
                    }
                }
                __syncwarp();
            }
            if (p.gate_done) {           // this warp's 32 rows of the tile are in global memory
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(p.gate_done + m0 / TBM, 1u);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// =================================================================================================================
// 2-CTA variant (cta_group::2): a CTA PAIR (cluster of 2, two SMs of one TPC) computes a 256 x 256 tile.  Each CTA stages
// its own 128 rows of A and its own 128 columns of B (32 KB per k-block instead of 32 KB for a 128 x 128 tile: twice the
// FLOPs per byte moved into and read out of shared memory, which is what bounds the fp32-operand 1-CTA kernel), the leader
// CTA's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 8) for both tensor cores, and each CTA keeps
// its 128 x 256 half of the fp32 accumulator in its own TMEM (double-buffered: all 512 columns).
//   barriers:  full[s]       leader only; 1 arrival (the leader's expect_tx of BOTH CTAs' bytes), both CTAs' TMA complete on it
//              empty[s]      per CTA; tcgen05.commit multicast to both CTAs when the MMAs that read the slot have retired
//              tmem_full[b]  per CTA; commit multicast when a tile's accumulator is complete
//              tmem_empty[b] leader only; 8 arrivals (4 epilogue warps x 2 CTAs, the peer's through mapa)
// All operand majors, split-K (weight gradients); no 3xTF32 (the 1-CTA kernel keeps that).
constexpr int T2_BN = 256, T2_STAGES = 6;
constexpr uint32_t T2_STAGE_BYTES = STAGE_A_BYTES + (T2_BN / 2) * TBK * 4;       // per CTA: 16 KB of A + 16 KB of B
constexpr uint32_t T2_SMEM = T2_STAGES * T2_STAGE_BYTES + TC_STG_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync2() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair into its OWN shared memory, completing bytes on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// Up to T2_MAXG same-shape problems in one launch ("grouped"): the tile space becomes (problem, split, m-tile, n-tile).  The encoder's
// ten 1024 x 256 x (T'B) weight gradients have four 256 x 256 tiles each: alone they need ~18 K-splits to fill the GPU (9 k-blocks
// per CTA, 18 atomic passes over the output); four at a time need 4-5.
constexpr int T2_MAXG = 4;
struct Tc2Maps { CUtensorMap a[T2_MAXG]; CUtensorMap b[T2_MAXG]; };
struct Tc2Group { int n; float* C[T2_MAXG]; const float* bias[T2_MAXG]; };

template <bool TA, bool NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ Tc2Maps maps, TcParams p, Tc2Group grp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + T2_STAGES * STAGE_A_BYTES;
    float* stg_base = reinterpret_cast<float*>(smem + T2_STAGES * T2_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg_base) + TC_STG_BYTES);
    uint64_t* full = bars;                       // [T2_STAGES]
    uint64_t* empty = bars + T2_STAGES;          // [T2_STAGES]
    uint64_t* tmem_full = bars + 2 * T2_STAGES;  // [2]
    uint64_t* tmem_empty = bars + 2 * T2_STAGES + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T2_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int nkb_total = (p.K + TBK - 1) / TBK;
    const int tiles_n = (p.N + T2_BN - 1) / T2_BN, tiles_m = (p.M + 2 * TBM - 1) / (2 * TBM);
    const int tiles_mn = tiles_m * tiles_n;
    const int tiles_p = tiles_mn * p.splits;     // split-K: (split, m-tile, n-tile), fp32 red.add into a zeroed C
    const int ntiles = tiles_p * grp.n;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {         // collective over the pair: warp 1 of both CTAs, same slot offset
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync2();         // both CTAs' barriers exist before any remote completion / arrival
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 rows of A, own 128 columns of B =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                const int g = t / tiles_p, tp = t - g * tiles_p;
                const CUtensorMap* mapA = &maps.a[g];
                const CUtensorMap* mapB = &maps.b[g];
                const int z = tp / tiles_mn, mn = tp - z * tiles_mn;
                const int m0 = (mn / tiles_n) * (2 * TBM) + (int)rank * TBM;
                const int n0 = (mn % tiles_n) * T2_BN + (int)rank * (T2_BN / 2);
                const int kb0 = z * p.kb_per_split;
                const int nkb = min(p.kb_per_split, nkb_total - kb0);
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (leader) mbar_expect_tx(&full[s], 2 * T2_STAGE_BYTES);
                    const uint32_t lbar = mapa_u32(smem_u32(&full[s]), 0);
                    const int k0 = (kb0 + i) * TBK;
                    uint8_t* a = sA + s * STAGE_A_BYTES;
                    uint8_t* b = sB + s * ((T2_BN / 2) * TBK * 4);
                    if (!TA) tma_load_2d_pair(mapA, lbar, a, k0, m0);                // box {32 k, 128 m}
                    else
#pragma unroll
                        for (int j = 0; j < TBM / 32; ++j) tma_load_2d_pair(mapA, lbar, a + j * (TBK * 128), m0 + 32 * j, k0);   // box {32 m, 32 k}
                    if (!NB) tma_load_2d_pair(mapB, lbar, b, k0, n0);                // box {32 k, 128 n}
                    else
#pragma unroll
                        for (int j = 0; j < T2_BN / 2 / 32; ++j) tma_load_2d_pair(mapB, lbar, b + j * (TBK * 128), n0 + 32 * j, k0);   // box {32 n, 32 k}
                    if (++s == T2_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's elected thread drives both tensor cores =====
        if (leader && lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TA ? 1u : 0u) << 15) | ((NB ? 1u : 0u) << 16) |
                                   ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)((2 * TBM) >> 4) << 24);
            const uint64_t a_base = TA ? make_smem_desc(smem_u32(sA), TBK * 128, 512, 1) : make_smem_desc(smem_u32(sA), 16, 1024, 2);
            const uint64_t b_base = NB ? make_smem_desc(smem_u32(sB), TBK * 128, 512, 1) : make_smem_desc(smem_u32(sB), 16, 1024, 2);
            constexpr uint32_t a_kstep = (TA ? 1024 : 32) >> 4, b_kstep = (NB ? 1024 : 32) >> 4;
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int t = pair; t < ntiles; t += npairs, ++it) {
                const int z = (t % tiles_p) / tiles_mn;
                const int nkb = min(p.kb_per_split, nkb_total - z * p.kb_per_split);
                const int buf = it & 1;
                const uint32_t acc = tmem_base + buf * T2_BN;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t ad0 = a_base + (uint64_t)(s * (STAGE_A_BYTES >> 4));
                    const uint64_t bd0 = b_base + (uint64_t)(s * (((T2_BN / 2) * TBK * 4) >> 4));
#pragma unroll
                    for (int k = 0; k < TBK / 8; ++k)
                        umma_tf32_pair(acc, ad0 + (uint64_t)(k * a_kstep), bd0 + (uint64_t)(k * b_kstep), idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(&empty[s]);
                    if (++s == T2_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_pair(&tmem_full[buf]);
            }
        }
    } else {
        // ===== epilogue (both CTAs): own 128 rows x 256 columns =====
        const int wq = warp & 3;
        float* stg = stg_base + wq * (32 * 33);
        int it = 0;
        for (int t = pair; t < ntiles; t += npairs, ++it) {
            const int g = t / tiles_p, tp = t - g * tiles_p;
            float* Cg = grp.C[g];
            const float* biasg = grp.bias[g];
            const int z = tp / tiles_mn, mn = tp - z * tiles_mn;
            const int m0 = (mn / tiles_n) * (2 * TBM) + (int)rank * TBM, n0 = (mn % tiles_n) * T2_BN;
            const int buf = it & 1;
            const uint32_t acc = tmem_base + buf * T2_BN + ((uint32_t)(wq * 32) << 16);
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < T2_BN / 32; ++c) {
                float v[32];
                tmem_ld32(acc + c * 32, v);
                if (c == T2_BN / 32 - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(&tmem_empty[buf]), 0)) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
                __syncwarp();
                const int n = n0 + c * 32 + lane;
                if (n < p.N && p.atomic != 2) {
                    const float bv = (biasg && z == 0) ? biasg[n] : 0.f;
                    const int mrow0 = m0 + wq * 32;
                    const int rmax = min(32, p.M - mrow0);
                    float* cp = Cg + (size_t)mrow0 * p.ldc + n;
                    for (int r = 0; r < rmax; ++r, cp += p.ldc) {
                        const float x = stg[r * 33 + lane] + bv;
                        if (p.atomic == 1 || p.beta == 1.f) atomicAdd(cp, x);
                        else *cp = (p.beta != 0.f) ? x + p.beta * (*cp) : x;
                    }
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync2();         // the peer may still be reading its accumulator / our barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map: dims {inner, outer}, row stride ld (floats), box {32, box_outer}, 128B swizzle, zero OOB fill.
static bool make_map(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, int ld, uint32_t box_outer, bool mn_major) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    if ((uintptr_t)ptr % 16 != 0 || ld % 4 != 0) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, box_outer};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Grid cap for gemm_tc launches (0 = none): GEMMs that are off the critical path and run beside the persistent encoder wavefront
// must leave the recurrence clusters their SMs - a 148-CTA grid of single CTAs re-takes every SM that frees up and starves the
// 8-CTA cluster launches (stream priorities do not reserve SMs for a pending cluster).
static int g_tc_cta_cap = 0;
void gemm_tc_set_cta_cap(int cap) { g_tc_cta_cap = cap; }
static int tc_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

template <bool TA, bool NB, bool S3>
static int launch_tc(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mal, const CUtensorMap& mbl,
                     const TcParams& p, dim3 grid) {
    auto kern = gemm_tc_kernel<TA, NB, S3>;
    constexpr uint32_t smem = S3 ? (2 * 3 * (STAGE_A_BYTES + STAGE_B_BYTES) + TC_STG_BYTES + 1024 + 256) : TC_SMEM;
    static bool attr_set = false;
    if (!attr_set) {
        AST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<grid, TC_THREADS, smem, st>>>(ma, mb, mal, mbl, p);
    AST_LAUNCH_OK();
    return 0;
}

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 h, l;
        h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
        l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
}
// hi = rna_tf32(x), lo = rna_tf32(x - hi): both exactly representable in TF32, so the tensor core's truncation of its
// fp32 inputs is exact and the remaining error (the dropped lo.lo term and the rounding of lo) is unbiased, ~2^-23.
// (Feeding raw x and lo = x - trunc(x) instead leaves truncation biases that add up linearly in K: measured 5e-5.)
int split_tf32(cudaStream_t st, const float* x, float* hi, float* lo, size_t n) {
    AST_CHECK(n % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)hi % 16 == 0 && (uintptr_t)lo % 16 == 0,
              "split_tf32: need 16-byte aligned buffers, n %% 4 == 0");
    const size_t n4 = n / 4;
    const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
    split_tf32_kernel<<<std::max(blocks, 1), 256, 0, st>>>(x, hi, lo, n4);
    AST_LAUNCH_OK();
    return 0;
}

// C = A . B^T (+bias) with fp32-faithful 3xTF32 (NT form only).  (A, Alo) / (B, Blo) = split_tf32 of the operands, same layouts.
int gemm_tc3_nt(cudaStream_t st, int M, int N, int K, const float* A, const float* Alo, int lda, const float* B, const float* Blo, int ldb,
                float* C, int ldc, const float* bias, bool allow_split) {
    if (M <= 0 || N <= 0 || K <= 0) return 1;
    CUtensorMap ma, mb, mal, mbl;
    const bool ok = make_map(&ma, A, (uint64_t)K, (uint64_t)M, lda, TBM, false) && make_map(&mal, Alo, (uint64_t)K, (uint64_t)M, lda, TBM, false) &&
                    make_map(&mb, B, (uint64_t)K, (uint64_t)N, ldb, TBN, false) && make_map(&mbl, Blo, (uint64_t)K, (uint64_t)N, ldb, TBN, false);
    if (!ok) return 1;
    const int nkb = cdiv(K, TBK);
    // few output tiles (a decode step: M = utterances x hypotheses <= 320 rows): split K so that about one wave of CTAs shares the
    // operand traffic (a 48-tile launch with K = 1152 was bound by each CTA's own 2.3 MB of TMA loads: 23 us), partial sums by
    // red.add onto a zeroed C, the bias added by split 0
    const int tiles = cdiv(N, TBN) * cdiv(M, TBM);
    int splits = 1;
    static const bool no_split = getenv("AST_TC3_NOSPLIT") != nullptr;      // diagnostics
    if (allow_split && !no_split && tiles * 2 <= tc_num_sms()) splits = std::max(1, std::min(nkb / 4, tc_num_sms() / tiles));
    const int kbps = cdiv(nkb, splits);
    splits = cdiv(nkb, kbps);
    if (splits > 1) AST_CUDA_OK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
    TcParams p{M, N, K, C, ldc, bias, 0.f, splits > 1 ? 1 : 0, kbps, splits, nullptr, 0u, 1, 1, 0, 0, nullptr};
    dim3 grid(std::min(tiles * splits, tc_num_sms()));
    return launch_tc<false, false, true>(st, ma, mb, mal, mbl, p, grid);
}

// General entry: returns 1 when the tensor-core path cannot take the problem (caller falls back to SIMT).
int gemm_tc(cudaStream_t st, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
            int ldc, const float* bias, float beta, int split_k) {
    if (M <= 0 || N <= 0 || K <= 0) return 1;
    CUtensorMap ma, mb;
    const bool okA = ta ? make_map(&ma, A, (uint64_t)M, (uint64_t)K, lda, 32, true) : make_map(&ma, A, (uint64_t)K, (uint64_t)M, lda, TBM, false);
    const bool okB = tb ? make_map(&mb, B, (uint64_t)K, (uint64_t)N, ldb, TBN, false) : make_map(&mb, B, (uint64_t)N, (uint64_t)K, ldb, 32, true);
    if (!okA || !okB) return 1;
    const int nkb = cdiv(K, TBK);
    const int tiles = cdiv(M, TBM) * cdiv(N, TBN);
    int splits = 1;
    if (split_k != 0) {
        // fill ~2 waves of 148 SMs, keep >= 8 k-blocks per split
        splits = split_k > 0 ? split_k : std::max(1, std::min(nkb / 8, (2 * 148) / std::max(tiles, 1)));
        splits = std::max(1, std::min(splits, nkb));
    }
    const int kbps = cdiv(nkb, splits);
    splits = cdiv(nkb, kbps);
    TcParams p{M, N, K, C, ldc, bias, beta, splits > 1 ? 1 : 0, kbps, splits, nullptr, 0u, 1, 1, 0, 0, nullptr};
    if (splits > 1) {
        if (beta == 0.f) AST_CUDA_OK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
        else if (beta != 1.f) return 1;
    }
    dim3 grid(std::min(cdiv(N, TBN) * cdiv(M, TBM) * splits, g_tc_cta_cap > 0 ? std::min(g_tc_cta_cap, tc_num_sms()) : tc_num_sms()));
    if (!ta && tb) return launch_tc<false, false, false>(st, ma, mb, ma, mb, p, grid);
    if (!ta && !tb) return launch_tc<false, true, false>(st, ma, mb, ma, mb, p, grid);
    if (ta && !tb) return launch_tc<true, true, false>(st, ma, mb, ma, mb, p, grid);
    return launch_tc<true, false, false>(st, ma, mb, ma, mb, p, grid);
}

// C = A . op(B) + bias (tb: B^T, K-major weight; !tb: N-major weight) as ONE small persistent launch (`ctas` CTAs) that runs beside the kernel PRODUCING A: tiles are walked in
// row order, the TMA warp waits for gate.wait[row chunk] >= gate.target before it reads a tile's rows, and each finished tile
// counts 4 (epilogue warps) into gate.done[m-tile]; the consumer waits for 4 * tiles_per_row() there.  Returns 1 if the TMA path
// cannot take the operands (the caller must then not use the gated scheme).
int gemm_tc_tiles_per_row(int N) { return cdiv(N, TBN); }
int gemm_tc_gated(cudaStream_t st, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* bias, const TcGate& gate, int ctas) {
    if (M <= 0 || N <= 0 || K <= 0) return 1;
    CUtensorMap ma, mb;
    const bool okB = tb ? make_map(&mb, B, (uint64_t)K, (uint64_t)N, ldb, TBN, false) : make_map(&mb, B, (uint64_t)N, (uint64_t)K, ldb, 32, true);
    if (!make_map(&ma, A, (uint64_t)K, (uint64_t)M, lda, TBM, false) || !okB) return 1;
    TcParams p{M, N, K, C, ldc, bias, 0.f, 0, cdiv(K, TBK), 1, gate.wait, gate.target, gate.B, gate.chunk, gate.T, gate.rev ? 1 : 0, gate.done};
    dim3 grid(std::max(1, std::min(cdiv(N, TBN) * cdiv(M, TBM), ctas)));
    if (tb) return launch_tc<false, false, false>(st, ma, mb, ma, mb, p, grid);
    return launch_tc<false, true, false>(st, ma, mb, ma, mb, p, grid);
}

// 2-CTA launch: returns 1 when the pair kernel does not apply (the caller then uses the 1-CTA kernel).
template <bool TA, bool NB>
static int launch_tc2(cudaStream_t st, const Tc2Maps& maps, const TcParams& p, const Tc2Group& grp, int pairs) {
    auto kern = gemm_tc2_kernel<TA, NB>;
    static bool attr_set = false;
    if (!attr_set) {
        AST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = T2_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AST_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, p, grp));
    ++g_kernel_launches;
    return 0;
}
// n same-shape problems C[g] = op(A[g]) . op(B[g]) in one launch (n = 1: the plain GEMM).  split_k: 0 = none, -1 = automatic (one
// wave of CTA pairs, >= 8 k-blocks per split), > 0 = that many.  bias[g] (or null) per problem.
int gemm_tc2_grouped(cudaStream_t st, int n, bool ta, bool tb, int M, int N, int K, const float* const* A, int lda, const float* const* B, int ldb,
                     float* const* C, int ldc, const float* const* bias, float beta, int split_k) {
    if (M <= 0 || N <= 0 || K <= 0 || n < 1 || n > T2_MAXG) return 1;
    Tc2Maps maps;
    Tc2Group grp{};
    grp.n = n;
    for (int g = 0; g < T2_MAXG; ++g) {
        const int q = g < n ? g : 0;
        const bool okA = ta ? make_map(&maps.a[g], A[q], (uint64_t)M, (uint64_t)K, lda, 32, true) : make_map(&maps.a[g], A[q], (uint64_t)K, (uint64_t)M, lda, TBM, false);
        const bool okB = tb ? make_map(&maps.b[g], B[q], (uint64_t)K, (uint64_t)N, ldb, T2_BN / 2, false) : make_map(&maps.b[g], B[q], (uint64_t)N, (uint64_t)K, ldb, 32, true);
        if (!okA || !okB) return 1;
        grp.C[g] = C[q];
        grp.bias[g] = bias ? bias[q] : nullptr;
    }
    int cap = tc_num_sms() / 2;
    if (g_tc_cta_cap > 0) cap = std::max(1, std::min(cap, g_tc_cta_cap / 2));
    const int nkb = cdiv(K, TBK);
    const int tiles = cdiv(M, 2 * TBM) * cdiv(N, T2_BN);
    int splits = 1;
    if (split_k != 0) {
        splits = split_k > 0 ? split_k : std::max(1, std::min(nkb / 8, (tc_num_sms() / 2) / std::max(tiles * n, 1)));
        splits = std::max(1, std::min(splits, nkb));
    }
    const int kbps = cdiv(nkb, splits);
    splits = cdiv(nkb, kbps);
    if (splits > 1) {
        if (beta == 0.f) { for (int g = 0; g < n; ++g) AST_CUDA_OK(cudaMemset2DAsync(C[g], sizeof(float) * ldc, 0, sizeof(float) * N, M, st)); }
        else if (beta != 1.f) return 1;
    }
    static const bool nostore = getenv("AST_TC2_NOSTORE") != nullptr;      // diagnostics: epilogue without its global stores
    const int atomic = nostore ? 2 : (splits > 1 ? 1 : 0);
    TcParams p{M, N, K, C[0], ldc, nullptr, beta, atomic, kbps, splits, nullptr, 0u, 1, 1, 0, 0, nullptr};
    const int pairs = std::min(tiles * splits * n, cap);
    if (!ta && tb) return launch_tc2<false, false>(st, maps, p, grp, pairs);
    if (!ta && !tb) return launch_tc2<false, true>(st, maps, p, grp, pairs);
    if (ta && !tb) return launch_tc2<true, true>(st, maps, p, grp, pairs);
    return launch_tc2<true, false>(st, maps, p, grp, pairs);
}
int gemm_tc2(cudaStream_t st, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             const float* bias, float beta, int split_k) {
    return gemm_tc2_grouped(st, 1, ta, tb, M, N, K, &A, lda, &B, ldb, &C, ldc, bias ? &bias : nullptr, beta, split_k);
}

int gemm_tc_nt(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
               const float* bias, float beta) {
    return gemm_tc(st, false, true, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, 0);
}

}  // namespace ast

// Persistent LSTM recurrence on 5th-generation tensor cores (training mode, per-direction h = 256); seq2seq.py:293-315 is the
// reference loop this replaces (L.NStepBiLSTM per layer), nn.py:176-199 its backward through Chainer's autograd.
//
// Same cluster organisation as lstm_seq.cu (8 CTAs per chain, each owning 32 hidden units = 128 gate rows, W_h resident in
// shared memory for all T steps), but the per-step GEMM is issued as tcgen05.mma with the operands SWAPPED: gates^T (128 gate
// rows x 16 batch rows) = W_slice (128 x 256, the M side, K-major, resident) · h^T (256 x 16, the N side).  The fp32
// accumulator lives in TMEM (16 columns) and is read back with tcgen05.ld for the fused gate / cell / dropout epilogue.
//
// A 128 x 16 x K instruction costs ~60 cycles whatever K is (measured: from smem or TMEM, with 1, 2 or 4 accumulators), so the
// operands are FP16 (kind::f16, K = 16: 16 instructions per step instead of TF32's 32) - the same 11-bit significand; see
// FW_F16 / BW_F16 below for the range argument (forward) and the exact per-step scale (backward).  -DFW_OPERAND_TF32 /
// -DBW_OPERAND_TF32 build the TF32 variants (SWIZZLE_128B operands, M-major W^T in SWIZZLE_128B_BASE32B for backward).
//
// Warp roles: warps 0..7 = epilogue (gates, cell, sends, bookkeeping); warp 8 = MMA issuer (nothing else, so the tensor core
// starts the moment the last slice of h lands; descriptors precomputed and advanced by immediates); then the chunk signaller
// (forward warp 9, backward warp 10: publishes a finished chunk device-wide, see chunk_arrive / signaller_loop) and, backward,
// the loader warp 9 (cp.async prefetch of the next steps' operands into a two-deep shared-memory ring).
//
// Operands are rounded to nearest when they are written to shared memory (W once, h / dG by the producing CTA), so the tensor
// core never truncates.  tools/enc_step_probe.py prints the cycle breakdown of a step inside a real training step.
#include "cluster_dev.cuh"
#include <cuda_fp16.h>
#include "kernels.h"

namespace ast {

__global__ void wait_resident_kernel(const unsigned* counter, unsigned target) { spin_until_ge(counter, target); }
int wait_resident(cudaStream_t st, const unsigned* counter, unsigned target) {
    wait_resident_kernel<<<1, 1, 0, st>>>(counter, target);
    AST_LAUNCH_OK();
    return 0;
}

// optional cycle probe (tools/lstm_step_probe.py): CTA 0 stores clock64() stamps of steps 8..23 of the forward kernel
static unsigned long long* g_lstm_prof = nullptr;
void lstm_tc_set_prof(unsigned long long* p) { g_lstm_prof = p; }

constexpr int TNC = 8;           // CTAs per cluster
constexpr int TH = 256;          // per-direction hidden size handled by this kernel
constexpr int TU = TH / TNC;     // 32 hidden units per CTA
constexpr int TROWS = 16;        // batch rows per chain = UMMA N
constexpr int TC_EPI = 256;      // epilogue threads
constexpr int TC_THREADS = TC_EPI + 64;       // forward: epilogue warps 0..7, MMA issuer warp 8, chunk signaller warp 9

__device__ __forceinline__ float rnd_tf32(float x) { return __uint_as_float(f2tf32(x)); }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }
// D[tmem] (+)= A[tmem, lane = row m, column = k] . B[smem]: the weight operand read from TENSOR MEMORY instead of shared memory
// (cute::SM100_MMA_TF32_TS).  With the weight on the M side (swap-AB) every step re-reads the whole 128 KB slice; from shared
// memory that alone was ~1900 cycles of a 4740-cycle step (tools/lstm_step_probe.py).
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI) : "memory"); }   // epilogue warps only
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Chunk hand-off through a dedicated signaller warp.  The epilogue warps only count their arrival in shared memory (release at CTA
// scope, one lane per warp after __syncwarp); the signaller acquires the count, makes the chunk's global stores visible device-wide
// with ONE gpu-scope fence (cumulative over everything it acquired) and bumps the chunk's counter.  That fence waits for every
// outstanding store of the SM: executed by the 256 epilogue threads (+ a CTA barrier) it cost 1.4 us (forward) / 3.3 us (backward)
// per chunk of 8 steps on the recurrence's critical path (tools/enc_timeline.py against tools/enc_step_probe.py).
__device__ __forceinline__ void chunk_arrive(uint32_t cnt_addr, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(cnt_addr), "r"(1u) : "memory");
}
__device__ __forceinline__ void signaller_loop(uint32_t cnt_addr, int nchunks, const LstmGate& gt, bool stamp) {
    for (int c = 0; c < nchunks; ++c) {
        const unsigned target = (unsigned)(TC_EPI / 32) * (unsigned)(c + 1);
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(cnt_addr) : "memory");
            if (v >= target) break;
            __nanosleep(64);
        }
        if (stamp) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); gt.ts[c] = t; }
        if (gt.done) { __threadfence(); atomicAdd(gt.done + c, 1u); }
    }
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B operand made of k-blocks of 32 floats:
// [kb][rows][128 B], 16-byte chunks XOR-ed with (row % 8)
__device__ __forceinline__ uint32_t kmajor_off(int row, int k, int rows_per_block) {
    const int kb = k >> 5, c = (k & 31) >> 2;
    return (uint32_t)(kb * rows_per_block * 128 + row * 128 + ((c ^ (row & 7)) << 4) + ((k & 3) << 2));
}

// the same for 2-byte operands: K-major SWIZZLE_64B, k-blocks of 32 elements = [rows][64 B], 16-byte chunks XOR-ed with (row / 2) % 4
__device__ __forceinline__ uint32_t kmajor_off_h(int row, int k, int rows_per_block) {
    const int kb = k >> 5, c = (k & 31) >> 3;
    return (uint32_t)(kb * rows_per_block * 64 + row * 64 + ((c ^ ((row >> 1) & 3)) << 4) + ((k & 7) << 1));
}
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
    const __half2 hh = __floats2half2_rn(x, y);
    return *reinterpret_cast<const uint32_t*>(&hh);
}

// explicit shared-state-space accesses: the carve-up pointers come from an aligned-up integer, through which the compiler can only
// emit generic LD/ST
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }

// ================================================================================================
// forward
// ================================================================================================
// Operand type of the forward recurrence.  FP16 carries the same 11-bit significand as TF32 (h is in (-1, 1), recurrent weights are
// far inside +-65504, values under 6e-5 keep an absolute error of 3e-8), but one tcgen05.mma covers K = 16 instead of 8: the 128x16xK
// instruction costs ~60 cycles whatever K is (tools/enc_step_probe.py: 32 of them = 1920 of a step's 4050 cycles, independent
// accumulators change nothing), so 16 instructions instead of 32 - and the h all-gather moves half the bytes.
#ifdef FW_OPERAND_TF32
constexpr bool FW_F16 = false;
#else
constexpr bool FW_F16 = true;
#endif
constexpr uint32_t FW_ROWB = FW_F16 ? 64 : 128;         // bytes of one operand row inside a k-block of 32 hidden units
constexpr uint32_t FW_A_BYTES = 8 * 128 * FW_ROWB;      // W slice: 8 k-blocks x 128 gate rows
constexpr uint32_t FW_TMEM_D = 0;                       // accumulator column
constexpr uint32_t FW_TMEM_COLS = 32;
constexpr uint32_t FW_H_BYTES = 8 * TROWS * FW_ROWB;    // one h buffer: 8 k-blocks x 16 rows
constexpr uint32_t FW_XG_BYTES = 4 * TROWS * TU * 4;    // gate exchange [gate][batch][unit]
constexpr uint32_t FW_STG_BYTES = 2 * TROWS * FW_ROWB;  // [2] staging of this CTA's h slice (= one k-block of the operand layout)
constexpr uint32_t FW_SMEM = FW_A_BYTES + 2 * FW_H_BYTES + FW_XG_BYTES + FW_STG_BYTES + 64 + 1024;

__global__ void __launch_bounds__(TC_THREADS, 1)
lstm_seq_fwd_tc_kernel(LstmChains ch, int T, int B, float drop, unsigned long long seed, unsigned long long* prof, LstmGate gt) {
    const bool probe = prof != nullptr && blockIdx.x == 0 && ch.c[0].drop_stream <= 1;      // stand-alone launches (0) or (layer 0, forward direction)
    // stamps stay in registers until the end of the step: a global store in front of a fence would be waited for by that fence
    uint32_t pst[8];
#define PROBE(slot) do { if (probe) pst[slot] = (uint32_t)clock(); } while (0)
#define PROBE_FLUSH(first, last) do { if (probe && i >= 8 && i < 24) { for (int q_ = (first); q_ <= (last); ++q_) prof[(i - 8) * 8 + q_] = pst[q_]; } } while (0)
    const int rank = (int)cluster_rank();
    if (gt.resident && threadIdx.x == 0) atomicAdd(gt.resident, 1u);       // this CTA holds its SM from here on
    const LstmChain a = ch.c[blockIdx.x / TNC];
    constexpr int h = TH, H4 = 4 * TH;
    const int nb = a.nb, b0 = a.b0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                                  // resident W slice, rows p = 32*gate + unit
    uint8_t* sH = sm + FW_A_BYTES;                     // [2] h buffers (UMMA B operand)
    float* xg = reinterpret_cast<float*>(sm + FW_A_BYTES + 2 * FW_H_BYTES);
    uint8_t* sStg = sm + FW_A_BYTES + 2 * FW_H_BYTES + FW_XG_BYTES;
    uint64_t* mbar_h = reinterpret_cast<uint64_t*>(sm + FW_A_BYTES + 2 * FW_H_BYTES + FW_XG_BYTES + FW_STG_BYTES);   // [2]
    uint64_t* mbar_mma = mbar_h + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_h + 3);
    uint32_t* chunk_cnt = tmem_slot + 1;               // epilogue warps that finished the current chunk's stores (cumulative)

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    for (int idx = tid; idx < (int)(2 * FW_H_BYTES / 16); idx += TC_THREADS) reinterpret_cast<float4*>(sH)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        mbar_init(&mbar_h[0], 1); mbar_init(&mbar_h[1], 1); mbar_init(mbar_mma, 1);
        *chunk_cnt = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(tmem_slot)), "n"(FW_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // Resident weight slice in shared memory: row p = 32*gate + unit  <-  W_lat row 4*(32*rank + unit) + gate, rounded to TF32.
    // (The operand was also tried in TENSOR MEMORY - umma_tf32_ts, A read from TMEM lanes: numerically identical, but the
    // issue cost of the 32 small MMAs stayed at ~56 cycles each, and a 512-column allocation made co-resident GEMM CTAs
    // stall in tcgen05.alloc for the length of a chunk.)  Loads are issued in batches of 8 per thread before the first
    // store: a load -> round -> store loop serialises on L2 latency.
    for (int base = 0; base < 128 * (h / 4); base += 8 * TC_THREADS) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * TC_THREADS + tid;
            if (idx < 128 * (h / 4)) {
                const int p = idx / (h / 4), k4 = idx % (h / 4);
                const int gate = p >> 5, unit = p & 31;
                v[u] = __ldg(reinterpret_cast<const float4*>(a.Wl + (size_t)(4 * (TU * rank + unit) + gate) * h + k4 * 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * TC_THREADS + tid;
            if (idx < 128 * (h / 4)) {
                const int p = idx / (h / 4), k4 = idx % (h / 4);
                if constexpr (FW_F16)
                    *reinterpret_cast<uint2*>(sA + kmajor_off_h(p, 4 * k4, 128)) =
                        make_uint2(pack_h2(v[u].x, v[u].y), pack_h2(v[u].z, v[u].w));
                else
                    *reinterpret_cast<float4*>(sA + kmajor_off(p, 4 * k4, 128)) =
                        make_float4(rnd_tf32(v[u].x), rnd_tf32(v[u].y), rnd_tf32(v[u].z), rnd_tf32(v[u].w));
            }
        }
    }
    for (int idx = tid; idx < nb * (h / 4); idx += TC_THREADS) {          // h_{-1} (slot 0 of Hs) into buffer 0
        const int m = idx / (h / 4), k4 = idx % (h / 4);
        float4 v = *reinterpret_cast<const float4*>(a.Hs + (size_t)(b0 + m) * h + k4 * 4);
        if constexpr (FW_F16) {
            *reinterpret_cast<uint2*>(sH + kmajor_off_h(m, 4 * k4, TROWS)) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
        } else {
            v.x = rnd_tf32(v.x); v.y = rnd_tf32(v.y); v.z = rnd_tf32(v.z); v.w = rnd_tf32(v.w);
            *reinterpret_cast<float4*>(sH + kmajor_off(m, 4 * k4, TROWS)) = v;
        }
    }
    fence_proxy_async();                  // generic-proxy writes of the operands -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sH_addr = saddr(sH), xg_addr = saddr(xg), sStg_addr = saddr(sStg), cnt_addr = saddr(chunk_cnt);
    cluster_sync_all();                   // all CTAs have initialised buffers / barriers before any remote store

    if (w == 8) {
        // ===== MMA issuer: one thread, nothing else on its plate =====
        if (lane == 0) {
            const uint32_t idesc = FW_F16 ? umma_idesc_f16(128, TROWS, false, false) : umma_idesc_tf32(128, TROWS, false, false);
            const uint64_t a0 = umma_smem_desc(saddr(sA), 16, 8 * FW_ROWB, FW_F16 ? 4 : 2);       // SWIZZLE_64B / SWIZZLE_128B, 8-row groups
            const uint64_t b0d = umma_smem_desc(sH_addr, 16, 8 * FW_ROWB, FW_F16 ? 4 : 2);
            for (int i = 0; i < T; ++i) {
                const int cur = i & 1;
                if (i + 1 < T) mbar_expect_tx(&mbar_h[cur ^ 1], FW_H_BYTES);    // arm the buffer this step fills
                if (i > 0) mbar_wait(&mbar_h[cur], ((i - 1) >> 1) & 1);         // all 8 slices of h_{i-1} have landed
                PROBE(0);
                fence_proxy_async();
                tc_fence_after();
                const uint64_t bcur = b0d + (uint64_t)((cur * FW_H_BYTES) >> 4);
                constexpr int KS = FW_ROWB / 32;             // instructions per k-block: 32 bytes of K each
#pragma unroll
                for (int kb = 0; kb < 8; ++kb)
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint64_t ad = a0 + (uint64_t)((kb * (128 * FW_ROWB) + ks * 32) >> 4);
                        const uint64_t bd = bcur + (uint64_t)((kb * (TROWS * FW_ROWB) + ks * 32) >> 4);
                        if constexpr (FW_F16) umma_f16_ss(tmem_base, ad, bd, idesc, (kb | ks) ? 1u : 0u);
                        else umma_tf32_ss(tmem_base, ad, bd, idesc, (kb | ks) ? 1u : 0u);
                    }
                umma_commit_arrive(mbar_mma);
                PROBE(1);
                PROBE_FLUSH(0, 1);
            }
        }
    } else if (w == 9) {
        if (lane == 0 && (gt.done || (gt.ts && blockIdx.x == 0)))
            signaller_loop(cnt_addr, (T + gt.chunk - 1) / gt.chunk, gt, gt.ts && blockIdx.x == 0);
    } else {
        // ===== epilogue: thread (w, lane) owns hidden unit `lane`, batch rows 2w and 2w+1 =====
        const int ju = TU * rank + lane;
        float creg[2];
        float4 gx[2];
        int tiles_ok = 0;          // leading 128-row tiles of G known complete (a.tile_ready gating)
        // rows of step i handled here: [i*B + b0, i*B + b0 + nb) -> needs every tile up to the one holding the last row
#define TILE_WAIT(step) do { if (a.tile_ready) { const int need = ((step) * B + b0 + nb - 1) >> 7; \
            while (tiles_ok <= need) { spin_until_ge(a.tile_ready + tiles_ok, a.tile_target); ++tiles_ok; } } } while (0)
        TILE_WAIT(0);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = 2 * w + j;
            creg[j] = b < nb ? a.Cs[(size_t)(b0 + b) * h + ju] : 0.f;
            gx[j] = b < nb ? __ldcg(reinterpret_cast<const float4*>(a.G + (size_t)(b0 + b) * H4 + 4 * ju)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int gate = w & 3, bh = w >> 2;          // TMEM quadrant = gate; this warp reads batch columns 8*bh..+7
        for (int i = 0; i < T; ++i) {
            const int nxt = (i & 1) ^ 1;
            mbar_wait(mbar_mma, i & 1);
            if (tid == 0) PROBE(2);
            tc_fence_after();
            {
                float v[8];
                tmem_ld8(tmem_base + ((uint32_t)(32 * gate) << 16) + FW_TMEM_D + 8 * bh, v);
                tc_fence_before();
#pragma unroll
                for (int b = 0; b < 8; ++b) sts_f1(xg_addr + (uint32_t)(((gate * TROWS + 8 * bh + b) * TU + lane) * 4), v[b]);
            }
            epi_barrier();
            if (tid == 0) PROBE(3);
            float4 actv[2]; float cv[2], hv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = 2 * w + j;
                const float ga = fast_tanh(lds_f1(xg_addr + (uint32_t)(((0 * TROWS + b) * TU + lane) * 4)) + gx[j].x);
                const float gi = fast_sigmoid(lds_f1(xg_addr + (uint32_t)(((1 * TROWS + b) * TU + lane) * 4)) + gx[j].y);
                const float gf = fast_sigmoid(lds_f1(xg_addr + (uint32_t)(((2 * TROWS + b) * TU + lane) * 4)) + gx[j].z);
                const float go = fast_sigmoid(lds_f1(xg_addr + (uint32_t)(((3 * TROWS + b) * TU + lane) * 4)) + gx[j].w);
                const float c = ga * gi + gf * creg[j];
                const float hval = b < nb ? go * fast_tanh(c) : 0.f;
                actv[j] = make_float4(ga, gi, gf, go); cv[j] = c; hv[j] = hval;
                if (b < nb) creg[j] = c;
            }
            if (tid == 0) PROBE(4);
            if (i + 1 < T) {
                // This CTA's 32 units are exactly k-block `rank` of the next step's K-major operand: a contiguous 2 KB block
                // [16 rows][128 B, 16-byte chunks XOR-ed with row % 8].  Stage it locally (TF32-rounded) and push it to all 8
                // CTAs with ONE bulk copy each (cp.async.bulk smem -> dsmem, completing bytes on the receiver's mbarrier)
                // instead of 4 st.async + 8 mapa per thread (measured: the send section was 1340 of a step's 4500 cycles).
                const uint32_t stg = sStg_addr + (uint32_t)(nxt * (TROWS * FW_ROWB));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int b = 2 * w + j;
                    if constexpr (FW_F16) {
                        const unsigned short hb = __half_as_ushort(__float2half_rn(hv[j]));
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)(b * 64 + ((((lane >> 3) ^ ((b >> 1) & 3))) << 4) + ((lane & 7) << 1))), "h"(hb) : "memory");
                    } else {
                        sts_f1(stg + (uint32_t)(b * 128 + ((((lane >> 2) ^ (b & 7))) << 4) + ((lane & 3) << 2)), rnd_tf32(hv[j]));
                    }
                }
                fence_proxy_async();
                epi_barrier();
                if (tid < TNC) {
                    const uint32_t dst = mapa(sH_addr + nxt * FW_H_BYTES + (uint32_t)rank * (TROWS * FW_ROWB), tid);
                    const uint32_t bar = mapa(saddr(&mbar_h[nxt]), tid);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "r"(stg), "r"((uint32_t)(TROWS * FW_ROWB)), "r"(bar) : "memory");
                }
            }
            if (tid == 0) PROBE(5);
            // bookkeeping: overlaps the other CTAs' sends and the next step's MMA
            if (i + 1 < T) TILE_WAIT(i + 1);      // x-projection of step i+1 published?
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = 2 * w + j;
                if (b < nb) {
                    const size_t r = (size_t)i * B + b0 + b;
                    *reinterpret_cast<float4*>(a.G + r * H4 + 4 * ju) = actv[j];
                    a.Cs[(r + B) * h + ju] = cv[j];
                    a.Hs[(r + B) * h + ju] = hv[j];
                    const float dm = dropout_scale(seed, a.drop_stream, (uint32_t)(r * h + ju) + a.drop_off, drop);
                    a.out[(long long)i * a.out_si + (long long)(b0 + b) * a.out_sb + ju] = hv[j] * dm;
                    if (i + 1 < T) gx[j] = __ldcg(reinterpret_cast<const float4*>(a.G + (r + B) * H4 + 4 * ju));
                }
            }
            if ((i + 1) % gt.chunk == 0 || i + 1 == T) chunk_arrive(cnt_addr, lane);
            if (tid == 0) { PROBE(6); PROBE_FLUSH(2, 6); }
        }
    }
#undef PROBE
#undef PROBE_FLUSH
#undef TILE_WAIT
    tc_fence_before();
    __syncthreads();
    if (w == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FW_TMEM_COLS));
    cluster_sync_all();
}

// ================================================================================================
// backward
// ================================================================================================
// Operand type of the backward recurrence.  FP16 (K = 16 per instruction: 16 MMAs instead of 32, see the forward kernel) needs a
// range guard for dG: every step the CTA takes the exact maximum |dG| of its operand (warp redux + shared atomic + one barrier of
// the epilogue warps), scales by the power of two that puts it in [256, 512) and un-scales its partial dh after the accumulator
// read - exact, so the operand keeps TF32's 11-bit significand for everything within 2^-22 of the maximum and an absolute floor
// 2^-33 of it below.  The weight sits transposed (K-major, k = gate row) so that both operands use the plain SWIZZLE_64B layout.
#ifdef BW_OPERAND_TF32
constexpr bool BW_F16 = false;
#else
constexpr bool BW_F16 = true;
#endif
constexpr uint32_t BW_A_BYTES = BW_F16 ? 4 * 256 * 64   // W^T: 4 k-blocks (32 gate rows) x 256 unit rows x 64 B (K-major)
                                       : 8 * 128 * 128; // W slice as M-major operand: 8 unit-chunks x 128 gate rows x 128 B
constexpr uint32_t BW_G_BYTES = 4 * TROWS * (BW_F16 ? 64 : 128);   // dG: 4 k-blocks x 16 batch rows (K-major)
constexpr uint32_t BW_R_BYTES = TNC * TU * TROWS * 4;   // one reduce buffer [src][unit][batch]
constexpr uint32_t BW_STG_BYTES = 2 * TNC * TU * TROWS * 4;   // [2][owner] staging of the partial dh blocks (2 KB per owner)
constexpr int BW_THREADS = TC_EPI + 96;                 // epilogue warps 0..7, MMA issuer warp 8, loader warp 9, chunk signaller warp 10
constexpr int BW_PAIRS = TU * TROWS;                    // (unit, batch row) pairs of a CTA = 2 per epilogue thread
constexpr uint32_t BW_PRE_BYTES = 2 * BW_PAIRS * (16 + 4 + 4);   // [2] prefetched (gate activations, c_{t-1}, dout) of a step
constexpr uint32_t BW_SMEM = BW_A_BYTES + BW_G_BYTES + 2 * BW_R_BYTES + BW_STG_BYTES + BW_PRE_BYTES + 128 + 1024;
// slot of pair idx = 16*unit + row inside a prefetch buffer: the loader warp writes with lane = unit (fixed row), the epilogue reads
// with consecutive idx - the XOR keeps both sides off the same banks
// asynchronous global -> shared copies: no register staging, so the loader warp keeps several steps in flight
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ int pre_slot4(int idx) { return idx ^ ((idx >> 4) & 7); }
__device__ __forceinline__ int pre_slot1(int idx) { return idx ^ ((idx >> 4) & 15); }

__global__ void __launch_bounds__(BW_THREADS, 1)
lstm_seq_bwd_tc_kernel(LstmChains ch, int T, int B, float drop, unsigned long long seed, LstmGate gt) {
    const int rank = (int)cluster_rank();
    if (gt.resident && threadIdx.x == 0) atomicAdd(gt.resident, 1u);       // this CTA holds its SM from here on
    const LstmChain a = ch.c[blockIdx.x / TNC];
    // cycle probe (tools/lstm_step_probe.py --bwd): CTA 0 of the launch whose first chain is (layer 0, forward direction) stamps
    // clock64 for its 9th..24th processed step into probe[128 + 8 k + slot]
    unsigned long long* const bprof = (gt.probe && blockIdx.x == 0 && a.drop_stream == 1) ? gt.probe + 128 : nullptr;
    // layout: [16 steps][16] = epilogue thread 0 slots 0..7, epilogue thread 224 (warp 7) slots 8..15; then [16 steps][2] of the issuer.
    // Stamps stay in registers until the end of the step (a global store in front of a fence would be waited for by that fence).
    uint32_t pst[8];
#define BPROBE(slot) do { if (bprof) pst[slot] = (uint32_t)clock(); } while (0)
    constexpr int h = TH, H4 = 4 * TH;
    const int nb = a.nb, b0 = a.b0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                                   // A[m = unit][k = gate row p] (M-major, SWIZZLE_128B_BASE32B)
    uint8_t* sG = sm + BW_A_BYTES;                      // B[n = batch][k = gate row p] (K-major, SWIZZLE_128B)
    float* red = reinterpret_cast<float*>(sm + BW_A_BYTES + BW_G_BYTES);     // [2][src][unit][batch]
    uint8_t* sStg = sm + BW_A_BYTES + BW_G_BYTES + 2 * BW_R_BYTES;
    float4* sPa = reinterpret_cast<float4*>(sStg + BW_STG_BYTES);            // [2][pairs] gate activations of the step
    float* sPc = reinterpret_cast<float*>(sPa + 2 * BW_PAIRS);               // [2][pairs] c_{t-1}
    float* sPd = sPc + 2 * BW_PAIRS;                                         // [2][pairs] dout_t
    uint64_t* mbar_r = reinterpret_cast<uint64_t*>(sm + BW_A_BYTES + BW_G_BYTES + 2 * BW_R_BYTES + BW_STG_BYTES + BW_PRE_BYTES);   // [2]
    uint64_t* mbar_mma = mbar_r + 2;
    uint64_t* mbar_g = mbar_r + 3;                      // dG operand written by all 256 epilogue threads
    uint64_t* mbar_full = mbar_r + 4;                   // [2] prefetch buffer filled (loader warp)
    uint64_t* mbar_empty = mbar_r + 6;                  // [2] prefetch buffer consumed (one arrival per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_r + 8);
    uint32_t* chunk_cnt = tmem_slot + 1;               // epilogue warps that finished the current chunk's stores (cumulative)

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    // W slice rows p = 4*unit + gate (natural order, = dG's k index), all 256 unit columns n:
    // chunk j = n/32 at j*16 KB, k-row p at p*128 B, 32-byte sub-chunk ((n%32)/8) XOR (p % 4)
    if constexpr (BW_F16) {
        // item = (quad of gate rows p4, unit column n): four coalesced loads, one 8-byte store of element (row n, k = 4 p4 ..)
        for (int base = 0; base < 32 * h; base += 4 * BW_THREADS) {
            float v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * BW_THREADS + tid;
                if (idx < 32 * h) {
                    const int p4 = idx / h, n = idx % h;
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[u][q] = __ldg(a.Wl + (size_t)(128 * rank + 4 * p4 + q) * h + n);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * BW_THREADS + tid;
                if (idx < 32 * h) {
                    const int p4 = idx / h, n = idx % h;
                    *reinterpret_cast<uint2*>(sA + kmajor_off_h(n, 4 * p4, h)) = make_uint2(pack_h2(v[u][0], v[u][1]), pack_h2(v[u][2], v[u][3]));
                }
            }
        }
    } else {
    for (int base = 0; base < 128 * (h / 4); base += 8 * BW_THREADS) {      // batched loads, see the forward kernel
            float4 v[8];
    #pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * BW_THREADS + tid;
                if (idx < 128 * (h / 4)) {
                    const int p = idx / (h / 4), n = 4 * (idx % (h / 4));
                    v[u] = __ldg(reinterpret_cast<const float4*>(a.Wl + (size_t)(128 * rank + p) * h + n));
                }
            }
    #pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * BW_THREADS + tid;
                if (idx < 128 * (h / 4)) {
                    const int p = idx / (h / 4), n = 4 * (idx % (h / 4));
                    const uint32_t off = (uint32_t)((n >> 5) * 16384 + p * 128 + (((((n & 31) >> 3) ^ (p & 3))) << 5) + ((n & 7) << 2));
                    *reinterpret_cast<float4*>(sA + off) = make_float4(rnd_tf32(v[u].x), rnd_tf32(v[u].y), rnd_tf32(v[u].z), rnd_tf32(v[u].w));
                }
            }
        }
    }
    for (int idx = tid; idx < (int)(BW_G_BYTES / 16); idx += BW_THREADS) reinterpret_cast<float4*>(sG)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int idx = tid; idx < (int)(2 * BW_R_BYTES / 16); idx += BW_THREADS) reinterpret_cast<float4*>(red)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        mbar_init(&mbar_r[0], 1); mbar_init(&mbar_r[1], 1); mbar_init(mbar_mma, 1); mbar_init(mbar_g, TC_EPI);
        *chunk_cnt = 0u; chunk_cnt[1] = 0u; chunk_cnt[2] = 0u;      // [1], [2]: the per-step |dG| maximum (FP16 operand scale)
        mbar_init(&mbar_full[0], 32); mbar_init(&mbar_full[1], 32); mbar_init(&mbar_empty[0], TC_EPI / 32); mbar_init(&mbar_empty[1], TC_EPI / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(tmem_slot)), "n"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sA_addr = saddr(sA), sG_addr = saddr(sG), red_addr = saddr(red), sStg_addr = saddr(sStg);
    const uint32_t sPa_addr = saddr(sPa), sPc_addr = saddr(sPc), sPd_addr = saddr(sPd), cnt_addr = saddr(chunk_cnt), amax_addr = cnt_addr + 4;
    cluster_sync_all();

    if (w == 8) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = BW_F16 ? umma_idesc_f16(128, TROWS, false, false) : umma_idesc_tf32(128, TROWS, true, false);
            const uint64_t a0 = BW_F16 ? umma_smem_desc(sA_addr, 16, 512, 4) : umma_smem_desc(sA_addr, 16384, 512, 1);
            const uint64_t g0 = BW_F16 ? umma_smem_desc(sG_addr, 16, 512, 4) : umma_smem_desc(sG_addr, 16, 1024, 2);
            int step = 0;
            const int i_last = a.dh0 ? 0 : 1;       // dh0 requested: step 0 also sends (gradient w.r.t. the initial h)
            for (int i = T - 1; i >= i_last; --i, ++step) {
                mbar_expect_tx(&mbar_r[i & 1], BW_R_BYTES);       // arm the reduce buffer this step's sends fill
                mbar_wait(mbar_g, step & 1);                      // dG_i operand complete in smem
                BPROBE(0);
                tc_fence_after();
                // dh^T (256 x 16) = W_slice^T (256 x 128) · dG^T (128 x 16): two M = 128 halves, 16 k-steps each
#pragma unroll
                for (int hm = 0; hm < 2; ++hm) {
                    if constexpr (BW_F16) {
                        // K-major W^T: k-block kb at kb * 256 rows * 64 B, unit half hm at + 128 rows * 64 B; 8 k-steps of 16
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_f16_ss(tmem_base + hm * TROWS, a0 + (uint64_t)(((ks >> 1) * (256 * 64) + hm * (128 * 64) + (ks & 1) * 32) >> 4),
                                        g0 + (uint64_t)(((ks >> 1) * (TROWS * 64) + (ks & 1) * 32) >> 4), idesc, ks ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks)
                            umma_tf32_ss(tmem_base + hm * TROWS, a0 + (uint64_t)((hm * 4 * 16384 + ks * 1024) >> 4),
                                         g0 + (uint64_t)(((ks >> 2) * (TROWS * 128) + (ks & 3) * 32) >> 4), idesc, ks ? 1u : 0u);
                    }
                }
                umma_commit_arrive(mbar_mma);
                BPROBE(1);
                if (bprof && step >= 8 && step < 24) { bprof[256 + (step - 8) * 2] = pst[0]; bprof[256 + (step - 8) * 2 + 1] = pst[1]; }
            }
        }
    } else if (w == 10) {
        if (lane == 0 && (gt.done || (gt.ts && blockIdx.x == 0)))
            signaller_loop(cnt_addr, (T + gt.chunk - 1) / gt.chunk, gt, gt.ts && blockIdx.x == 0);
    } else if (w == 9) {
        // ===== loader warp: stages (gate activations, c_{t-1}, dout_t) of a step in shared memory one to two steps ahead (cp.async).
        // In the epilogue warps these global loads were still in flight at the proxy fence in front of the bulk send
        // (MEMBAR.ALL.CTA waits for every outstanding access of the thread: +650 cycles per step, tools/enc_step_probe.py). =====
        const int ju = TU * rank + lane;
        int tile_next = (T * B - 1) >> 7;          // a.tile_ready gating: tiles above this one are known complete (dout arrives last tile first)
        int step = 0;
        for (int i = T - 1; i >= 0; --i, ++step) {
            const int pb = step & 1;
            if (step >= 2) mbar_wait(&mbar_empty[pb], ((step >> 1) - 1) & 1);      // epilogue done with this buffer (step - 2)
            if (a.tile_ready) {
                const int need = (i * B + b0) >> 7;
                while (tile_next >= need) { spin_until_ge(a.tile_ready + tile_next, a.tile_target); --tile_next; }
            }
#pragma unroll
            for (int m = 0; m < TROWS; ++m)
                if (m < nb) {
                    const size_t r = (size_t)i * B + b0 + m;
                    const int idx = lane * TROWS + m;
                    cp_async16(sPa_addr + (uint32_t)((pb * BW_PAIRS + pre_slot4(idx)) * 16), a.G + r * H4 + 4 * ju);
                    cp_async4(sPc_addr + (uint32_t)((pb * BW_PAIRS + pre_slot1(idx)) * 4), a.Cs + r * h + ju);
                    // (each 128-byte line of dout is read exactly once, after the acquire on its tile flag: L1 cannot hold it stale)
                    cp_async4(sPd_addr + (uint32_t)((pb * BW_PAIRS + pre_slot1(idx)) * 4),
                              a.dout + (long long)i * a.out_si + (long long)(b0 + m) * a.out_sb + ju);
                }
            cp_async_arrive(&mbar_full[pb]);       // this lane's arrival fires when its copies above have landed
        }
    } else {
        // ===== epilogue / elementwise: pair e -> batch row m = idx % 16, unit ul = idx / 16 =====
        float dc[2];
        float4 p_act[2]; float p_c[2], p_cp[2], p_dout[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid + e * TC_EPI;
            const int m = idx & 15, ul = idx >> 4, ju = TU * rank + ul;
            const bool v = m < nb;
            dc[e] = (v && a.dc_fin) ? a.dc_fin[(size_t)(b0 + m) * a.ld_dc_fin + ju] : 0.f;
            p_c[e] = v ? a.Cs[((size_t)T * B + b0 + m) * h + ju] : 0.f;        // c_{T-1}; afterwards the previous step's c_{t-1}
            p_act[e] = make_float4(0.f, 0.f, 0.f, 0.f); p_cp[e] = p_dout[e] = 0.f;
        }
        int step = 0;
        for (int i = T - 1; i >= 0; --i, ++step) {
            const int buf = i & 1;
            const bool send = i > 0 || a.dh0 != nullptr;
            const uint32_t rprev_addr = red_addr + (uint32_t)(buf ^ 1) * BW_R_BYTES;
            {   // this step's prefetched operands (loader warp), read before the wait on the critical-path barrier
                const int pb = step & 1;
                mbar_wait(&mbar_full[pb], (step >> 1) & 1);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int idx = tid + e * TC_EPI;
                    if ((idx & 15) < nb) {
                        p_act[e] = lds_f4(sPa_addr + (uint32_t)((pb * BW_PAIRS + pre_slot4(idx)) * 16));
                        p_cp[e] = lds_f1(sPc_addr + (uint32_t)((pb * BW_PAIRS + pre_slot1(idx)) * 4));
                        p_dout[e] = lds_f1(sPd_addr + (uint32_t)((pb * BW_PAIRS + pre_slot1(idx)) * 4));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&mbar_empty[pb]);
            }
            if (i < T - 1) mbar_wait(&mbar_r[buf ^ 1], ((T - 2 - i) >> 1) & 1);     // partial dh of step i+1 from all CTAs
            BPROBE(0);
            // 1. dG_t for the owned units (K-major UMMA operand, TF32-rounded)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int idx = tid + e * TC_EPI;
                const int m = idx & 15, ul = idx >> 4, ju = TU * rank + ul;
                float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < nb) {
                    float dh;
                    if (i == T - 1) {
                        dh = a.dh_fin ? a.dh_fin[(size_t)(b0 + m) * a.ld_dh_fin + ju] : 0.f;
                    } else {
                        dh = 0.f;
#pragma unroll
                        for (int s = 0; s < TNC; ++s) dh += lds_f1(rprev_addr + (uint32_t)(((s * TU + ul) * TROWS + m) * 4));
                    }
                    const size_t r = (size_t)i * B + b0 + m;
                    const float dm = dropout_scale(seed, a.drop_stream, (uint32_t)(r * h + ju) + a.drop_off, drop);
                    dh += p_dout[e] * dm;
                    const float4 act = p_act[e];
                    const float c = p_c[e], cp = p_cp[e];
                    const float tc = fast_tanh(c);
                    const float dct = dc[e] + dh * act.w * (1.f - tc * tc);
                    dg.x = dct * act.y * (1.f - act.x * act.x);
                    dg.y = dct * act.x * act.y * (1.f - act.y);
                    dg.z = dct * cp * act.z * (1.f - act.z);
                    dg.w = dh * tc * act.w * (1.f - act.w);
                    dc[e] = dct * act.z;
                    p_act[e] = dg;
                    p_c[e] = cp;          // rotate here, where c_{t-1} is already in hand: a move after the prefetch load would wait for it
                }
                if constexpr (!BW_F16) {
                    if (send)
                        sts_f4(sG_addr + kmajor_off(m, 4 * ul, TROWS), make_float4(rnd_tf32(dg.x), rnd_tf32(dg.y), rnd_tf32(dg.z), rnd_tf32(dg.w)));
                }
            }
            float unscale = 1.f;
            if constexpr (BW_F16) {
                if (send) {
                    // exact |dG| maximum of the CTA's operand -> power-of-two scale into fp16's comfortable range
                    uint32_t mx = 0u;
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        mx = max(max(mx, __float_as_uint(fabsf(p_act[e].x))), max(max(__float_as_uint(fabsf(p_act[e].y)), __float_as_uint(fabsf(p_act[e].z))),
                                                                                  __float_as_uint(fabsf(p_act[e].w))));
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    const uint32_t slot = amax_addr + (uint32_t)((step & 1) * 4);
                    if (lane == 0) asm volatile("red.shared::cta.max.u32 [%0], %1;" ::"r"(slot), "r"(mx) : "memory");
                    epi_barrier();
                    uint32_t am;
                    asm volatile("ld.shared::cta.u32 %0, [%1];" : "=r"(am) : "r"(slot) : "memory");
                    if (tid == 0) asm volatile("st.shared::cta.u32 [%0], %1;" ::"r"(amax_addr + (uint32_t)(((step & 1) ^ 1) * 4)), "r"(0u) : "memory");
                    float scale = 1.f;
                    if (am != 0u) {
                        const int E = min(max((int)((am >> 23) & 0xFFu), 9), 253);
                        scale = __uint_as_float((uint32_t)(262 - E) << 23);        // max * scale in [256, 512)
                        unscale = __uint_as_float((uint32_t)(E - 8) << 23);        // exactly 1 / scale
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int idx = tid + e * TC_EPI;
                        const int m = idx & 15, ul = idx >> 4;
                        const float4 dg = p_act[e];
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sG_addr + kmajor_off_h(m, 4 * ul, TROWS)),
                                     "r"(pack_h2(dg.x * scale, dg.y * scale)), "r"(pack_h2(dg.z * scale, dg.w * scale)) : "memory");
                    }
                }
            }
            BPROBE(1);
            if (send) {
                fence_proxy_async();
                mbar_arrive(mbar_g);              // hand the operand to the issuer warp
            }
            BPROBE(2);
#define BW_STORE_DG() do { _Pragma("unroll") for (int e = 0; e < 2; ++e) { \
                const int idx = tid + e * TC_EPI; const int m = idx & 15, ul = idx >> 4, ju = TU * rank + ul; \
                if (m < nb) *reinterpret_cast<float4*>(a.G + ((size_t)i * B + b0 + m) * H4 + 4 * ju) = p_act[e]; } \
            if ((T - i) % gt.chunk == 0 || i == 0) chunk_arrive(cnt_addr, lane); } while (0)
            BW_STORE_DG();                         // 2. bookkeeping while the tensor core works: write dG_t in place
            BPROBE(3);
            if (send) {
                // 3. reduce-scatter: TMEM lane = unit n (half hm = w/4, quadrant w%4) -> owner CTA n/32, 16 batch partials
                mbar_wait(mbar_mma, step & 1);
                BPROBE(4);
                tc_fence_after();
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + (w >> 2) * TROWS, v);
                tc_fence_before();
                if constexpr (BW_F16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] *= unscale;
                }
                BPROBE(5);
                // warp w holds exactly the block owner CTA w needs ([32 units][16 batch] partials = 2 KB contiguous at the
                // receiver): stage it and push it with ONE bulk copy per warp instead of 4 st.async per lane
                const int owner = (w >> 2) * 4 + (w & 3);
                const uint32_t stg = sStg_addr + (uint32_t)((buf * TNC + owner) * (TU * TROWS * 4));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)(lane * TROWS * 4 + 16 * j)),
                                 "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
                fence_proxy_async();
                __syncwarp();
                BPROBE(6);
                if (lane == 0) {
                    const uint32_t dst = mapa(red_addr + (uint32_t)(((buf * TNC + rank) * TU) * TROWS * 4), owner);
                    const uint32_t bar = mapa(saddr(&mbar_r[buf]), owner);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "r"(stg), "r"((uint32_t)(TU * TROWS * 4)), "r"(bar) : "memory");
                }
                BPROBE(7);
            }
            if (bprof && step >= 8 && step < 24 && (tid == 0 || tid == 224)) {
#pragma unroll
                for (int q = 0; q < 8; ++q) bprof[(step - 8) * 16 + (tid ? 8 : 0) + q] = pst[q];
            }
        }
        // gradients w.r.t. the initial state (slot 0): the carry into the previous chunk of a longer sequence
        if (a.dh0 || a.dc0) {
            if (a.dh0) mbar_wait(&mbar_r[0], ((T - 1) >> 1) & 1);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int idx = tid + e * TC_EPI;
                const int m = idx & 15, ul = idx >> 4, ju = TU * rank + ul;
                if (m < nb) {
                    if (a.dh0) {
                        float s = 0.f;
#pragma unroll
                        for (int sidx = 0; sidx < TNC; ++sidx) s += lds_f1(red_addr + (uint32_t)(((sidx * TU + ul) * TROWS + m) * 4));
                        a.dh0[(size_t)(b0 + m) * h + ju] = s;
                    }
                    if (a.dc0) a.dc0[(size_t)(b0 + m) * h + ju] = dc[e];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (w == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32));
    cluster_sync_all();
}

template <class KernT>
static int launch_tc(KernT kern, cudaStream_t st, int nchains, size_t smem, const LstmChains& ch, int T, int B,
                     float drop, unsigned long long seed, LstmGate gate = LstmGate{nullptr, 1, nullptr}) {
    gate.probe = g_lstm_prof;
    AST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nchains * TNC);
    cfg.blockDim = dim3(BW_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TNC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AST_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ch, T, B, drop, seed, gate));
    ++g_kernel_launches;
    return 0;
}
static int launch_tc_fwd(cudaStream_t st, int nchains, size_t smem, const LstmChains& ch, int T, int B, float drop, unsigned long long seed,
                         LstmGate gate = LstmGate{nullptr, 1, nullptr}) {
    AST_CUDA_OK(cudaFuncSetAttribute(lstm_seq_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nchains * TNC);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TNC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AST_CUDA_OK(cudaLaunchKernelEx(&cfg, lstm_seq_fwd_tc_kernel, ch, T, B, drop, seed, g_lstm_prof, gate));
    ++g_kernel_launches;
    return 0;
}

// How many 8-CTA clusters of the recurrence kernels the device can hold at once (cudaOccupancyMaxActiveClusters): the
// persistent encoder wavefront needs every cluster of all layers co-resident, its caller checks that before choosing it.
int lstm_seq_tc_max_clusters(bool backward) {
    static int cached[2] = {-1, -1};
    int& c = cached[backward ? 1 : 0];
    if (c >= 0) return c;
    const size_t smem = backward ? BW_SMEM : FW_SMEM;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(TNC * 64);
    cfg.blockDim = dim3(backward ? BW_THREADS : TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TNC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e;
    if (backward) {
        e = cudaFuncSetAttribute(lstm_seq_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, lstm_seq_bwd_tc_kernel, &cfg);
    } else {
        e = cudaFuncSetAttribute(lstm_seq_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, lstm_seq_fwd_tc_kernel, &cfg);
    }
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    c = n;
    return c;
}
int lstm_seq_tc_cluster_size() { return TNC; }

// `ch` already holds 16-row chains (lstm_seq.cu::expand_chains).  Preconditions checked by the caller: h == 256.
int lstm_seq_fwd_tc(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed) {
    return launch_tc_fwd(st, nchains, FW_SMEM, ch, T, B, drop, seed);
}
int lstm_seq_bwd_tc(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed) {
    return launch_tc(lstm_seq_bwd_tc_kernel, st, nchains, BW_SMEM, ch, T, B, drop, seed);
}
int lstm_seq_fwd_tc_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed, const LstmGate& gate) {
    AST_CHECK(gate.chunk >= 1, "lstm_seq_fwd_tc_gated: chunk must be >= 1");
    return launch_tc_fwd(st, nchains, FW_SMEM, ch, T, B, drop, seed, gate);
}
int lstm_seq_bwd_tc_gated(cudaStream_t st, const LstmChains& ch, int nchains, int T, int B, float drop, unsigned long long seed, const LstmGate& gate) {
    AST_CHECK(gate.chunk >= 1, "lstm_seq_bwd_tc_gated: chunk must be >= 1");
    return launch_tc(lstm_seq_bwd_tc_kernel, st, nchains, BW_SMEM, ch, T, B, drop, seed, gate);
}

}  // namespace ast

// Persistent decoder-sequence kernels: the whole teacher-forced / scheduled-sampling decoder pass of
// forward_loss (seq2seq.py:423-470: embed -> 3 LSTM cells -> attention -> context/tanh -> out -> softmax-CE ->
// argmax, L-1 times) and its backward run as ONE cooperative launch each.  One CTA per SM stays resident for
// all steps; the phases of a step are separated by grid barriers instead of kernel boundaries, so a decoder
// step costs ~9 barrier latencies instead of ~10 (forward) / ~12 (backward) kernel launches + drains.
// Scheduled sampling stays on the device: the CE phase computes the argmax and writes the next step's
// embedding row itself.  Arithmetic is shared with the per-step kernels (decoder_dev.cuh).
#include <cooperative_groups.h>
#include "decoder_dev.cuh"

namespace cg = cooperative_groups;

namespace ast {

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
// grid barrier + optional timestamp (phase-timing probe, tools/dec_phase_times.py)
// Grid barrier on one monotonically increasing global counter (zeroed by the host before the launch; the cooperative
// launch guarantees co-residency).  One release-add per CTA, one acquire-poll loop in thread 0: measured ~X us against
// cooperative_groups' grid.sync() (tools/dec_phase_times.py prints both).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
}
#define GRID_SYNC() do { if (p.bar) grid_barrier(p.bar, bar_target); else grid.sync(); \
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) { p.prof[++nprof] = gtimer(); p.prof[0] = (unsigned long long)nprof; } } while (0)

// x0[s][b][0:E] = Emb[word] * dropmask ; words_used[s][b] = word ; (s == 0) x0[0][b][E:] = 0
__device__ __forceinline__ void embed_row(const DecSeq& p, int s, int b, int word) {
    word = min(max(word, 0), p.V - 1);
    float* dst = p.x0 + ((size_t)s * p.B + b) * (p.E + p.A);
    if (threadIdx.x == 0) p.words_used[(size_t)s * p.B + b] = word;
    for (int j = threadIdx.x; j < p.E; j += blockDim.x) {
        const float dm = dropout_scale(p.seed, 32, (uint32_t)(((size_t)s * p.B + b) * p.E + j), p.drop_embed);
        dst[j] = __ldg(p.emb + (size_t)word * p.E + j) * dm;
    }
    if (s == 0)
        for (int j = threadIdx.x; j < p.A; j += blockDim.x) dst[p.E + j] = 0.f;
}

template <int MT, bool EXACT>
__global__ void __launch_bounds__(SK_THREADS, 1)
dec_seq_fwd_kernel(DecSeq p) {
    cg::grid_group grid = cg::this_grid();
    int nprof = 0;
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[1] = gtimer();
    nprof = 1;
    unsigned bar_target = 0;
    if (p.prof) {      // barrier micro-benchmark: 4 x cooperative_groups, then 4 x the counter barrier
        for (int i = 0; i < 4; ++i) { grid.sync(); if (blockIdx.x == 0 && threadIdx.x == 0) p.prof[++nprof] = gtimer(); }
        if (p.bar) for (int i = 0; i < 4; ++i) { grid_barrier(p.bar, bar_target); if (blockIdx.x == 0 && threadIdx.x == 0) p.prof[++nprof] = gtimer(); }
    }
    __shared__ SkinnySmem sm;
    __shared__ float scratch[32];
    __shared__ int iscratch[32];
    extern __shared__ float dsm[];
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int B = p.B, H = p.H, E = p.E, A = p.A, Tp = p.Tp, NL = p.NL, S = p.S, L = p.L;
    const int ldx0 = E + A;

    for (int b = cta; b < B; b += ncta) embed_row(p, 0, b, p.y[(size_t)b * L]);
    GRID_SYNC();

    for (int s = 0; s < S; ++s) {
        // ---- LSTM stack (seq2seq.py:375) ----------------------------------------------------------------
        for (int l = 0; l < NL; ++l) {
            SkinnyArgs a{};
            const int in = l == 0 ? E + A : H;
            if (l == 0) { a.X[0] = p.x0 + (size_t)s * B * ldx0; a.ldx[0] = ldx0; }
            else { a.X[0] = p.hdd[l - 1] + (size_t)s * B * H; a.ldx[0] = H; }
            a.K[0] = in; a.W[0] = p.Wup[l]; a.ldw[0] = in;
            a.X[1] = p.Hd[l] + (size_t)s * B * H; a.ldx[1] = H; a.K[1] = H; a.W[1] = p.Wlat[l]; a.ldw[1] = H;
            a.bias = p.bup[l]; a.B = B; a.N = 4 * H; a.epi = EPI_LSTM;
            a.Y = p.act[l] + (size_t)s * B * 4 * H; a.ldy = 4 * H;
            a.c_prev = p.Cd[l] + (size_t)s * B * H; a.c_out = p.Cd[l] + (size_t)(s + 1) * B * H;
            a.h_out = p.Hd[l] + (size_t)(s + 1) * B * H;
            if (l == NL - 1) { a.hd_out = p.cvh + (size_t)s * B * 2 * H + H; a.ld_hd = 2 * H; }
            else { a.hd_out = p.hdd[l] + (size_t)s * B * H; a.ld_hd = H; }
            a.drop = p.drop_rnn; a.seed = p.seed; a.drop_stream = 16 + l; a.drop_base = (size_t)s * B * H;
            for (int g = cta; g < (4 * H) / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        const float* htop = p.cvh + (size_t)s * B * 2 * H + H;
        float* q = p.q + (size_t)s * B * H;
        {   // q = attn_Wa(h)  (:341)
            SkinnyArgs a{};
            a.X[0] = htop; a.ldx[0] = 2 * H; a.K[0] = H; a.W[0] = p.Wa; a.ldw[0] = H; a.bias = p.ba;
            a.B = B; a.N = H; a.epi = EPI_NONE; a.Y = q; a.ldy = H;
            for (int g = cta; g < H / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        const long long ebs = (long long)Tp * H;
        {   // scores (:342)
            const int ntb = (Tp + 7) / 8;
            for (int i = cta; i < ntb * B; i += ncta) attn_dot_block(p.enc, ebs, q, H, p.scores, Tp, H, i / ntb, i % ntb);
            GRID_SYNC();
        }
        {   // softmax over T' + context (:351-355)
            const int njb = (H + 127) / 128;
            for (int i = cta; i < njb * B; i += ncta)
                attn_ctx_block(p.enc, ebs, p.scores, p.alpha + (size_t)s * B * Tp, p.cvh + (size_t)s * B * 2 * H, 2 * H, Tp, H,
                               i / njb, i % njb, dsm, dsm + ((Tp + 3) & ~3), scratch);
            GRID_SYNC();
        }
        {   // ht = tanh(context([cv;h]))  (:386-390); also the next step's input-feeding slot
            SkinnyArgs a{};
            a.X[0] = p.cvh + (size_t)s * B * 2 * H; a.ldx[0] = 2 * H; a.K[0] = 2 * H; a.W[0] = p.Wc; a.ldw[0] = 2 * H; a.bias = p.bc;
            a.B = B; a.N = A; a.epi = EPI_TANH; a.Y = p.ht + (size_t)s * B * A; a.ldy = A;
            if (s + 1 < S) { a.Y2 = p.x0 + (size_t)(s + 1) * B * ldx0 + E; a.ldy2 = ldx0; }
            for (int g = cta; g < A / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        float* z = p.logits + (size_t)s * B * p.Vp;
        {   // logits = out(ht)  (:394)
            SkinnyArgs a{};
            a.X[0] = p.ht + (size_t)s * B * A; a.ldx[0] = A; a.K[0] = A; a.W[0] = p.Wo; a.ldw[0] = A; a.bias = p.bo;
            a.B = B; a.N = p.V; a.epi = EPI_NONE; a.Y = z; a.ldy = p.Vp;
            for (int g = cta; g < (p.V + SK_COLS - 1) / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        // softmax-CE (+ gradient in place) + argmax (:448,468) + next decoder input (:431-436)
        for (int b = cta; b < B; b += ncta) {
            const int target = p.y[(size_t)b * L + s + 1];
            const int mi = softmax_ce_row(z + (size_t)b * p.Vp, p.Vp, p.V, target, B, p.row_loss + (size_t)s * B + b, 1, scratch, iscratch);
            if (threadIdx.x == 0) p.argmax_steps[(size_t)s * B + b] = mi;
            if (s + 1 < S) {
                const bool ut = (p.use_true == nullptr) || p.use_true[s + 1];
                embed_row(p, s + 1, b, ut ? target : mi);
            }
            __syncthreads();
        }
        GRID_SYNC();
    }
}

template <int MT, bool EXACT>
__global__ void __launch_bounds__(SK_THREADS, 1)
dec_seq_bwd_kernel(DecSeq p) {
    cg::grid_group grid = cg::this_grid();
    int nprof = 0;
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[1] = gtimer();
    nprof = 1;
    unsigned bar_target = 0;
    if (p.prof) {      // barrier micro-benchmark: 4 x cooperative_groups, then 4 x the counter barrier
        for (int i = 0; i < 4; ++i) { grid.sync(); if (blockIdx.x == 0 && threadIdx.x == 0) p.prof[++nprof] = gtimer(); }
        if (p.bar) for (int i = 0; i < 4; ++i) { grid_barrier(p.bar, bar_target); if (blockIdx.x == 0 && threadIdx.x == 0) p.prof[++nprof] = gtimer(); }
    }
    __shared__ SkinnySmem sm;
    __shared__ float scratch[32];
    extern __shared__ float dsm[];
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int B = p.B, H = p.H, E = p.E, A = p.A, Tp = p.Tp, NL = p.NL, S = p.S, Vp = p.Vp;
    const long long ebs = (long long)Tp * H;
    const int ld0 = E + A + H;

    for (int s = S - 1; s >= 0; --s) {
        float* du = p.du + (size_t)s * B * A;
        {   // du = (dz . Wo + dht_feed) * (1 - ht^2)
            SkinnyArgs a{};
            a.X[0] = p.logits + (size_t)s * B * Vp; a.ldx[0] = Vp; a.K[0] = Vp; a.W[0] = p.WoT; a.ldw[0] = Vp;
            a.B = B; a.N = A; a.epi = EPI_TANHBWD; a.Y = du; a.ldy = A;
            if (s < S - 1) { a.add = p.dxh[0] + E; a.ld_add = ld0; }
            a.aux = p.ht + (size_t)s * B * A; a.ld_aux = A;
            for (int g = cta; g < A / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        {   // dcvh = du . Wc
            SkinnyArgs a{};
            a.X[0] = du; a.ldx[0] = A; a.K[0] = A; a.W[0] = p.WcT; a.ldw[0] = A;
            a.B = B; a.N = 2 * H; a.epi = EPI_NONE; a.Y = p.dcvh; a.ldy = 2 * H;
            for (int g = cta; g < (2 * H) / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        {   // dalpha[b][t] = enc[b][t][:] . dcv[b][:]
            const int ntb = (Tp + 7) / 8;
            for (int i = cta; i < ntb * B; i += ncta) attn_dot_block(p.enc, ebs, p.dcvh, 2 * H, p.dalpha, Tp, H, i / ntb, i % ntb);
            GRID_SYNC();
        }
        float* dq = p.dq + (size_t)s * B * H;
        {   // softmax / scores / context backward, d_enc accumulation
            const int njb = (H + 127) / 128;
            for (int i = cta; i < njb * B; i += ncta)
                attn_bwd_block(p.enc, p.d_enc, ebs, p.alpha + (size_t)s * B * Tp, p.dalpha, p.dcvh, 2 * H, p.q + (size_t)s * B * H, H,
                               dq, H, Tp, H, i / njb, i % njb, dsm, dsm + ((2 * Tp + 3) & ~3), scratch);
            GRID_SYNC();
        }
        {   // dh_top = dcvh[:, H:] + dq . Wa ; fused: cell backward of the top LSTM layer
            const int l = NL - 1, in = l == 0 ? E + A : H;
            SkinnyArgs a{};
            a.X[0] = dq; a.ldx[0] = H; a.K[0] = H; a.W[0] = p.WaT; a.ldw[0] = H;
            a.B = B; a.N = H; a.epi = EPI_CELLBWD; a.add = p.dcvh + H; a.ld_add = 2 * H;
            a.cb_act = p.act[l] + (size_t)s * B * 4 * H; a.cb_c = p.Cd[l] + (size_t)(s + 1) * B * H; a.cb_c_prev = p.Cd[l] + (size_t)s * B * H;
            a.cb_dc = p.dcd[l]; a.cb_H = H; a.cb_ncols = H;
            if (s < S - 1) { a.cb_dh_rec = p.dxh[l] + in; a.cb_ld_dh_rec = in + H; }
            a.drop = p.drop_rnn; a.seed = p.seed; a.drop_stream = 16 + l; a.drop_base = (size_t)s * B * H;
            for (int g = cta; g < H / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
        for (int l = NL - 1; l >= 0; --l) {
            // [dx | dh_rec] = dG_l . [W_up | W_lat] ; fused: cell backward of layer l-1 on the dx columns,
            // or (l == 0) the EmbedID scatter-add on the embedding columns
            const int in = l == 0 ? E + A : H;
            SkinnyArgs a{};
            a.X[0] = p.act[l] + (size_t)s * B * 4 * H; a.ldx[0] = 4 * H; a.K[0] = 4 * H; a.W[0] = p.WcatT[l]; a.ldw[0] = 4 * H;
            a.B = B; a.N = in + H; a.Y = p.dxh[l]; a.ldy = in + H; a.epi = EPI_NONE;
            if (l > 0) {
                const int lb = l - 1, inb = lb == 0 ? E + A : H;
                a.epi = EPI_CELLBWD;
                a.cb_act = p.act[lb] + (size_t)s * B * 4 * H; a.cb_c = p.Cd[lb] + (size_t)(s + 1) * B * H;
                a.cb_c_prev = p.Cd[lb] + (size_t)s * B * H; a.cb_dc = p.dcd[lb]; a.cb_H = H; a.cb_ncols = H;
                if (s < S - 1) { a.cb_dh_rec = p.dxh[lb] + inb; a.cb_ld_dh_rec = inb + H; }
                a.drop = p.drop_rnn; a.seed = p.seed; a.drop_stream = 16 + lb; a.drop_base = (size_t)s * B * H;
            } else {
                a.sc_demb = p.demb; a.sc_words = p.words_used + (size_t)s * B; a.sc_E = E;
                a.drop = p.drop_embed; a.seed = p.seed; a.drop_stream = 32; a.drop_base = (size_t)s * B * E;
            }
            for (int g = cta; g < (in + H) / SK_COLS; g += ncta) skinny_tile<MT, EXACT>(a, g * SK_COLS, sm);
            GRID_SYNC();
        }
    }
}

template <class KernT>
static int launch_coop(KernT kern, cudaStream_t st, const DecSeq& p, size_t dsm_bytes) {
    int dev = 0, sms = 0, occ = 0;
    AST_CUDA_OK(cudaGetDevice(&dev));
    AST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AST_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SK_THREADS, dsm_bytes));
    AST_CHECK(occ >= 1, "decoder sequence kernel does not fit on an SM");
    DecSeq pp = p;
    if (pp.bar) AST_CUDA_OK(cudaMemsetAsync(pp.bar, 0, sizeof(unsigned), st));
    void* args[] = {&pp};
    AST_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(sms), dim3(SK_THREADS), args, dsm_bytes, st));
    ++g_kernel_launches;
    return 0;
}

static int check(const DecSeq& p) {
    AST_CHECK(p.B >= 1 && p.B <= 32, "dec_seq: batch %d unsupported (1..32)", p.B);
    AST_CHECK(p.H % 16 == 0 && p.A % 16 == 0 && p.E % 16 == 0 && p.Vp % 16 == 0, "dec_seq: H, A, E, Vp must be multiples of 16");
    AST_CHECK(p.S >= 1, "dec_seq: need at least one decode step");
    return 0;
}

int dec_seq_fwd(cudaStream_t st, const DecSeq& p, bool exact) {
    AST_TRY(check(p));
    const size_t dsm = sizeof(float) * (p.Tp + 4 + 8 * 128);
    if (p.B <= 16) return exact ? launch_coop(dec_seq_fwd_kernel<1, true>, st, p, dsm) : launch_coop(dec_seq_fwd_kernel<1, false>, st, p, dsm);
    return exact ? launch_coop(dec_seq_fwd_kernel<2, true>, st, p, dsm) : launch_coop(dec_seq_fwd_kernel<2, false>, st, p, dsm);
}

int dec_seq_bwd(cudaStream_t st, const DecSeq& p, bool exact) {
    AST_TRY(check(p));
    const size_t dsm = sizeof(float) * (2 * p.Tp + 4 + 8 * 128);
    if (p.B <= 16) return exact ? launch_coop(dec_seq_bwd_kernel<1, true>, st, p, dsm) : launch_coop(dec_seq_bwd_kernel<1, false>, st, p, dsm);
    return exact ? launch_coop(dec_seq_bwd_kernel<2, true>, st, p, dsm) : launch_coop(dec_seq_bwd_kernel<2, false>, st, p, dsm);
}

}  // namespace ast

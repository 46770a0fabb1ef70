// Kernel (b): fused backward of GlobalAttentionGeneral on the 5th-generation tensor cores
// (SBA_ALGO_TCGEN05, bf16 tensors).  One persistent CTA (320 threads, 2 per SM) = 1 TMA producer
// warp + 1 MMA-issuing warp + 4 first-stage warps + 4 second-stage warps (both stages thread = pixel,
// each group covers the four TMEM lane quarters), contiguous ranges of 128-pixel tiles.  Per tile
// (formulas: SURVEY.md §8a-4, oracle/attention.py):
//   TMA         : the [idf x 128 px] tiles of g_c and x land interleaved box by box, so that the same
//                 bytes are (i) two MN-major A operands (pixels = M) and (ii) ONE K-major A operand
//                 [g ; x] with 2*idf rows (pixels = K).
//   MMA1        : S  = x^T . sourceT      dP = g^T . sourceT                     (K = idf)
//   first stage : P = masked softmax(S) (recomputed, not re-read from HBM)
//                 dS = P * (dP [+ g_attn] - sum_l P dP)
//                 P and dS are written (bf16, 128-byte swizzled rows [word][pixel]) into the
//                 shared-memory operand buffer PB[tile parity].
//   MMA2        : dX   = dS . sourceT^T          (A = the dS rows of PB, MN-major; K = words)
//                 dSrc += [g ; x] . [P | dS]     (A = the staged tiles, B = PB, K = 128 pixels;
//                                                 the accumulator stays in TMEM across all tiles of a
//                                                 sample: rows = g / x channels, columns = P / dS words;
//                                                 its diagonal blocks are g.P and x.dS)
//   second stage: dX row of the pixel -> staged -> TMA box store; at the end of a sample the two
//                 diagonal blocks of the TMEM accumulator are added to dSrc[b] with fp32 atomics.
// Two {S, dP} TMEM buffers and two PB buffers form a software pipeline: MMA1(j+1) is issued before
// MMA2(j), so the first stage of tile j+1 overlaps the tensor core and the second stage of tile j.
// At a sample boundary the pipeline drains (the operands are rebuilt only after the last MMA of the
// old sample has retired, the accumulator is handed back after it has been read).
// dSrc / dW are zeroed by a small kernel in front (programmatic dependent launch: this kernel only
// waits for it before its first atomic), and dW = sum_b dSrc[b] . ctx[b]^T, dCtx[b] = W^T . dSrc[b]
// are formed by a small kernel behind it (attn_bwd_post; also a programmatic dependent, resident
// early): doing that per sample inside this kernel stalls CTAs mid-stream and piles the whole
// reduction up at the end of the stream.
#include <cstdlib>

#include "kernels.h"
#include "tc5_common.cuh"

namespace sba {
namespace {
using namespace tc5;

// development aid (SBA_TC5_TIMELINE): globaltimer stamps of the kernels of consecutive backward calls
__device__ unsigned long long g_timeline[16 * 8];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_min(int slot) { if (slot >= 0) atomicMin(&g_timeline[slot], gtime()); }
__device__ __forceinline__ void tl_max(int slot) { if (slot >= 0) atomicMax(&g_timeline[slot], gtime()); }

// development aid: SBA_TC5_TIMELINE=1 stamps the kernels of calls 20..35 and prints them at call 36
static int g_tl_call = 0;
static int timeline_slot() {
    static const bool on = getenv("SBA_TC5_TIMELINE") != nullptr;
    if (!on) return -1;
    const int c = g_tl_call - 20;
    return (c >= 0 && c < 16) ? c * 8 : -1;
}


struct Tc5BwdParams {
    const float* srcT;
    const uint8_t* mask;
    const uint32_t* mask_bits;   // [B] caption mask words written by the forward (scratch)
    const void* ga;       // [B, L, Q] nullable
    float* dSrc;          // [B, idf, L]
    const float* ctx;     // [B, cdf, L]   (epilogue)
    const float* W;       // [idf, cdf]    (epilogue, dCtx only)
    float* dW;            // [idf, cdf]    nullable
    float* dCtx;          // [B, cdf, L]   nullable
    int B, L, Q, cdf, mask_mode;
    int tiles_per_sample;
    int n_tiles;
    long long* trace;     // development aid (SBA_TC5_TRACE): per-phase clock64 stamps of CTA 0, else NULL
    int tl;               // development aid (SBA_TC5_TIMELINE): slot base in g_timeline, else -1
};

// producer warp, MMA warp, 4 first-stage ("math") warps, 4 second-stage ("epilogue") warps
constexpr int kBwdThreads = 64 + 128 + 128;
constexpr int kFirstEpilogueWarp = 6;

constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int IDF, int NQ, int NST_ = 3>
struct Tc5BwdCfg {
    static constexpr int ES = 2;                                // bf16
    static constexpr int LP = 4 * NQ;                           // words held per thread
    static constexpr int RP = (LP + 7) / 8 * 8;                 // P rows of PB (whole 8-row swizzle atoms)
    static constexpr int K2 = (LP + 15) / 16 * 16;              // dS rows of PB = K extent of the dX MMA
    static constexpr int NR = RP + K2;                          // PB rows
    static constexpr int ND = 2 * RP;                           // N of the dSrc MMA: P words | dS words
    static constexpr int BOX_PX = 64, NBOX = TQ / BOX_PX;       // 128-byte box rows
    static constexpr int BOX_BYTES = IDF * 128;
    static constexpr int STAGE_BYTES = 2 * NBOX * BOX_BYTES;    // g and x tiles, boxes interleaved (g0 x0 g1 x1)
    static constexpr int KS1 = IDF / 16;                        // MMA1 k-steps (channels)
    static constexpr int KS2 = K2 / 16;                         // dX k-steps (words)
    static constexpr int KS3 = TQ / 16;                         // dSrc k-steps (pixels)
    static constexpr int NS = 32;                               // MMA1 N (words, zero padded)
    static constexpr int KCH1 = IDF * ES / 16, KCH2 = K2 * ES / 16;
    static constexpr int B1_BYTES = NS * IDF * ES, B2_BYTES = IDF * K2 * ES;
    static constexpr int PB_KBLOCK = NR * 128;                  // bytes of one 64-pixel block of PB
    static constexpr int PB_BYTES = NBOX * PB_KBLOCK;           // one PB buffer; two of them alternate (software pipeline)
    static constexpr int NST = NST_;                            // g/x ring depth: 3, or 4 for long streams at idf 32 (measured +2-4 %)
    static constexpr int MD = 2 * IDF <= 64 ? 64 : 128;         // M of the dSrc MMA
    // TMEM columns: two {S, dP} buffers (tile parity), dX, the dSrc accumulator
    static constexpr int COL_S = 0, COL_DP = 32, COL_BUF = 64, COL_DX = 128, COL_ACC = 128 + IDF;
    static constexpr int TMEM_COLS = pow2_cols(COL_ACC + ND);
    static constexpr int OUT_WARP_BYTES = IDF * 32 * ES;        // per-warp dX staging [channel][32 px]
    static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 2 * PB_BYTES + B1_BYTES + B2_BYTES + 4 * OUT_WARP_BYTES;
    static constexpr uint32_t IDESC1 = make_idesc(1, 1, 0, TQ, NS);      // A = tile, MN-major
    static constexpr uint32_t IDESC2 = make_idesc(1, 1, 0, TQ, IDF);     // A = dS rows of PB, MN-major
    static constexpr uint32_t IDESC3 = make_idesc(1, 0, 0, MD, ND);      // A = [g ; x], B = PB, both K-major
    static constexpr int CTAS_PER_SM = 512 / TMEM_COLS;
    static_assert(IDF % 16 == 0 && 2 * IDF <= 128, "idf must be a multiple of 16, at most 64");
    static_assert(LP <= 32 && IDF <= 64, "at most 32 words");
};

// zero dSrc, the per-sample counters and dW; the streaming kernel waits for this grid only before its
// first atomic (griddepcontrol.wait), so the fill overlaps its prologue and first tiles
__global__ void __launch_bounds__(256) k_zero_tc5(float* __restrict__ a, size_t na, float* __restrict__ b, size_t nb, int tl) {
    // programmatic dependent of whatever precedes it (hides its launch latency); upstream work is complete
    // before its own dependent - the streaming kernel, which reads x / g_c at once - may start
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) tl_min(tl);
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < na; i += step) a[i] = 0.f;
    if (b != nullptr)
        for (size_t i = i0; i < nb; i += step) b[i] = 0.f;
    if (threadIdx.x == 0) tl_max(tl < 0 ? tl : tl + 1);
}

// dW += sum_{b in group} dSrc[b] . ctx[b]^T for a [idf x 32] slice, dCtx[b] = W^T . dSrc[b]; a programmatic
// dependent of the streaming kernel (griddepcontrol.wait = that grid is complete and flushed).
//   blocks [0, n_dw)          : (32 input channels c, one of 64 sample groups).  The K = (sample, word) axis
//                               of up to 4 samples is staged flat - ds [idf][K], cs [K][32 c] - so the inner
//                               loop is branch-free: one broadcast LDS + one LDS.128 per 4 FMAs per thread
//                               (thread = channel i x 4 channels c).  The 64 groups meet in fp32 atomics on
//                               dW (dW zeroed by k_zero_tc5); many small blocks hide each other's latency.
//   blocks [n_dw, n_dw + B)   : dCtx of one sample (only when words need a gradient)
constexpr int kPostCS = 36;        // row stride (floats) of cs: 16-byte aligned rows, 4-bank skew
constexpr int kPostDS = 132;       // row stride (floats) of ds: K <= 128, 4-bank skew between channels
template <int IDF>
__global__ void __launch_bounds__(256) k_bwd_post_tc5(const float* __restrict__ dSrc, const float* __restrict__ ctx,
                                                      const float* __restrict__ W, float* __restrict__ dW,
                                                      float* __restrict__ dCtx, int B, int cdf, int L, int n_dw, int tl) {
    extern __shared__ __align__(16) float sm[];
    constexpr bool HI = IDF > 32;          // a thread owns channel i0 and, for idf > 32, i0 + 32
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // (the head kernel of the next call, see above)
    if (threadIdx.x == 0) tl_min(tl < 0 ? tl : tl + 6);
    if ((int)blockIdx.x < n_dw) {
        float* cs = sm;                          // [K <= 128][kPostCS]
        float* ds = sm + 128 * kPostCS;          // [64][kPostDS]
        const int cg = blockIdx.x >> 6, grp = blockIdx.x & 63;
        const int c0 = cg * 32, nc = cdf - c0 < 32 ? cdf - c0 : 32;
        const int b_lo = (B * grp) >> 6, b_hi = (B * (grp + 1)) >> 6;
        const int i0 = tid >> 3, cq = (tid & 7) * 4;
        const int row = tid >> 1, half = tid & 1;          // staging: thread = half a row of L words
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        bool waited = false;
        for (int bb = b_lo; bb < b_hi; bb += 4) {
            const int nb = b_hi - bb < 4 ? b_hi - bb : 4;
            const int K = nb * L;
            __syncthreads();
            // ctx rows (sample s, channel c) - 4 x 32 = 128 rows - do not depend on the streaming kernel: they are
            // staged BEFORE griddepcontrol.wait, while that grid is still running
            {
                const int sidx = row >> 5, c = row & 31;
                float cv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int l = half * 16 + j;
                    cv[j] = (sidx < nb && c < nc && l < L) ? __ldg(ctx + ((size_t)(bb + sidx) * cdf + c0 + c) * L + l) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int l = half * 16 + j;
                    if (sidx < nb && l < L) cs[(sidx * L + l) * kPostCS + c] = cv[j];
                }
            }
            if (!waited) {
                asm volatile("griddepcontrol.wait;" ::: "memory");      // dSrc is complete and visible
                if (threadIdx.x == 0) tl_min(tl < 0 ? tl : tl + 4);
                waited = true;
            }
            // dSrc rows (sample s, channel i) - up to 4 x 64 = 256 rows, two passes of 128 rows; all loads of the
            // round are issued before the first store (one memory round trip)
            float dv[2][16];
#pragma unroll
            for (int ps = 0; ps < (HI ? 2 : 1); ++ps) {
                const int r = row + 128 * ps, sidx = r / IDF, i = r - sidx * IDF;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int l = half * 16 + j;
                    dv[ps][j] = (sidx < nb && l < L) ? __ldcg(dSrc + ((size_t)(bb + sidx) * IDF + i) * L + l) : 0.f;
                }
            }
#pragma unroll
            for (int ps = 0; ps < (HI ? 2 : 1); ++ps) {
                const int r = row + 128 * ps, sidx = r / IDF, i = r - sidx * IDF;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int l = half * 16 + j;
                    if (sidx < nb && l < L) ds[i * kPostDS + sidx * L + l] = dv[ps][j];
                }
            }
            __syncthreads();
            const float* d0p = ds + i0 * kPostDS;
            const float* ccol = cs + cq;
#pragma unroll 8
            for (int k = 0; k < K; ++k) {
                const float4 c4 = *reinterpret_cast<const float4*>(ccol + k * kPostCS);
                const float d0 = d0p[k];
                acc[0][0] = fmaf(d0, c4.x, acc[0][0]); acc[0][1] = fmaf(d0, c4.y, acc[0][1]);
                acc[0][2] = fmaf(d0, c4.z, acc[0][2]); acc[0][3] = fmaf(d0, c4.w, acc[0][3]);
                if constexpr (HI) {
                    const float d1 = d0p[32 * kPostDS + k];       // rows >= IDF are never staged; results discarded
                    acc[1][0] = fmaf(d1, c4.x, acc[1][0]); acc[1][1] = fmaf(d1, c4.y, acc[1][1]);
                    acc[1][2] = fmaf(d1, c4.z, acc[1][2]); acc[1][3] = fmaf(d1, c4.w, acc[1][3]);
                }
            }
        }
        if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");     // (empty sample group) dW is zeroed upstream
#pragma unroll
        for (int h = 0; h < (HI ? 2 : 1); ++h) {
            const int i = i0 + 32 * h;
            if (i < IDF)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (cq + k < nc) atomicAdd(dW + (size_t)i * cdf + c0 + cq + k, acc[h][k]);
        }
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        float* ds = sm;                  // [idf][L]
        const int b = blockIdx.x - n_dw;
        for (int o = tid; o < IDF * L; o += blockDim.x) ds[o] = __ldcg(dSrc + (size_t)b * IDF * L + o);
        __syncthreads();
        for (int o = tid; o < cdf * L; o += blockDim.x) {
            const int c = o / L, l = o - c * L;
            float a = 0.f;
            for (int i = 0; i < IDF; ++i) a = fmaf(__ldg(W + (size_t)i * cdf + c), ds[i * L + l], a);
            dCtx[(size_t)b * cdf * L + o] = a;
        }
    }
    if (threadIdx.x == 0) tl_max(tl < 0 ? tl : tl + 5);
}

__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

template <int IDF, int NQ, bool HAS_GA, int NST_>
__global__ void __launch_bounds__(kBwdThreads, Tc5BwdCfg<IDF, NQ, NST_>::CTAS_PER_SM)
    k_attn_bwd_tc5(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                   const __grid_constant__ CUtensorMap tm_dx, const Tc5BwdParams p) {
    using C = Tc5BwdCfg<IDF, NQ, NST_>;
    using T = __nv_bfloat16;
    constexpr int LP = C::LP, RP = C::RP, NST = C::NST;
    constexpr float kLog2e = 1.4426950408889634f;

    // no static shared memory in this kernel: the dynamic segment starts 1024-byte aligned (swizzle atoms)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    unsigned char* sgen = smem_raw;
    const uint32_t s_st = sbase;                                  // [NST] staged g / x tiles
    const uint32_t s_pb = s_st + NST * C::STAGE_BYTES;            // PB: [64-px block][row][128 B]
    const uint32_t s_b1 = s_pb + 2 * C::PB_BYTES;                 // sourceT, rows = words
    const uint32_t s_b2 = s_b1 + C::B1_BYTES;                     // sourceT, rows = channels
    const uint32_t s_out = s_b2 + C::B2_BYTES;                    // [4 warps] dX staging
    unsigned char* g_pb = sgen + NST * C::STAGE_BYTES;
    unsigned char* g_b1 = g_pb + 2 * C::PB_BYTES;
    unsigned char* g_b2 = g_b1 + C::B1_BYTES;
    unsigned char* g_out = g_b2 + C::B2_BYTES;

    unsigned long long* bars = reinterpret_cast<unsigned long long*>(g_out + 4 * C::OUT_WARP_BYTES);
    unsigned long long* bar_x_full = bars;
    unsigned long long* bar_x_empty = bars + NST;
    unsigned long long* bar_s_full = bars + 2 * NST;          // [2] by tile parity
    unsigned long long* bar_s_free = bars + 2 * NST + 2;      // [2]
    unsigned long long* bar_ds_ready = bars + 2 * NST + 11;   // [2] by tile parity: the first-stage warps may be a tile apart
    unsigned long long& bar_c_full = bars[2 * NST + 5];
    unsigned long long& bar_dx_free = bars[2 * NST + 6];
    unsigned long long& bar_b_ready = bars[2 * NST + 7];
    unsigned long long* bar_pb_free = bars + 2 * NST + 8;     // [2] MMA2 of that parity has completed (PB, operands)
    unsigned long long& bar_acc_free = bars[2 * NST + 10];    // the dSrc accumulator of a finished sample has been read
    uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(bars + 2 * NST + 13);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L, Q = p.Q, TPS = p.tiles_per_sample;
    // let the post kernel's blocks become resident wherever there is room; they park in griddepcontrol.wait
    // until this grid has completed, which takes its launch latency off the critical path
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) tl_min(p.tl < 0 ? p.tl : p.tl + 2);
    if (p.trace != nullptr && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        p.trace[256 + 2 * blockIdx.x] = (long long)gt;
    }
    if (p.trace != nullptr && blockIdx.x == 0 && tid == 64) p.trace[240] = clock64();

    const int w_begin = (int)(((long long)blockIdx.x * p.n_tiles) / gridDim.x);
    const int w_end = (int)(((long long)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    const int n_local = w_end - w_begin;
    const int b0 = w_begin / TPS, t0 = w_begin - b0 * TPS;
    const int n_pre = n_local < NST ? n_local : NST;       // tiles whose loads thread 0 issues before the prologue

    if (tid == 0) {
        if (sbase & 1023u) __trap();
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            mbar_init(smem_u32(&bar_x_full[s]), 1);
            mbar_init(smem_u32(&bar_x_empty[s]), 1);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_s_full[s]), 1);
            mbar_init(smem_u32(&bar_s_free[s]), 4);
        }
        mbar_init(smem_u32(&bar_pb_free[0]), 1);
        mbar_init(smem_u32(&bar_pb_free[1]), 1);
        mbar_init(smem_u32(&bar_acc_free), 4);
        mbar_init(smem_u32(&bar_ds_ready[0]), 4);
        mbar_init(smem_u32(&bar_ds_ready[1]), 4);
        mbar_init(smem_u32(&bar_c_full), 1);
        mbar_init(smem_u32(&bar_dx_free), 4);
        mbar_init(smem_u32(&bar_b_ready), 4);
        fence_barrier_init();
        // the first ring of g / x tiles is requested before the rest of the prologue (TMEM allocation, operand
        // buffers) so that its DRAM latency runs under it
        {
            int b = b0, t = t0;
            for (int j = 0; j < n_pre; ++j) {
                const uint32_t full = smem_u32(&bar_x_full[j]);
                mbar_expect_tx(full, (uint32_t)C::STAGE_BYTES);
                const uint32_t dst = s_st + j * C::STAGE_BYTES;
#pragma unroll
                for (int bx = 0; bx < C::NBOX; ++bx) {
                    tma_load_2d(dst + (2 * bx) * C::BOX_BYTES, &tm_g, t * TQ + bx * C::BOX_PX, b * IDF, full);
                    tma_load_2d(dst + (2 * bx + 1) * C::BOX_BYTES, &tm_x, t * TQ + bx * C::BOX_PX, b * IDF, full);
                }
                if (++t == TPS) { t = 0; ++b; }
            }
        }
        prefetch_tensormap(&tm_dx);
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
    // zero PB and the B operand buffers once: padding rows / words are never written again
    for (int o = tid; o < (2 * C::PB_BYTES + C::B1_BYTES + C::B2_BYTES) / 16; o += kBwdThreads)
        reinterpret_cast<uint4*>(g_pb)[o] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);     // provably warp-uniform
    if (p.trace != nullptr && blockIdx.x == 0 && tid == 64) p.trace[241] = clock64();


    if (warp == kProducerWarp) {
        // --------------------------------- TMA producer -----------------------------------------
        // (the whole warp runs the loop; one elected lane issues - see elect_one())
        int b = b0, t = t0 + n_pre;
        while (t >= TPS) { t -= TPS; ++b; }
        for (int j = n_pre; j < n_local; ++j) {
            const int stage = j % NST;
            mbar_wait(smem_u32(&bar_x_empty[stage]), (uint32_t)((j / NST) - 1) & 1u);
            const uint32_t full = smem_u32(&bar_x_full[stage]);
            const uint32_t dst = s_st + stage * C::STAGE_BYTES;
            if (elect_one()) {
                mbar_expect_tx(full, (uint32_t)C::STAGE_BYTES);
#pragma unroll
                for (int bx = 0; bx < C::NBOX; ++bx) {
                    tma_load_2d(dst + (2 * bx) * C::BOX_BYTES, &tm_g, t * TQ + bx * C::BOX_PX, b * IDF, full);
                    tma_load_2d(dst + (2 * bx + 1) * C::BOX_BYTES, &tm_x, t * TQ + bx * C::BOX_PX, b * IDF, full);
                }
            }
            __syncwarp();
            if (++t == TPS) { t = 0; ++b; }
        }
    } else if (warp == kMmaWarp) {
        // --------------------------------- MMA issuer -------------------------------------------
        // (the whole warp runs the loop; one elected lane issues - see elect_one())
        // descriptor halves (tc5_common.cuh)
        constexpr uint32_t kTileMnHi = desc_hi(1024, kSwizzle128B), kKHi = desc_hi(1024, kSwizzle128B);
        constexpr uint32_t kBHi1 = desc_hi(C::KCH1 * 128, kSwizzleNone), kBHi2 = desc_hi(C::KCH2 * 128, kSwizzleNone);
        const uint32_t tg_lo = desc_lo(s_st, 2 * C::BOX_BYTES), tx_lo = desc_lo(s_st + C::BOX_BYTES, 2 * C::BOX_BYTES);
        const uint32_t tk_lo = desc_lo(s_st, 16);                                  // [g ; x] as one K-major operand
        const uint32_t pbA_lo = desc_lo(s_pb + RP * 128, C::PB_KBLOCK), pbB_lo = desc_lo(s_pb, 16);
        const uint32_t b1_lo = desc_lo(s_b1, 128), b2_lo = desc_lo(s_b2, 128);
        // MMA1(j): S, dP of tile j into the {S, dP} buffer of its parity
        auto mma1 = [&](int j) {
            const int stage = j % NST, buf = j & 1;
            mbar_wait(smem_u32(&bar_x_full[stage]), (uint32_t)(j / NST) & 1u);
            if (j >= 2) mbar_wait(smem_u32(&bar_s_free[buf]), (uint32_t)((j >> 1) - 1) & 1u);
            tc_fence_after();
            const uint32_t st_lo = (uint32_t)(stage * (C::STAGE_BYTES >> 4));
            const uint32_t dS_ = tmem_base + C::COL_BUF * buf + C::COL_S, dP_ = tmem_base + C::COL_BUF * buf + C::COL_DP;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < C::KS1; ++ks) {
                    const uint32_t ka = (uint32_t)(ks * 128), kb = (uint32_t)(ks * 16);
                    umma_ss<false>(dS_, tx_lo + st_lo + ka, kTileMnHi, b1_lo + kb, kBHi1, C::IDESC1, ks > 0 ? 1u : 0u);
                    umma_ss<false>(dP_, tg_lo + st_lo + ka, kTileMnHi, b1_lo + kb, kBHi1, C::IDESC1, ks > 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&bar_s_full[buf]));
            }
            __syncwarp();
        };
        // MMA2(j): dX = dS . B2 and acc (+)= [g ; x] . [P | dS] from PB[j & 1]
        uint32_t n_flush = 0;        // accumulator hand-backs waited for so far
        auto mma2 = [&](int j, bool first_of_sample) {
            const int stage = j % NST;
            mbar_wait(smem_u32(&bar_ds_ready[j & 1]), (uint32_t)(j >> 1) & 1u);
            if (j > 0) mbar_wait(smem_u32(&bar_dx_free), (uint32_t)(j - 1) & 1u);      // dX(j-1) is in registers
            if (first_of_sample && j > 0) {                                            // the previous sample's accumulator
                mbar_wait(smem_u32(&bar_acc_free), n_flush & 1u);                      // has been read out
                ++n_flush;
            }
            tc_fence_after();
            const uint32_t st_lo = (uint32_t)(stage * (C::STAGE_BYTES >> 4));
            const uint32_t pb_off = (uint32_t)((j & 1) * (C::PB_BYTES >> 4));
            if (elect_one()) {
                // (A = dS rows of PB, MN-major: pixel blocks PB_KBLOCK apart)
#pragma unroll
                for (int ks = 0; ks < C::KS2; ++ks)
                    umma_ss<false>(tmem_base + C::COL_DX, pbA_lo + pb_off + (uint32_t)(ks * 128), kTileMnHi,
                                   b2_lo + (uint32_t)(ks * 16), kBHi2, C::IDESC2, ks > 0 ? 1u : 0u);
                // (both K-major, K = pixels: 16 per step, 4 steps per 64-px block)
#pragma unroll
                for (int ks = 0; ks < C::KS3; ++ks)
                    umma_ss<false>(tmem_base + C::COL_ACC,
                                   tk_lo + st_lo + (uint32_t)((ks >> 2) * (2 * C::BOX_BYTES >> 4) + (ks & 3) * 2), kKHi,
                                   pbB_lo + pb_off + (uint32_t)((ks >> 2) * (C::PB_KBLOCK >> 4) + (ks & 3) * 2), kKHi,
                                   C::IDESC3, (ks > 0 || !first_of_sample) ? 1u : 0u);
                umma_commit(smem_u32(&bar_x_empty[stage]));
                umma_commit(smem_u32(&bar_pb_free[j & 1]));
                umma_commit(smem_u32(&bar_c_full));
            }
            __syncwarp();
        };
        // Software pipeline: S/dP of tile j+1 are produced BEFORE the second-stage MMAs of tile j, so the
        // consumers' softmax of tile j+1 overlaps MMA2(j).  At a sample boundary the pipeline drains: the
        // operands of the next sample are rebuilt only after every MMA of this one has completed.
        int t = t0;
        uint32_t nb = 0;
        if (n_local > 0) {
            mbar_wait(smem_u32(&bar_b_ready), nb & 1u);
            ++nb;
            mma1(0);
        }
        bool first_of_sample = true;
        for (int j = 0; j < n_local; ++j) {
            const bool has_next = j + 1 < n_local;
            const bool next_same = has_next && (t + 1 < TPS);
            if (next_same) mma1(j + 1);
            mma2(j, first_of_sample);
            first_of_sample = false;
            if (has_next && !next_same) {
                mbar_wait(smem_u32(&bar_b_ready), nb & 1u);     // operands of the next sample are in place
                ++nb;
                mma1(j + 1);
                first_of_sample = true;
            }
            if (++t == TPS) t = 0;
        }
    } else if (warp < kFirstEpilogueWarp) {
        // --------------------------------- first-stage warps: thread = pixel --------------------
        // S, dP -> P, dS -> PB.  They run ahead of the second stage by up to one tile (two {S, dP} buffers,
        // two PB buffers) and rebuild the operands at sample boundaries.
        const int ct = tid - 64;
        const int cw = warp & 3;
        const int px = cw * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(cw * 32) << 16);
        const uint32_t pad_bits = (L < 32) ? ~((1u << L) - 1u) : 0u;
        const uint32_t Bu = (uint32_t)p.B;
        const uint32_t step_mod = (uint32_t)TQ % Bu;
        // reference mask order: pixel n = b*Q + q uses caption n mod B (GlobalAttention.py:104-108)
        uint32_t cap = (uint32_t)(((unsigned long long)w_begin * TQ + px) % Bu);
        int b = b0, t = t0, cur_b = -1;
        // PB element (row n, this pixel): block (px / 64), row n, 16-byte chunks XOR-swizzled with n & 7
        unsigned char* pb_px = g_pb + (px >> 6) * C::PB_KBLOCK;
        const uint32_t pb_col = (uint32_t)(px & 63) * 2;

        for (int j = 0; j < n_local; ++j) {
            const int buf = j & 1;
            if (b != cur_b) {
                // ---- operands of sample b: B1[word][channel] = B2[channel][word] = srcT -----------------
                // every MMA that read the previous sample's operands has completed: MMA2(j - 1)
                if (j > 0) mbar_wait(smem_u32(&bar_pb_free[(j - 1) & 1]), (uint32_t)((j - 1) >> 1) & 1u);
                cur_b = b;
                const float* sb = p.srcT + (size_t)b * IDF * L;
                // thread -> (channel ct / 4 [+ 32 g], words (ct % 4) + 4 k): all loads of a thread are issued
                // before the first is consumed (one L2 round trip), and no division by the runtime L
                constexpr int NG = IDF / 32 + (IDF % 32 != 0), NK = LP / 4;
                float sv[NG][NK];
#pragma unroll
                for (int g = 0; g < NG; ++g)
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const int ch = (ct >> 2) + 32 * g, l = (ct & 3) + 4 * k;
                        sv[g][k] = (ch < IDF && l < L) ? __ldg(sb + ch * L + l) : 0.f;
                    }
#pragma unroll
                for (int g = 0; g < NG; ++g)
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const int ch = (ct >> 2) + 32 * g, l = (ct & 3) + 4 * k;
                        if (ch < IDF && l < L) {
                            const __nv_bfloat16 v = __float2bfloat16_rn(sv[g][k]);
                            *reinterpret_cast<__nv_bfloat16*>(g_b1 + kmajor_off<2>(l, ch, C::KCH1)) = v;
                            *reinterpret_cast<__nv_bfloat16*>(g_b2 + kmajor_off<2>(ch, l, C::KCH2)) = v;
                        }
                    }
                fence_proxy_async();
                warp_arrive(smem_u32(&bar_b_ready), lane);
            }
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && ct == 0 && j < 16;
            if (tr) p.trace[j * 16 + 0] = clock64();
            const int q = t * TQ + px;
            uint32_t mb = pad_bits;
            if (p.mask != nullptr) mb |= __ldg(p.mask_bits + (p.mask_mode == SBA_MASK_PER_SAMPLE ? (uint32_t)b : cap));
            float ga[HAS_GA ? LP : 1];
            if constexpr (HAS_GA) {
                const T* gp = static_cast<const T*>(p.ga) + (size_t)b * L * Q + q;
#pragma unroll
                for (int l = 0; l < LP; ++l) ga[l] = (l < L) ? __bfloat162float(gp[(size_t)l * Q]) : 0.f;
            }
            // ---- S and dP rows of this pixel ---------------------------------------------------------
            mbar_wait(smem_u32(&bar_s_full[buf]), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            if (tr) p.trace[j * 16 + 1] = clock64();
            uint32_t sr[LP], dr[LP];
            tmem_ld<LP>(tl + C::COL_BUF * buf + C::COL_S, sr);
            tmem_ld<LP>(tl + C::COL_BUF * buf + C::COL_DP, dr);
            tmem_wait_ld();
            tc_fence_before();
            warp_arrive(smem_u32(&bar_s_free[buf]), lane);
            // ---- P = masked softmax over words (recomputed; GlobalAttention.py:104-109) ---------------
            float s[LP];
            float m = -INFINITY;
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                s[l] = ((mb >> l) & 1u) ? -INFINITY : __uint_as_float(sr[l]);
                m = fmaxf(m, s[l]);
            }
            const float ml = m * kLog2e;
            float sum = 0.f;
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                s[l] = mma::ex2_approx(fmaf(s[l], kLog2e, -ml));    // all-masked row: NaN, as the reference
                sum += s[l];
            }
            const float inv = mma::rcp_approx(sum);
            // ---- dS = P * (dP [+ g_attn] - sum_l P dP)  (masked / padded words have P = 0) --------------
            float d[LP];
            float dot = 0.f;
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                s[l] *= inv;
                d[l] = __uint_as_float(dr[l]);
                if constexpr (HAS_GA) d[l] += ga[l];
                dot = fmaf(s[l], d[l], dot);
            }
            // ---- P and dS into PB[buf]: rows [0, LP) and [RP, RP + LP), bf16, swizzled --------------------
            if (j >= 2) mbar_wait(smem_u32(&bar_pb_free[buf]), (uint32_t)((j >> 1) - 1) & 1u);    // MMA2(j - 2) is done with it
            unsigned char* pb = pb_px + buf * C::PB_BYTES;
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                const float ds = s[l] * (d[l] - dot);
                *reinterpret_cast<__nv_bfloat16*>(pb + l * 128 + (pb_col ^ ((l & 7) << 4))) = __float2bfloat16_rn(s[l]);
                *reinterpret_cast<__nv_bfloat16*>(pb + (RP + l) * 128 + (pb_col ^ ((l & 7) << 4))) = __float2bfloat16_rn(ds);
            }
            fence_proxy_async();
            warp_arrive(smem_u32(&bar_ds_ready[buf]), lane);
            if (tr) p.trace[j * 16 + 3] = clock64();

            if (++t == TPS) { t = 0; ++b; }
            cap += step_mod;
            if (cap >= Bu) cap -= Bu;
        }
    } else {
        // --------------------------------- second-stage warps: thread = pixel -------------------
        // dX row of the pixel -> staged [channel][32 px] per warp -> one TMA box store; at the end of a sample
        // the diagonal blocks of the TMEM accumulator are added to dSrc[b] with fp32 atomics.
        const int cw = warp & 3;
        const uint32_t tl = tmem_base + ((uint32_t)(cw * 32) << 16);
        const uint32_t so = s_out + cw * C::OUT_WARP_BYTES;
        T* go = reinterpret_cast<T*>(g_out + cw * C::OUT_WARP_BYTES) + lane;
        int b = b0, t = t0;
        bool waited_zero = false;
        for (int j = 0; j < n_local; ++j) {
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && warp == kFirstEpilogueWarp && lane == 0 && j < 16;
            mbar_wait(smem_u32(&bar_c_full), (uint32_t)j & 1u);
            tc_fence_after();
            if (tr) p.trace[j * 16 + 5] = clock64();
            uint32_t cr[IDF];
            tmem_ld<IDF>(tl + C::COL_DX, cr);
            tmem_wait_ld();
            tc_fence_before();
            warp_arrive(smem_u32(&bar_dx_free), lane);
            if (tr) p.trace[j * 16 + 6] = clock64();
            if (lane == 0) bulk_wait_read<0>();        // the previous dX store has finished reading the staging
            __syncwarp();
#pragma unroll
            for (int i = 0; i < IDF; ++i) go[i * 32] = __float2bfloat16_rn(__uint_as_float(cr[i]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_dx, t * TQ + cw * 32, b * IDF, so);
                bulk_commit();
            }
            if (tr) p.trace[j * 16 + 7] = clock64();

            const bool last_of_sample = (t + 1 == TPS) || (j + 1 == n_local);
            if (last_of_sample) {
                // every MMA of sample b issued by this CTA has completed (c_full(j)): add its share of dSrc[b]
                if (!waited_zero) {
                    asm volatile("griddepcontrol.wait;" ::: "memory");      // k_zero_tc5 has cleared dSrc / dW
                    waited_zero = true;
                }
                // accumulator row r of this lane: M = 64 -> lanes 0..15 of each quarter hold rows 16*cw + lane
                const int row = C::MD == 64 ? 16 * cw + lane : 32 * cw + lane;
                const bool valid = (C::MD == 64 ? lane < 16 : true) && row < 2 * IDF;
                const bool is_x = row >= IDF;
                const int ch = is_x ? row - IDF : row;
                // column block: g rows take the P columns [0, LP), x rows the dS columns [RP, RP + LP)
                uint32_t a0[LP], a1[LP];
                tc_fence_after();
                tmem_ld<LP>(tl + C::COL_ACC, a0);
                tmem_ld<LP>(tl + C::COL_ACC + RP, a1);
                tmem_wait_ld();
                tc_fence_before();
                if (j + 1 < n_local) warp_arrive(smem_u32(&bar_acc_free), lane);     // the next sample may overwrite it
                if (valid) {
                    float* db = p.dSrc + ((size_t)b * IDF + ch) * L;
#pragma unroll
                    for (int l = 0; l < LP; ++l)
                        if (l < L) atomicAdd(db + l, __uint_as_float(is_x ? a1[l] : a0[l]));
                }
            }
            if (++t == TPS) { t = 0; ++b; }
        }
        if (lane == 0) bulk_wait<0>();
        tc_fence_before();
    }

    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (p.trace != nullptr && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        p.trace[256 + 2 * blockIdx.x + 1] = (long long)gt;
    }
    if (tid == 0) tl_max(p.tl < 0 ? p.tl : p.tl + 3);
}

template <int IDF, int NQ, bool HAS_GA, int NST_>
int launch_bwd_tc5(const void* x, const void* g, void* dX, const Tc5BwdParams& p, cudaStream_t st) {
    using C = Tc5BwdCfg<IDF, NQ, NST_>;
    auto kern = k_attn_bwd_tc5<IDF, NQ, HAS_GA, NST_>;
    const size_t smem = (size_t)C::SMEM_BYTES + (2 * C::NST + 14) * 8;
    // per device (a process may drive several): SM count, and whether this kernel's dynamic shared memory
    // limit has been raised there (smem is a compile-time constant of the instantiation)
    static int sms_of[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) {
        set_error("%s(tcgen05): device index %d not supported", "attn_bwd", dev);
        return SBA_ERR_UNSUPPORTED;
    }
    if (smem > 220 * 1024) {
        set_error("%s(tcgen05): %zu bytes of shared memory needed", "attn_bwd", smem);
        return SBA_ERR_UNSUPPORTED;
    }
    if (sms_of[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess || n < 1) {
            set_error("%s(tcgen05): cudaFuncSetAttribute(%zu B): %s", "attn_bwd", smem, cudaGetErrorString(e));
            return SBA_ERR_CUDA;
        }
        sms_of[dev] = n;
    }
    const int sms = sms_of[dev];
    const size_t smem_set = smem;
    int per_sm = (int)((227 * 1024) / (smem_set + 1024));
    if (per_sm > C::CTAS_PER_SM) per_sm = C::CTAS_PER_SM;
    if (per_sm < 1) per_sm = 1;
    if (getenv("SBA_TC5_CTAS_PER_SM")) per_sm = atoi(getenv("SBA_TC5_CTAS_PER_SM"));
    const int max_ctas = sms * per_sm;
    CUtensorMap tm_x, tm_g, tm_dx;
    int rc = make_tile_map(&tm_x, x, SBA_BF16, p.B * IDF, p.Q, IDF, 64, true);
    if (!rc) rc = make_tile_map(&tm_g, g, SBA_BF16, p.B * IDF, p.Q, IDF, 64, true);
    if (!rc) rc = make_tile_map(&tm_dx, dX, SBA_BF16, p.B * IDF, p.Q, IDF, 32, false);
    if (rc) return rc;
    const size_t n_src = (size_t)p.B * IDF * p.L + p.B + 1;       // dSrc and the counter words behind it
    rc = attn_bwd_zero(p.dSrc, n_src, p.dW, p.dW ? (size_t)IDF * p.cdf : 0, st);
    if (rc) return rc;
    const int grid = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kBwdThreads);
    cfg.dynamicSmemBytes = smem_set;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    Tc5BwdParams pk = p;
    pk.tl = timeline_slot();
    static long long* trace_buf = nullptr;
    if (getenv("SBA_TC5_TRACE")) {
        if (!trace_buf) cudaMalloc(&trace_buf, 4096 * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, 4096 * sizeof(long long), st);
        pk.trace = trace_buf;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_x, tm_g, tm_dx, pk);
    if (e != cudaSuccess) {
        set_error("attn_bwd(tcgen05): launch: %s", cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    if (pk.trace) {
        static long long h[4096];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        {
            long long t0 = h[256];
            for (int k = 0; k < grid && k < 1900; ++k) if (h[256 + 2 * k] < t0) t0 = h[256 + 2 * k];
            {
                double sum = 0; long long mx = 0, mn = 1LL << 62; int n = 0;
                for (int k = 0; k < grid && k < 1900; ++k) {
                    const long long d = h[256 + 2 * k + 1] - h[256 + 2 * k];
                    sum += (double)d; ++n;
                    if (d > mx) mx = d;
                    if (d < mn) mn = d;
                }
                fprintf(stderr, "CTA lifetime over %d CTAs: min %lld mean %.0f max %lld ns\n", n, mn, sum / n, mx);
            }
            fprintf(stderr, "CTA start/end (ns since the first start), every 37th CTA:\n");
            for (int k = 0; k < grid && k < 1900; k += 37)
                fprintf(stderr, "  cta %4d: %7lld .. %7lld\n", k, h[256 + 2 * k] - t0, h[256 + 2 * k + 1] - t0);
        }
        const char* names[12] = {"top", "s_full", "ld_S", "math+PB", "ds_rdy", "c_full", "dX_stg", "store", "M:x_full", "M:mma1", "M:ds_rdy", "M:mma2"};
        fprintf(stderr, "CTA 0: entry %lld, prologue done +%lld, B operands built +%lld, first tile top +%lld (cycles)\n", 0LL,
                h[241] - h[240], h[242] - h[240], h[0] - h[240]);
        fprintf(stderr, "tile");
        for (int k = 0; k < 12; ++k) fprintf(stderr, " %9s", names[k]);
        fprintf(stderr, "   (cycles since the first stamp)\n");
        for (int j = 0; j < 16 && h[j * 16] != 0; ++j) {
            fprintf(stderr, "%4d", j);
            for (int k = 0; k < 12; ++k) fprintf(stderr, " %9lld", h[j * 16 + k] - h[0]);
            fprintf(stderr, "\n");
        }
    }
    add_launches(1);
    rc = check_launch("attn_bwd(tcgen05)");
    if (rc) return rc;
    return attn_bwd_post(p.dSrc, p.ctx, p.W, p.dW, p.dCtx, p.B, IDF, p.cdf, p.L, st);
}

template <int IDF, bool HAS_GA>
int dispatch_nq(const void* x, const void* g, void* dX, const Tc5BwdParams& p, cudaStream_t st) {
    // long streams at idf 32 (>= 16 tiles per CTA of a 2-per-SM grid, e.g. 128x128 at B >= 40) take the 4-deep ring:
    // it still fits two CTAs per SM there and measured 2-4 % faster; short streams lose to its longer prologue
    constexpr bool kDeep = IDF == 32;
    const bool deep = kDeep && p.n_tiles >= 16 * 2 * 148;
#define SBA_BWD_CASE(n)                                                                     \
    case n:                                                                                 \
        if constexpr (kDeep) {                                                              \
            if (deep) return launch_bwd_tc5<IDF, n, HAS_GA, 4>(x, g, dX, p, st);            \
        }                                                                                   \
        return launch_bwd_tc5<IDF, n, HAS_GA, 3>(x, g, dX, p, st);
    switch ((p.L + 3) / 4) {
        SBA_BWD_CASE(1) SBA_BWD_CASE(2) SBA_BWD_CASE(3) SBA_BWD_CASE(4)
        SBA_BWD_CASE(5) SBA_BWD_CASE(6) SBA_BWD_CASE(7) SBA_BWD_CASE(8)
        default: return -1;
    }
#undef SBA_BWD_CASE
}

template <int IDF>
int dispatch_ga(const void* x, const void* g, void* dX, const Tc5BwdParams& p, cudaStream_t st) {
    return p.ga != nullptr ? dispatch_nq<IDF, true>(x, g, dX, p, st) : dispatch_nq<IDF, false>(x, g, dX, p, st);
}

}  // namespace

int attn_bwd_zero(float* dSrc, size_t n_src, float* dW, size_t n_dw, cudaStream_t st) {
    if (getenv("SBA_TC5_TIMELINE")) {
        ++g_tl_call;
        if (g_tl_call == 20) {
            unsigned long long init[16 * 8];
            for (int i = 0; i < 16 * 8; ++i) init[i] = (i % 8 == 0 || i % 8 == 2 || i % 8 == 4 || i % 8 == 6) ? ~0ull : 0ull;
            cudaMemcpyToSymbol(g_timeline, init, sizeof(init));
        }
        if (g_tl_call == 37) {
            unsigned long long h[16 * 8];
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(h, g_timeline, sizeof(h));
            fprintf(stderr, "call: zero start..end | main start..end | post entry, past wait..end   (us since zero start of the first call)\n");
            for (int c = 0; c < 16; ++c) {
                const unsigned long long* r = h + c * 8;
                auto us = [&](unsigned long long v) { return (double)(long long)(v - h[0]) * 1e-3; };
                fprintf(stderr, "%3d: %8.2f..%8.2f | %8.2f..%8.2f | %8.2f, %8.2f..%8.2f\n", c, us(r[0]), us(r[1]), us(r[2]), us(r[3]),
                        us(r[6]), us(r[4]), us(r[5]));
            }
        }
    }
    {
        cudaLaunchConfig_t zc = {};
        zc.gridDim = dim3(64);
        zc.blockDim = dim3(256);
        zc.stream = st;
        cudaLaunchAttribute za[1];
        za[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        za[0].val.programmaticStreamSerializationAllowed = 1;
        zc.attrs = za;
        zc.numAttrs = 1;
        cudaError_t ze = cudaLaunchKernelEx(&zc, k_zero_tc5, dSrc, n_src, dW, n_dw, timeline_slot());
        if (ze != cudaSuccess) {
            set_error("attn_bwd(zero): launch: %s", cudaGetErrorString(ze));
            return SBA_ERR_CUDA;
        }
    }
    add_launches(1);
    return check_launch("attn_bwd(zero)");
}

int attn_bwd_post(const float* dSrc, const float* ctx, const float* W, float* dW, float* dCtx, int B, int idf, int cdf,
                  int L, cudaStream_t st) {
    if (dW == nullptr && dCtx == nullptr) return SBA_OK;
    const int n_dw = dW != nullptr ? 64 * ((cdf + 31) / 32) : 0;
    const int grid = n_dw + (dCtx != nullptr ? B : 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = (size_t)(128 * kPostCS + 64 * kPostDS) * sizeof(float);       // 52 KB: above the default limit
    static bool attr_set_of[64] = {false};
    int pdev = 0;
    cudaGetDevice(&pdev);
    bool& attr_set = attr_set_of[pdev & 63];
    if (!attr_set) {
        cudaFuncSetAttribute(k_bwd_post_tc5<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        cudaFuncSetAttribute(k_bwd_post_tc5<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        cudaFuncSetAttribute(k_bwd_post_tc5<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        attr_set = true;
    }
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int tl = timeline_slot();
    cudaError_t e;
    if (idf == 32) e = cudaLaunchKernelEx(&cfg, k_bwd_post_tc5<32>, dSrc, ctx, W, dW, dCtx, B, cdf, L, n_dw, tl);
    else if (idf == 48) e = cudaLaunchKernelEx(&cfg, k_bwd_post_tc5<48>, dSrc, ctx, W, dW, dCtx, B, cdf, L, n_dw, tl);
    else if (idf == 64) e = cudaLaunchKernelEx(&cfg, k_bwd_post_tc5<64>, dSrc, ctx, W, dW, dCtx, B, cdf, L, n_dw, tl);
    else {
        set_error("attn_bwd(post): idf=%d not covered", idf);
        return SBA_ERR_UNSUPPORTED;
    }
    if (e != cudaSuccess) {
        set_error("attn_bwd(post): launch: %s", cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    add_launches(1);
    return check_launch("attn_bwd(post)");
}

bool tc5_bwd_supports(const AttnShape& s) { return tc5_supports(s) && s.dtype == SBA_BF16; }

int tc5_attn_bwd(const void* x, const float* ctx, const float* W, const float* srcT, const uint8_t* mask,
                 const uint32_t* mask_bits, const void* g_c,
                 const void* g_attn, void* dX, float* dSrc, float* dW, float* dCtx, const AttnShape& s, cudaStream_t st) {
    Tc5BwdParams p{};
    p.srcT = srcT; p.mask = mask; p.mask_bits = mask_bits; p.ga = g_attn; p.dSrc = dSrc; p.ctx = ctx; p.W = W; p.dW = dW; p.dCtx = dCtx;
    p.B = s.B; p.L = s.L; p.Q = s.Q; p.cdf = s.cdf; p.mask_mode = s.mask_mode;
    p.tiles_per_sample = s.Q / tc5::TQ;
    p.n_tiles = s.B * p.tiles_per_sample;
    int rc = -1;
    if (s.idf == 32) rc = dispatch_ga<32>(x, g_c, dX, p, st);
    else if (s.idf == 48) rc = dispatch_ga<48>(x, g_c, dX, p, st);
    else if (s.idf == 64) rc = dispatch_ga<64>(x, g_c, dX, p, st);
    if (rc == -1) {
        set_error("attn_bwd(tcgen05): unsupported shape idf=%d L=%d", s.idf, s.L);
        return SBA_ERR_UNSUPPORTED;
    }
    return rc;
}

}  // namespace sba

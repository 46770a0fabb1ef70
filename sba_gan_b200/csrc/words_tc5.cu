// Kernel (c) on the 5th-generation tensor cores - forward here, backward further down (phase A = this forward kernel with
// a template flag, phase B = k_words_bwd_tc5, then the d_img / d_words GEMMs k_words_dimg_tc5; formulas at each kernel
// and in DESIGN.md 4.4).  Forward: the B x B DAMSM region-word similarity of words_loss
// (AttnGAN2/code/miscc/losses.py:72-123 calling func_attention, GlobalAttention.py:31-69) as two chained
// 3xTF32 GEMMs per (image, block of captions) with the softmaxes in between, everything on chip:
//
//   G1  S[r, n]   = sum_c X_j[c, r] W[c, n]              M = regions (3 tiles of 128), N = 128 packed word columns, K = nef
//       a1        = softmax over the words of each caption (row-wise, per column segment)      (GlobalAttention.py:50-51)
//       e[r, n]   = exp(gamma1 (a1 - 1))                  = the un-normalised softmax over regions (:56-60; gamma1 a1 <= gamma1,
//                                                           so the shift by the constant gamma1 is the overflow guard)
//   G2  wc[n, c]  = sum_r e[r, n] X_j[c, r]               M = 128 word columns, N = nef, K = regions (chunks of 32)      (:67)
//       cos_n     = <w_n, wc_n> / max(|w_n| |wc_n| / Z_n, eps),  Z_n = sum_r e[r, n]            (losses.py:11-17)
//       sim[j, i] = gamma3 log sum_{n in caption i} exp(gamma2 cos_n)                            (losses.py:106-108, 123)
//
// A CUDA-core formulation of the same contraction runs at 24 % of the FFMA ceiling (words_loss.cu: bound by the
// shared-memory operand path); fp32 parity (1e-5 on a score amplified by gamma2 gamma3 = 50) rules out single-pass TF32,
// so every product is hi.hi + hi.lo + lo.hi with the operands split into tf32 hi / lo parts ONCE, by pre-pass kernels
// that also lay them out K-major for TMA: XT[r][c] for G1, X[c][r] (row stride padded to 16 bytes: 17 x 17 = 289 regions
// is not TMA-addressable as it lies) for G2, WT[n][c] with the captions' words packed into 64-column half blocks (no
// caption straddles one, so each epilogue warp owns whole captions).
//
// One CTA per (image j, 128-column block): TMA producer warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter, one
// per 64-column half).  S is double buffered in TMEM (2 x 128 columns), wc lives in the other 256; e goes through two
// 32 KB shared-memory chunk buffers ([128 n][32 regions], hi and lo) as the K-major A operand of G2.
#include "host_util.h"
#include "kernels.h"
#include "tc5_common.cuh"

namespace sba {
namespace {
using namespace tc5;

#ifdef SBA_DEV_AIDS
// development build: a barrier wait that times out records WHICH wait it was and carries on (garbage results, but the
// kernel ends and the flags can be read: sba_dev_words_timeouts) instead of trapping
} }  // (symbol at namespace sba scope for cudaMemcpyFromSymbol)
namespace sba { __device__ unsigned g_wt_dbg[16]; __device__ unsigned* g_wt_host = nullptr;
                __device__ unsigned long long g_wt_ns[8]; }
#define WT_NOW() ([] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }())
#define WT_ACC(i, dt) atomicAdd(&sba::g_wt_ns[i], (unsigned long long)(dt))
#define WT_MARK(i) do { if (sba::g_wt_host && (threadIdx.x & 31) == 0) atomicAdd(sba::g_wt_host + (i), 1u); } while (0)
namespace sba { namespace {
using namespace tc5;
__device__ __forceinline__ void wt_wait(uint32_t bar, uint32_t parity, int id) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (!done && ++spins > (1u << 14)) { atomicAdd(&g_wt_dbg[id], 1u); if (g_wt_host) atomicAdd(g_wt_host + id, 1u); return; }
    } while (!done);
}
#define WT_WAIT(bar, parity, id) wt_wait(bar, parity, id)
#else
#define WT_WAIT(bar, parity, id) mbar_wait(bar, parity)
#define WT_MARK(i) ((void)0)
#define WT_NOW() 0ull
#define WT_ACC(i, dt) ((void)0)
#endif

constexpr int kWtThreads = 64 + 256;      // producer warp, MMA warp, 8 epilogue warps
constexpr int kNB = 128;                  // word columns per CTA
constexpr int kHalf = 64;                 // columns per half block (one epilogue warp per lane quarter each)
constexpr int kKC = 32;                   // floats per 128-byte operand row = K per pipeline stage
constexpr int kTileBytes = 128 * kKC * 4; // one [128 rows][32 floats] operand tile: 16 KB
constexpr int kStageBytes = 4 * kTileBytes;       // G1: A hi, A lo, B hi, B lo;  G2: X hi, X lo ([256][32] each)
constexpr int kStages = 2;
constexpr int kEBufs = 2;
static_assert(kEBufs == 2, "the e_free barriers are indexed by k & 3 and waited on for chunk k - 2");
constexpr int kEBufBytes = 2 * kTileBytes;        // e hi, e lo
constexpr int kMaxChunks = 12;                    // regions / 32, R <= 384

struct WtPlan {          // device-resident packing of the captions' words into columns (written by k_wt_plan)
    int n_half;          // half blocks in use
};

struct WtParams {
    const float* wt_hi;        // [ncols][nef]  (only the final epilogue reads it directly; the MMAs go through TMA)
    const float* wt_lo;
    const float* ww;           // [ncols] |w_n|
    const int* col_cap;        // [ncols] caption of the column, -1 = padding
    const int* col_T;          // [ncols] length of the caption if this is its first column, else 0
    const WtPlan* plan;
    float* sim;                // [B_img][B_cap]   (forward)
    int B_cap, nef, R, MT, RKC;
    float g1l2e, g2, g3, eps;  // gamma1 * log2(e), gamma2, gamma3
    // backward, phase A (the forward again, with the upstream gradient): per (image, column) scalars and wc
    const float* d_sim;        // [B_img][B_cap]
    const int* cap_col;        // [B_cap] first column of each caption
    float4* scal;              // [B_img][ncols] (alpha = d num, beta = d|wc| / |wc|, D = sum_r a2 da2, 1 / Z)
    float* qt_hi;              // [B_img][ncols][nef]  q^T, q = alpha w + beta wc, in tf32 hi / lo: B operand of T = X^T q in phase B (written
                               //                      by TMA stores: the struct keeps the pointers for reference only)
    float* qt_lo;
    float* a1;                 // [B_img * ncols][RKP] the caption softmax a1[n][r] of every pair (phase B reads it back;
                               //                      regions contiguous: thread = region on both sides, coalesced)
    int RKP;
    float* kap;                // [B_img][ncols] d|w_n| * |w_n| share of this image (beta |wc|^2), or NULL (no word gradients)
    int ncols;
    float g1;
};

// ---- pre-pass 1: pack the captions into half blocks (greedy, in caption order) -----------------------------------------
// One block.  The lengths are staged in shared memory by all threads, one thread walks them (a few cycles per caption:
// no global round trip inside the sequential part), then all threads fill the per-column tables.
constexpr int kPlanMaxCaps = 4096;
constexpr int kMaxCapsPerHalf = 16;       // captions per half block (phase B keeps one total per caption and thread in shared memory)
__global__ void __launch_bounds__(256) k_wt_plan(const int* __restrict__ cap_lens, int B_cap, int Lw, int* __restrict__ cap_col,
                                                 WtPlan* plan, int* __restrict__ col_cap, int* __restrict__ col_T, int ncols) {
    __shared__ short s_T[kPlanMaxCaps];
    __shared__ int s_col[kPlanMaxCaps];
    for (int n = threadIdx.x; n < ncols; n += blockDim.x) { col_cap[n] = -1; col_T[n] = 0; }
    for (int i = threadIdx.x; i < B_cap; i += blockDim.x) {
        const int T = cap_lens[i];
        s_T[i] = (short)(T < 0 ? 0 : (T > Lw ? Lw : T));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int h = 0, off = 0, cnt = 0;
        for (int i = 0; i < B_cap; ++i) {
            const int T = s_T[i];
            if (off + T > kHalf || cnt == kMaxCapsPerHalf) { ++h; off = 0; cnt = 0; }
            s_col[i] = h * kHalf + off;
            off += T;
            ++cnt;
        }
        plan->n_half = h + 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B_cap; i += blockDim.x) {
        const int c0 = s_col[i], T = s_T[i];
        cap_col[i] = c0;
        if (T > 0) col_T[c0] = T;
        for (int t = 0; t < T; ++t) col_cap[c0 + t] = i;
    }
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = tf32_rna(v);
    lo = tf32_rna(v - hi);
}

// ---- pre-pass 2: WT hi / lo / unsplit [ncols][nef] and |w_n|; one block per 8 columns, thread = channel --------------------------
__global__ void __launch_bounds__(256) k_wt_words(const float* __restrict__ words, const int* __restrict__ col_cap,
                                                  const int* __restrict__ cap_col, float* __restrict__ wt_hi,
                                                  float* __restrict__ wt_lo, float* __restrict__ wt_f, float* __restrict__ ww,
                                                  float* __restrict__ wcp_hi, float* __restrict__ wcp_lo, int ncols, int nef,
                                                  int Lw) {
    __shared__ float red[8];
    for (int k = 0; k < 8; ++k) {
        const int n = blockIdx.x * 8 + k;
        const int i = col_cap[n];
        float sq = 0.f;
        for (int c = threadIdx.x; c < nef; c += 256) {
            float v = 0.f;
            if (i >= 0) v = __ldg(words + ((size_t)i * nef + c) * Lw + (n - cap_col[i]));
            float hi, lo;
            split_tf32(v, hi, lo);
            wt_hi[(size_t)n * nef + c] = hi;
            wt_lo[(size_t)n * nef + c] = lo;
            wt_f[(size_t)n * nef + c] = v;            // unsplit: the final epilogue's copy (cosine numerator, q)
            if (wcp_hi != nullptr) {                  // [nef][ncols]: B operand of the d_img GEMM's first term (backward only)
                wcp_hi[(size_t)c * ncols + n] = hi;
                wcp_lo[(size_t)c * ncols + n] = lo;
            }
            sq = fmaf(v, v, sq);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f;
            for (int w = 0; w < 8; ++w) a += red[w];
            ww[n] = sqrtf(a);
        }
    }
}

// ---- pre-pass 3: X hi / lo [B][nef][RKP] (regions contiguous, zero padded) and XT hi / lo [B][RMP][nef] -------------------
// block = (32 x 32 tile of (channel, region), image); through shared memory so that both layouts are written coalesced
__global__ void __launch_bounds__(256) k_wt_images(const float* __restrict__ img, float* __restrict__ x_hi,
                                                   float* __restrict__ x_lo, float* __restrict__ xt_hi,
                                                   float* __restrict__ xt_lo, int nef, int R, int RKP, int RMP) {
    __shared__ float t[32][33];
    const int j = blockIdx.z, c0 = blockIdx.y * 32, r0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows per pass
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        const float v = (c < nef && r < R) ? __ldg(img + ((size_t)j * nef + c) * R + r) : 0.f;
        t[ty + 8 * k][tx] = v;
        if (c < nef && r < RKP) {
            float hi, lo;
            split_tf32(v, hi, lo);
            x_hi[((size_t)j * nef + c) * RKP + r] = hi;
            x_lo[((size_t)j * nef + c) * RKP + r] = lo;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < RMP && c < nef) {
            float hi, lo;
            split_tf32(t[tx][ty + 8 * k], hi, lo);
            xt_hi[((size_t)j * RMP + r) * nef + c] = hi;
            xt_lo[((size_t)j * RMP + r) * nef + c] = lo;
        }
    }
}

// fp32 K-major operand tiles: rows of 32 floats (128 bytes), standard 128-byte swizzle; mn32 = true: [32 rows][32 floats]
// boxes in the 32-byte-atom flavour of the 128-byte swizzle, the layout of a 32-bit MN-major UMMA operand
int make_k128_map(CUtensorMap* out, const void* base, long long rows, int cols, int box_rows, bool mn32 = false);

// lane l ends with the sum over lanes of v[l] in v[0] (31 shuffles)
__device__ __forceinline__ void reduce_scatter32_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = upper ? v[k] : v[k + s];
            const float keep = upper ? v[k + s] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
}

__device__ __forceinline__ void warp_arrive1(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

template <bool BWD>
__global__ void __launch_bounds__(kWtThreads, 1)
    k_words_tc5(const __grid_constant__ CUtensorMap tm_xt_hi, const __grid_constant__ CUtensorMap tm_xt_lo,
                const __grid_constant__ CUtensorMap tm_wt_hi, const __grid_constant__ CUtensorMap tm_wt_lo,
                const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                const __grid_constant__ CUtensorMap tm_wf, const __grid_constant__ CUtensorMap tm_q_hi,
                const __grid_constant__ CUtensorMap tm_q_lo, const WtParams p) {
    const int nb = blockIdx.x, j = blockIdx.y;
    if (2 * nb >= p.plan->n_half) return;                      // the grid is sized for the worst packing
    [[maybe_unused]] const unsigned long long t_begin = WT_NOW();

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t s_ring = sbase;                                      // [kStages] operand stages
    const uint32_t s_e = s_ring + kStages * kStageBytes;                // [kEBufs] e chunks: hi, lo
    unsigned char* g_e = smem_raw + kStages * kStageBytes;
    float* zp = reinterpret_cast<float*>(g_e + kEBufs * kEBufBytes);    // [kMaxChunks][128] column sums of e per chunk
    float* exs = zp + kMaxChunks * kNB;                                 // [128] exp(gamma2 cos_n)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(exs + kNB);
    unsigned long long* bar_full = bars;                  // [kStages]
    unsigned long long* bar_empty = bars + kStages;       // [kStages]
    unsigned long long* bar_s_full = bars + 2 * kStages;  // [2]
    unsigned long long* bar_s_free = bar_s_full + 2;      // [2]
    unsigned long long* bar_e_ready = bar_s_free + 2;     // [kEBufs]
    // e_free is per chunk residue k & 3, not per buffer: the two warp pairs that share a buffer (q and q ^ 2) run
    // independently, so a waiter on a per-buffer barrier can be TWO completions behind and a parity wait would alias
    // (it passes on the stale phase and overwrites a chunk that has not been consumed).  Chunk k waits for chunk k - 2,
    // i.e. completion (k - 2) / 4 + 1 of bar_e_free[(k - 2) & 3], and saw completion (k - 2) / 4 one tile earlier.
    unsigned long long* bar_e_free = bar_e_ready + kEBufs;  // [4]
    unsigned long long* bar_d_full = bar_e_free + 4;        // [1]
    unsigned long long* bar_w = bar_d_full + 1;             // [1] this block's W^T rows are in the (retired) operand ring
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bar_w + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nef = p.nef, MT = p.MT, RKC = p.RKC, KCH = nef / kKC;

    if (tid == 0) {
        if (sbase & 1023u) __trap();
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_s_full[s]), 1); mbar_init(smem_u32(&bar_s_free[s]), 8); }
        for (int s = 0; s < kEBufs; ++s) mbar_init(smem_u32(&bar_e_ready[s]), 2);
        for (int s = 0; s < 4; ++s) mbar_init(smem_u32(&bar_e_free[s]), 1);
        mbar_init(smem_u32(bar_d_full), 1);
        mbar_init(smem_u32(bar_w), 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_s, 0);
    constexpr uint32_t COL_S = 0, COL_D = 256;

    if (warp == kProducerWarp) {
        // ------------------------------- TMA producer: the stages in the order the MMA warp consumes them ----------
        int s = 0;
        auto stage_g1 = [&](int m, int kc) {
            const int st = s % kStages;
            if (s >= kStages) WT_WAIT(smem_u32(&bar_empty[st]), (uint32_t)((s / kStages) - 1) & 1u, 9);
            if (elect_one()) {
                const uint32_t full = smem_u32(&bar_full[st]), dst = s_ring + st * kStageBytes;
                mbar_expect_tx(full, (uint32_t)kStageBytes);
                tma_load_2d(dst, &tm_xt_hi, kc * kKC, (j * MT + m) * 128, full);
                tma_load_2d(dst + kTileBytes, &tm_xt_lo, kc * kKC, (j * MT + m) * 128, full);
                tma_load_2d(dst + 2 * kTileBytes, &tm_wt_hi, kc * kKC, nb * kNB, full);
                tma_load_2d(dst + 3 * kTileBytes, &tm_wt_lo, kc * kKC, nb * kNB, full);
            }
            __syncwarp();
            ++s;
        };
        auto stage_g2 = [&](int k) {
            const int st = s % kStages;
            if (s >= kStages) WT_WAIT(smem_u32(&bar_empty[st]), (uint32_t)((s / kStages) - 1) & 1u, 9);
            if (elect_one()) {
                const uint32_t full = smem_u32(&bar_full[st]), dst = s_ring + st * kStageBytes;
                mbar_expect_tx(full, (uint32_t)(2 * nef * kKC * 4));
                tma_load_2d(dst, &tm_x_hi, k * kKC, j * nef, full);
                tma_load_2d(dst + 2 * kTileBytes, &tm_x_lo, k * kKC, j * nef, full);
            }
            __syncwarp();
            ++s;
        };
        for (int m = 0; m < MT; ++m) {
            for (int kc = 0; kc < KCH; ++kc) stage_g1(m, kc);
            if (m >= 1)
                for (int q = 0; q < 4; ++q)
                    if (4 * (m - 1) + q < RKC) stage_g2(4 * (m - 1) + q);
        }
        for (int q = 0; q < 4; ++q)
            if (4 * (MT - 1) + q < RKC) stage_g2(4 * (MT - 1) + q);
        // The final epilogue wants this block's 128 rows of W^T (thread = column, all nef channels): read per thread from
        // global memory that is a 16-byte piece of a different line per lane on every load (measured: 11 of a CTA's 40 us
        // in the forward).  The operand ring is idle once its last stages have been consumed: KCH tiles of [128 rows][32
        // channels] land there instead (nef <= 256: at most 8 x 16 KB = the ring), the same swizzled rows the MMAs read.
        for (int st = 0; st < kStages; ++st) {
            const int uses = (s - st + kStages - 1) / kStages;               // stages s' < s with s' % kStages == st
            if (uses > 0) WT_WAIT(smem_u32(&bar_empty[st]), (uint32_t)(uses - 1) & 1u, 9);
        }
        if (elect_one()) {
            const uint32_t full = smem_u32(bar_w);
            mbar_expect_tx(full, (uint32_t)(KCH * kTileBytes));
            for (int kc = 0; kc < KCH; ++kc) tma_load_2d(s_ring + kc * kTileBytes, &tm_wf, kc * kKC, nb * kNB, full);
        }
        __syncwarp();
    } else if (warp == kMmaWarp) {
        // ------------------------------- MMA issuer ------------------------------------------------------------------
        constexpr uint32_t kHi = desc_hi(1024, kSwizzle128B);
        const uint32_t idesc1 = make_idesc(2, 0, 0, 128, kNB);
        const uint32_t idesc2 = make_idesc(2, 0, 0, 128, nef);
        int s = 0;
        bool d_started = false;
        auto g1 = [&](int m) {
            const int buf = m & 1;
            if (m >= 2) WT_WAIT(smem_u32(&bar_s_free[buf]), (uint32_t)((m >> 1) - 1) & 1u, 10);
            const uint32_t d = tmem_base + COL_S + 128 * buf;
            for (int kc = 0; kc < KCH; ++kc, ++s) {
                const int st = s % kStages;
                WT_WAIT(smem_u32(&bar_full[st]), (uint32_t)(s / kStages) & 1u, 11);
                tc_fence_after();
                const uint32_t base = s_ring + st * kStageBytes;
                const uint32_t a_hi = desc_lo(base, 16), a_lo = desc_lo(base + kTileBytes, 16);
                const uint32_t b_hi = desc_lo(base + 2 * kTileBytes, 16), b_lo = desc_lo(base + 3 * kTileBytes, 16);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint32_t o = (uint32_t)(ks * 2);            // 32 bytes per k-step, in 16-byte units
                        umma_ss<true>(d, a_hi + o, kHi, b_hi + o, kHi, idesc1, (kc > 0 || ks > 0) ? 1u : 0u);
                        umma_ss<true>(d, a_hi + o, kHi, b_lo + o, kHi, idesc1, 1u);
                        umma_ss<true>(d, a_lo + o, kHi, b_hi + o, kHi, idesc1, 1u);
                    }
                    umma_commit(smem_u32(&bar_empty[st]));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&bar_s_full[buf]));
            __syncwarp();
        };
        auto g2 = [&](int m) {
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * m + q;
                if (k >= RKC) break;
                const int eb = k % kEBufs, st = s % kStages;
                WT_WAIT(smem_u32(&bar_e_ready[eb]), (uint32_t)(k / kEBufs) & 1u, 12);
                WT_WAIT(smem_u32(&bar_full[st]), (uint32_t)(s / kStages) & 1u, 11);
                tc_fence_after();
                const uint32_t eh = desc_lo(s_e + eb * kEBufBytes, 16), el = desc_lo(s_e + eb * kEBufBytes + kTileBytes, 16);
                const uint32_t xb = s_ring + st * kStageBytes;
                const uint32_t xh = desc_lo(xb, 16), xl = desc_lo(xb + 2 * kTileBytes, 16);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint32_t o = (uint32_t)(ks * 2);
                        umma_ss<true>(tmem_base + COL_D, eh + o, kHi, xh + o, kHi, idesc2, (d_started || ks > 0) ? 1u : 0u);
                        umma_ss<true>(tmem_base + COL_D, eh + o, kHi, xl + o, kHi, idesc2, 1u);
                        umma_ss<true>(tmem_base + COL_D, el + o, kHi, xh + o, kHi, idesc2, 1u);
                    }
                    umma_commit(smem_u32(&bar_empty[st]));
                    umma_commit(smem_u32(&bar_e_free[k & 3]));
                }
                __syncwarp();
                d_started = true;
                ++s;
            }
        };
        for (int m = 0; m < MT; ++m) {
            g1(m);
            if (m >= 1) g2(m - 1);
        }
        g2(MT - 1);
        if (elect_one()) umma_commit(smem_u32(bar_d_full));
        __syncwarp();
    } else {
        // ------------------------------- epilogue warps ----------------------------------------------------------------
        const int ew = warp - kFirstConsumerWarp;          // 0..7
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int hf = ew >> 2;                            // 64-column half block
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        const int ncol0 = nb * kNB + hf * kHalf;           // first global column of this warp's half block
        // segment structure of the half block (block-uniform per warp): bit c of `valid` = column is a caption word,
        // `first` = first word of its caption, `last` = last word
        unsigned long long valid = 0ull, first = 0ull;
        {
            const int c0v = p.col_cap[ncol0 + lane], c1v = p.col_cap[ncol0 + 32 + lane];
            const int t0v = p.col_T[ncol0 + lane], t1v = p.col_T[ncol0 + 32 + lane];
            valid = (unsigned long long)__ballot_sync(0xffffffffu, c0v >= 0) |
                    ((unsigned long long)__ballot_sync(0xffffffffu, c1v >= 0) << 32);
            first = (unsigned long long)__ballot_sync(0xffffffffu, t0v > 0) |
                    ((unsigned long long)__ballot_sync(0xffffffffu, t1v > 0) << 32);
        }
        const unsigned long long last = valid & ((first >> 1) | ~(valid >> 1));
        constexpr float kLog2e = 1.4426950408889634f;

        for (int m = 0; m < MT; ++m) {
            const int k = 4 * m + q;                       // region chunk this warp produces
            const bool exists = k < RKC;
            const int buf = m & 1;
            WT_WAIT(smem_u32(&bar_s_full[buf]), (uint32_t)(m >> 1) & 1u, 13);
            tc_fence_after();
            uint32_t sr[kHalf];
            if (exists) {
                tmem_ld<kHalf>(tl + COL_S + 128 * buf + hf * kHalf, sr);
                tmem_wait_ld();
            }
            tc_fence_before();
            warp_arrive1(smem_u32(&bar_s_free[buf]), lane);
            if (!exists) continue;
            const bool rv = (128 * m + 32 * q + lane) < p.R;
            float v[kHalf];
            // softmax over the words of each caption: forward running max, backward segment max, exp, forward running
            // sum, backward segment sum - fully unrolled, the segment structure is in uniform bit masks
            float run = -INFINITY;
#pragma unroll
            for (int c = 0; c < kHalf; ++c) {
                const float x = __uint_as_float(sr[c]);
                run = ((first >> c) & 1ull) ? x : fmaxf(run, x);
                v[c] = x;
                sr[c] = __float_as_uint(run);
            }
            float seg = 0.f;
#pragma unroll
            for (int c = kHalf - 1; c >= 0; --c) {
                seg = ((last >> c) & 1ull) ? __uint_as_float(sr[c]) : seg;
                v[c] = mma::ex2_approx((v[c] - seg) * kLog2e);
            }
            run = 0.f;
#pragma unroll
            for (int c = 0; c < kHalf; ++c) {
                run = ((first >> c) & 1ull) ? v[c] : run + v[c];
                sr[c] = __float_as_uint(run);
            }
            seg = 1.f;
            [[maybe_unused]] float* a1o = p.a1 + ((size_t)j * p.ncols + ncol0) * p.RKP + 128 * m + 32 * q + lane;
#pragma unroll
            for (int c = kHalf - 1; c >= 0; --c) {
                seg = ((last >> c) & 1ull) ? mma::rcp_approx(__uint_as_float(sr[c])) : seg;
                const float a1 = v[c] * seg;
                // e = exp(gamma1 (a1 - 1)); padding columns and regions beyond R contribute nothing
                v[c] = (rv && ((valid >> c) & 1ull)) ? mma::ex2_approx((a1 - 1.f) * p.g1l2e) : 0.f;
                // phase A of the backward: a1 goes out (phase B does not recompute it), one coalesced line per column
                if constexpr (BWD) a1o[(size_t)c * p.RKP] = a1;
            }
            // column sums over this warp's 32 regions -> zp[k][column] (added in chunk order by the final epilogue)
            {
                float t[32];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) t[c] = v[32 * h2 + c];
                    reduce_scatter32_sum(t, lane);
                    zp[k * kNB + hf * kHalf + 32 * h2 + lane] = t[0];
                }
            }
            // e hi / lo -> chunk buffer rows n = hf * 64 + c, 32 regions per 128-byte row, 16-byte chunks XOR-swizzled with n & 7
            const int eb = k % kEBufs;
            if (k >= kEBufs) WT_WAIT(smem_u32(&bar_e_free[(k - kEBufs) & 3]), (uint32_t)((k - kEBufs) >> 2) & 1u, 14);
            unsigned char* eh = g_e + eb * kEBufBytes;
#pragma unroll
            for (int c = 0; c < kHalf; ++c) {
                const int n = hf * kHalf + c;
                const uint32_t off = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((((lane >> 2) ^ (n & 7)) << 4) | ((lane & 3) << 2)));
                const float hi = tf32_rna(v[c]);
                *reinterpret_cast<float*>(eh + off) = hi;
                *reinterpret_cast<float*>(eh + kTileBytes + off) = tf32_rna(v[c] - hi);
            }
            fence_proxy_async();
            warp_arrive1(smem_u32(&bar_e_ready[eb]), lane);
        }

        // ---- final epilogue: thread = word column n (the four warps whose lanes cover 0..127) ---------------------------
        [[maybe_unused]] unsigned long long t_loop = 0, t_dfull = 0;
        if (ew < 4) {
            const int n = 32 * q + lane, ng = nb * kNB + n;
            t_loop = WT_NOW();
            WT_WAIT(smem_u32(bar_d_full), 0u, 15);
            t_dfull = WT_NOW();
            tc_fence_after();
            float num = 0.f, wn2 = 0.f;
            // row n of the W^T tiles in the ring: 128-byte rows in 8-row atoms, 16-byte chunks XOR-swizzled with n & 7
            const unsigned char* wrow = smem_raw + (n >> 3) * 1024 + (n & 7) * 128;
            WT_WAIT(smem_u32(bar_w), 0u, 8);
            for (int c0 = 0; c0 < nef; c0 += 32) {
                uint32_t d[32];
                tmem_ld<32>(tl + COL_D + c0, d);
                tmem_wait_ld();
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 a = *reinterpret_cast<const float4*>(wrow + (c0 >> 5) * kTileBytes + ((c4 ^ (n & 7)) << 4));
                    const float w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float wc = __uint_as_float(d[4 * c4 + e]);
                        num = fmaf(w[e], wc, num);
                        wn2 = fmaf(wc, wc, wn2);
                    }
                }
            }
            tc_fence_before();
            float Z = 0.f;
            mma::named_bar_sync(2, 128);                   // every zp entry of this CTA has been written (all chunks are done:
            for (int k = 0; k < RKC; ++k) Z += zp[k * kNB + n];     // d_full follows the last G2, which follows e_ready)
            const int cap = p.col_cap[ng];
            const bool cv = cap >= 0;
            float ex = 0.f, invZ = 0.f, den = 1.f, wn = 0.f;
            if (cv) {
                invZ = 1.0f / Z;
                wn = sqrtf(wn2) * invZ;                                              // |wc_n|
                den = fmaxf(p.ww[ng] * wn, p.eps);                                   // losses.py:17
                ex = expf(p.g2 * (num * invZ) / den);                                // losses.py:106
            }
            exs[n] = ex;
            mma::named_bar_sync(2, 128);
            const int T = p.col_T[ng];
            if constexpr (!BWD) {
                if (T > 0) {
                    float E = 0.f;
                    for (int t = 0; t < T; ++t) E += exs[n + t];
                    p.sim[(size_t)j * p.B_cap + cap] = p.g3 * logf(E);              // losses.py:107-108, 123
                }
            } else {
                // E of this column's caption: summed by the caption's first column, looked up by the others
                float* Es = zp;                                  // (zp has been consumed: reuse its first 128 words)
                mma::named_bar_sync(2, 128);
                if (T > 0) {
                    float E = 0.f;
                    for (int t = 0; t < T; ++t) E += exs[n + t];
                    Es[n] = E;
                }
                mma::named_bar_sync(2, 128);
                float alpha = 0.f, beta = 0.f, D = 0.f;
                if (cv) {
                    const float E = Es[p.cap_col[cap] - nb * kNB];
                    const float g = p.d_sim[(size_t)j * p.B_cap + cap];
                    const float numt = num * invZ;                                   // <w_n, wc_n>
                    const float gcos = g * p.g3 / E * p.g2 * ex;
                    const float prod = p.ww[ng] * wn;
                    alpha = gcos / den;                                              // d num
                    const float d_den = prod > p.eps ? -gcos * numt / (den * den) : 0.f;
                    beta = wn > 0.f ? d_den * p.ww[ng] / wn : 0.f;                   // d|wc| / |wc|
                    D = alpha * numt + beta * wn * wn;                               // sum_r a2[r, n] da2[r, n]
                }
                p.scal[(size_t)j * p.ncols + ng] = make_float4(alpha, beta, D, invZ);
                if (p.kap != nullptr) p.kap[(size_t)j * p.ncols + ng] = beta * wn * wn;      // = d_den |w| |wc| (losses.py:17)
                // second pass over wc: q^T = alpha w + beta wc split into tf32 hi / lo (this thread's row, 128 contiguous
                // bytes per chunk): <X_r, q_n> = alpha S[r, n] + beta V[r, n] is the whole d a2 that phase B needs - one GEMM
                // instead of two - and d_img = sum_n dS[r, n] w_n + sum_n a2[n, r] q_n needs no other per-image operand
                // q^T rows go out through the (retired) e chunk buffers: [128 columns][32 channels] tiles, hi and lo, in
                // the swizzled layout of a TMA box - one tensor store per tile instead of a 16-byte piece of a different
                // line per lane on every store; two tiles pairs alternate
                const uint32_t qrow = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
                const bool issuer = ew == 0 && lane == 0;
                __syncwarp();                    // (lane-dependent code above; tcgen05.ld is warp-collective)
                tc_fence_after();
                for (int c0 = 0; c0 < nef; c0 += 32) {
                    const int kq = c0 >> 5;
                    unsigned char* qs = g_e + (kq & 1) * kEBufBytes + qrow;           // hi tile; lo tile kTileBytes further
                    if (kq >= 2) {                                                    // the store that last read this pair
                        if (issuer) bulk_wait_read<1>();
                        mma::named_bar_sync(2, 128);
                    }
                    uint32_t d[32];
                    tmem_ld<32>(tl + COL_D + c0, d);
                    tmem_wait_ld();
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float hi[4], lo[4];
                        const float4 a = *reinterpret_cast<const float4*>(wrow + (c0 >> 5) * kTileBytes + ((c4 ^ (n & 7)) << 4));
                        const float w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float wc = __uint_as_float(d[4 * c4 + e]) * invZ;
                            const float qv = fmaf(alpha, w[e], beta * wc);
                            hi[e] = tf32_rna(qv);
                            lo[e] = tf32_rna(qv - hi[e]);
                        }
                        *reinterpret_cast<float4*>(qs + ((c4 ^ (n & 7)) << 4)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4*>(qs + kTileBytes + ((c4 ^ (n & 7)) << 4)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    fence_proxy_async();
                    mma::named_bar_sync(2, 128);
                    if (issuer) {
                        const uint32_t src = s_e + (kq & 1) * kEBufBytes;
                        tma_store_2d(&tm_q_hi, c0, j * p.ncols + nb * kNB, src);
                        tma_store_2d(&tm_q_lo, c0, j * p.ncols + nb * kNB, src + kTileBytes);
                        bulk_commit();
                    }
                }
                if (issuer) bulk_wait<0>();          // the stores have left shared memory and landed before the CTA retires
            }
        }
        if (ew == 2 && lane == 0) {          // (q = 0 warp of the final epilogue: its tile loop is the longest)
            const unsigned long long t_end = WT_NOW();
            WT_ACC(BWD ? 4 : 0, 1);
            WT_ACC(BWD ? 5 : 1, t_loop - t_begin);
            WT_ACC(BWD ? 6 : 2, t_dfull - t_loop);
            WT_ACC(BWD ? 7 : 3, t_end - t_dfull);
        }
        WT_MARK(22);
        tc_fence_before();
    }
    WT_MARK(23);
    __syncthreads();
    WT_MARK(24);
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward, phase B: per (image j, 128-column block) ONE GEMM  T = X^T q,  q_n = alpha_n w_n + beta_n wc_n  (phase A wrote
// q^T in tf32 hi / lo), so that  T[r, n] = alpha S[r, n] + beta V[r, n] = d a2[n, r]  without recomputing S or forming V;
// then, thread = region row, with a1[r][n] read back from phase A,
//   a2 = exp(gamma1 (a1 - 1)) / Z,   dz = a2 (T - D),   t = a1 gamma1 dz,   ds = t - a1 sum_{caption} t,   u = ds + alpha a2
// (D_n = sum_r a2 da2 = alpha <w_n, wc_n> + beta |wc_n|^2 is known from phase A: no reduction over regions here) and
// store a2^T[j][r][n], u^T[j][r][n] (tf32 hi / lo) for the d_img GEMM below (oracle/attention.py::words_loss_backward;
// the CUDA-core kernel of words_loss.cu materialises the same u / a2).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBStageBytes = 4 * kTileBytes;      // A hi, A lo, q hi, q lo: 64 KB
constexpr int kBStages = 2;
constexpr int kBStageF4 = 2 * 32 * 8;             // float4 per epilogue warp: 32 columns of 32 regions, hi and lo (8 KB)
struct WtBwdParams {
    const int* col_cap;
    const WtPlan* plan;
    const float4* scal;        // [B_img][ncols]  (alpha, beta, D, 1 / Z)
    const float* a1;           // [B_img * ncols][RKP]
    float* u_hi;               // [B_img * R][ncols]  u^T and a2^T (regions = rows, columns contiguous), tf32 hi / lo:
    float* u_lo;               //                     the K-major A operands of the d_img GEMM
    float* a2_hi;
    float* a2_lo;
    float* u2_hi;              // [B_img * ncols][RKP]  u again, columns = rows, regions contiguous (zero for r >= R):
    float* u2_lo;              //                       K-major A operand of the d_words GEMM; NULL = no word gradients
    int RKP;
    int nef, R, MT, RKC, ncols;
    float g1, g1l2e;
};

template <bool WORDS>
__global__ void __launch_bounds__(kWtThreads, 1)
    k_words_bwd_tc5(const __grid_constant__ CUtensorMap tm_xt_hi, const __grid_constant__ CUtensorMap tm_xt_lo,
                    const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo, const WtBwdParams p) {
    const int nb = blockIdx.x, j = blockIdx.y;
    if (2 * nb >= p.plan->n_half) return;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t s_ring = sbase;
    float4* scs = reinterpret_cast<float4*>(smem_raw + kBStages * kBStageBytes);        // [128] per-column scalars
    float* tot_all = reinterpret_cast<float*>(scs + kNB);                               // [16 captions][256 epilogue threads]
    float4* stg_all = reinterpret_cast<float4*>(tot_all + kMaxCapsPerHalf * 256);       // [8 warps][hi, lo][32 regions][8 x float4]
    float* stg2_all = reinterpret_cast<float*>(stg_all + 8 * kBStageF4);                // WORDS: [8 warps][hi, lo][8 columns][32 regions]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stg2_all + (WORDS ? 8 * 2 * 8 * 32 : 0));
    unsigned long long* bar_full = bars;                   // [kBStages]
    unsigned long long* bar_empty = bars + kBStages;       // [kBStages]
    unsigned long long* bar_t_full = bars + 2 * kBStages;  // [2]  T of a region tile is complete
    unsigned long long* bar_t_free = bar_t_full + 2;       // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bar_t_free + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nef = p.nef, MT = p.MT, RKC = p.RKC, KCH = nef / kKC;

    if (tid == 0) {
#ifdef SBA_DEV_AIDS
        if (sbase & 1023u) atomicAdd(&g_wt_dbg[8], 1u);
#else
        if (sbase & 1023u) __trap();
#endif
        for (int s = 0; s < kBStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_t_full[s]), 1); mbar_init(smem_u32(&bar_t_free[s]), 8); }
        fence_barrier_init();
    }
    if (tid >= 64 && tid < 64 + kNB) scs[tid - 64] = p.scal[(size_t)j * p.ncols + nb * kNB + tid - 64];
    if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_base_s), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_s, 0);
    WT_MARK(16);

    if (warp == kProducerWarp) {
        int s = 0;
        for (int m = 0; m < MT; ++m)
            for (int kc = 0; kc < KCH; ++kc, ++s) {
                const int st = s % kBStages;
                if (s >= kBStages) WT_WAIT(smem_u32(&bar_empty[st]), (uint32_t)((s / kBStages) - 1) & 1u, 0);
                if (elect_one()) {
                    const uint32_t full = smem_u32(&bar_full[st]), dst = s_ring + st * kBStageBytes;
                    mbar_expect_tx(full, (uint32_t)kBStageBytes);
                    tma_load_2d(dst, &tm_xt_hi, kc * kKC, (j * MT + m) * 128, full);
                    tma_load_2d(dst + kTileBytes, &tm_xt_lo, kc * kKC, (j * MT + m) * 128, full);
                    tma_load_2d(dst + 2 * kTileBytes, &tm_q_hi, kc * kKC, j * p.ncols + nb * kNB, full);
                    tma_load_2d(dst + 3 * kTileBytes, &tm_q_lo, kc * kKC, j * p.ncols + nb * kNB, full);
                }
                __syncwarp();
            }
        WT_MARK(17);
    } else if (warp == kMmaWarp) {
        constexpr uint32_t kHi = desc_hi(1024, kSwizzle128B);
        const uint32_t idesc = make_idesc(2, 0, 0, 128, kNB);
        int s = 0;
        for (int m = 0; m < MT; ++m) {
            const int buf = m & 1;
            if (m >= 2) WT_WAIT(smem_u32(&bar_t_free[buf]), (uint32_t)((m >> 1) - 1) & 1u, 1);
            const uint32_t dT = tmem_base + 128 * buf;
            for (int kc = 0; kc < KCH; ++kc, ++s) {
                const int st = s % kBStages;
                WT_WAIT(smem_u32(&bar_full[st]), (uint32_t)(s / kBStages) & 1u, 2);
                tc_fence_after();
                const uint32_t base = s_ring + st * kBStageBytes;
                const uint32_t a_hi = desc_lo(base, 16), a_lo = desc_lo(base + kTileBytes, 16);
                const uint32_t q_hi = desc_lo(base + 2 * kTileBytes, 16), q_lo = desc_lo(base + 3 * kTileBytes, 16);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint32_t o = (uint32_t)(ks * 2);
                        umma_ss<true>(dT, a_hi + o, kHi, q_hi + o, kHi, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        umma_ss<true>(dT, a_hi + o, kHi, q_lo + o, kHi, idesc, 1u);
                        umma_ss<true>(dT, a_lo + o, kHi, q_hi + o, kHi, idesc, 1u);
                    }
                    umma_commit(smem_u32(&bar_empty[st]));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&bar_t_full[buf]));
            __syncwarp();
            WT_MARK(18);
        }
    } else {
        const int ew = warp - kFirstConsumerWarp, q = warp & 3, hf = ew >> 2;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        const int ncol0 = nb * kNB + hf * kHalf;
        unsigned long long valid = 0ull, first = 0ull;
        {
            const int c0v = p.col_cap[ncol0 + lane], c1v = p.col_cap[ncol0 + 32 + lane];
            // (every lane takes part in every shuffle: none of them inside a short-circuit expression)
            const int prev0 = __shfl_up_sync(0xffffffffu, c0v, 1), prev1 = __shfl_up_sync(0xffffffffu, c1v, 1);
            const int last0 = __shfl_sync(0xffffffffu, c0v, 31);
            const bool f0 = c0v >= 0 && (lane == 0 || prev0 != c0v);
            const bool f1 = c1v >= 0 && ((lane == 0 ? last0 : prev1) != c1v);
            valid = (unsigned long long)__ballot_sync(0xffffffffu, c0v >= 0) |
                    ((unsigned long long)__ballot_sync(0xffffffffu, c1v >= 0) << 32);
            first = (unsigned long long)__ballot_sync(0xffffffffu, f0) | ((unsigned long long)__ballot_sync(0xffffffffu, f1) << 32);
        }
        const unsigned long long last = valid & ((first >> 1) | ~(valid >> 1));

        for (int m = 0; m < MT; ++m) {
            const bool exists = 4 * m + q < RKC;
            const int buf = m & 1;
            const int r = 128 * m + 32 * q + lane;
            const bool rv = r < p.R;
            // Both passes walk the 64 columns forward: a1 (written by phase A) is read from global memory once per tile,
            // all 16 loads in flight before the tile's T is waited for; T comes from tensor memory in 16-column pieces
            // in each pass, and the only state carried from the first pass to the second is one total per caption, in
            // shared memory.  (A version that also held 64 running sums in registers spilled 500-900 bytes per thread, and
            // with 194 KB of the SM's memory carved out for the operand ring that local memory lives in L2:
            // long-scoreboard stalls were 42 % of the kernel's samples.)
            const float* a1p = p.a1 + ((size_t)j * p.ncols + ncol0) * p.RKP + r;      // (every existing warp's r < RKP)
            float a1[kHalf];
            if (exists) {
#pragma unroll
                for (int c = 0; c < kHalf; ++c) a1[c] = __ldcs(a1p + (size_t)c * p.RKP);
            }
            WT_WAIT(smem_u32(&bar_t_full[buf]), (uint32_t)(m >> 1) & 1u, 3);
            tc_fence_after();
            if (!exists) {
                tc_fence_before();
                warp_arrive1(smem_u32(&bar_t_free[buf]), lane);
                continue;
            }
            const uint32_t tT = tl + 128 * buf + hf * kHalf;
            float* tot = tot_all + ew * 32 + lane;               // [caption ordinal * 256]
            // u^T / a2^T rows [region][column]: thread = region, so a thread's own row is a 16-byte piece of a different line
            // per lane on every store.  32 columns of the warp's 32 regions are staged (float4 slots XOR-swizzled with
            // region & 7: conflict-free both ways) and leave as whole 128-byte row pieces, four rows per store instruction.
            float4* stg = stg_all + ew * kBStageF4;
            const int r_base = 128 * m + 32 * q;
            const size_t obase = ((size_t)j * p.R + r_base) * p.ncols + ncol0;
            auto stage4 = [&](int c, const float (&h)[4], const float (&l)[4]) {      // columns c - 3 .. c of this lane's region
                const int slot = lane * 8 + ((((c & 31) >> 2)) ^ (lane & 7));
                stg[slot] = make_float4(h[0], h[1], h[2], h[3]);
                stg[256 + slot] = make_float4(l[0], l[1], l[2], l[3]);
            };
            auto flush32 = [&](float* __restrict__ dst_hi, float* __restrict__ dst_lo, int col0) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = 4 * i + (lane >> 3), k4 = lane & 7;
                    if (r_base + row < p.R) {
                        const int slot = row * 8 + (k4 ^ (row & 7));
                        const size_t o = obase + (size_t)row * p.ncols + col0 + 4 * k4;
                        *reinterpret_cast<float4*>(dst_hi + o) = stg[slot];
                        *reinterpret_cast<float4*>(dst_lo + o) = stg[256 + slot];
                    }
                }
                __syncwarp();                                    // read out before the next 32 columns are staged
            };
            // (1) t = a1 gamma1 a2 (T - D) summed per caption -> tot; a2 goes out
            {
                float run = 0.f;
                int ord = 0;
#pragma unroll
                for (int g0 = 0; g0 < kHalf; g0 += 16) {
                    uint32_t t16[16];
                    float ah[4], al[4];
                    __syncwarp();                                // (the stores below are lane-dependent)
                    tmem_ld<16>(tT + g0, t16);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int c = g0 + k;
                        const float4 cs4 = scs[hf * kHalf + c];  // alpha, beta, D, 1 / Z
                        const bool ok = rv && ((valid >> c) & 1ull);
                        const float a2 = ok ? mma::ex2_approx((a1[c] - 1.f) * p.g1l2e) * cs4.w : 0.f;
                        const float t = a1[c] * p.g1 * a2 * (__uint_as_float(t16[k]) - cs4.z);
                        run = ((first >> c) & 1ull) ? t : run + t;
                        if ((last >> c) & 1ull) {                // (block-uniform)
                            tot[ord * 256] = run;
                            ++ord;
                        }
                        split_tf32(a2, ah[k & 3], al[k & 3]);
                        if ((k & 3) == 3) stage4(c, ah, al);
                    }
                    if (g0 & 16) flush32(p.a2_hi, p.a2_lo, g0 - 16);
                }
            }
            // (2) ds = t - a1 (sum of t over the caption), u = ds + alpha a2: the same walk again
            {
                float uh[4], ul[4];
                float seg = 0.f;
                int ord = 0;
#pragma unroll
                for (int g0 = 0; g0 < kHalf; g0 += 16) {
                    uint32_t t16[16];
                    __syncwarp();
                    tmem_ld<16>(tT + g0, t16);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int c = g0 + k;
                        const float4 cs4 = scs[hf * kHalf + c];
                        const bool ok = rv && ((valid >> c) & 1ull);
                        if ((first >> c) & 1ull) {               // (block-uniform)
                            seg = tot[ord * 256];
                            ++ord;
                        }
                        const float a2 = ok ? mma::ex2_approx((a1[c] - 1.f) * p.g1l2e) * cs4.w : 0.f;
                        const float t = a1[c] * p.g1 * a2 * (__uint_as_float(t16[k]) - cs4.z);
                        const float dsv = ok ? t - a1[c] * seg : 0.f;                // d S[r, n]: the d_img GEMM's first term
                        split_tf32(dsv, uh[k & 3], ul[k & 3]);
                        if constexpr (WORDS) {
                            float u2h, u2l;                                          // u = ds + alpha a2: the d_words GEMM's operand
                            split_tf32(ok ? dsv + cs4.x * a2 : 0.f, u2h, u2l);
                            // u again as [column][region] (the d_words GEMM's K-major operand): 8 columns of the warp's 32
                            // regions are staged and leave as 16-byte pieces along the regions - two store instructions per
                            // array instead of eight (every existing warp's regions are < RKP; zero beyond R)
                            float* s2 = stg2_all + ew * 512;
                            s2[(k & 7) * 32 + lane] = u2h;
                            s2[256 + (k & 7) * 32 + lane] = u2l;
                            if ((k & 7) == 7) {
                                __syncwarp();
#pragma unroll
                                for (int i = 0; i < 2; ++i) {
                                    const int cc = 4 * i + (lane >> 3), r4 = (lane & 7) * 4;
                                    const size_t o2 = ((size_t)j * p.ncols + ncol0 + c - 7 + cc) * p.RKP + r_base + r4;
                                    *reinterpret_cast<float4*>(p.u2_hi + o2) = *reinterpret_cast<const float4*>(s2 + cc * 32 + r4);
                                    *reinterpret_cast<float4*>(p.u2_lo + o2) = *reinterpret_cast<const float4*>(s2 + 256 + cc * 32 + r4);
                                }
                                __syncwarp();
                            }
                        }
                        if ((k & 3) == 3) stage4(c, uh, ul);
                    }
                    if (g0 & 16) flush32(p.u_hi, p.u_lo, g0 - 16);
                }
            }
            tc_fence_before();
            warp_arrive1(smem_u32(&bar_t_free[buf]), lane);
            __syncwarp();
            WT_MARK(19);
        }
        tc_fence_before();
        WT_MARK(20);
    }
    __syncthreads();
    WT_MARK(21);
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward, d_img: per (image j, 128-region tile m)
//   d_img[j][c][r] = sum_n u[j][n][r] w[c][n] + sum_n a2[j][n][r] v[j][c][n]
// as ONE accumulation D[r][c] over the concatenated K = (columns of term 1 | columns of term 2): A = u^T / a2^T tiles
// [128 regions][32 columns], B = wcp / v[j] tiles [nef channels][32 columns], all K-major 128-byte swizzled fp32, 3xTF32.
// D (128 lanes x nef columns) stays in TMEM for the whole K loop (2 * 64 * n_half / 32 stages); the epilogue's thread =
// region, so every store instruction of a warp writes 32 consecutive regions of one channel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGThreads = 64 + 128;                       // producer warp, MMA warp, 4 epilogue warps
constexpr int kGBTileBytes = 256 * kKC * 4;               // B operand tile: up to 256 channels
constexpr int kGStageBytes = 2 * kTileBytes + 2 * kGBTileBytes;      // A hi, A lo, B hi, B lo: 96 KB
struct WtGemmParams {
    const WtPlan* plan;
    float* d_img;              // [B_img][nef][R]
    int nef, R;
    // d_words mode: part[split][n][c] = sum over the split's images j and all regions r of u[j][n][r] X_j[c][r]
    float* part;               // [splits][ncols][nef]
    int ncols, B_img, RKC;
};

// WORDS = false: the d_img GEMM above, CTA = (region tile, image).
// WORDS = true : the word-gradient GEMM  part[n][c] = sum_{j in split} sum_r u[j][n][r] X_j[c][r]  (M = 128 word columns,
//                N = nef, K = (image, region chunk)), CTA = (128-column block, split of the images); A = u2 hi / lo (maps 0, 1),
//                B = X hi / lo (maps 4, 5: the forward's G2 operand).  The splits are added in order by k_wt_dwords.
template <bool WORDS>
__global__ void __launch_bounds__(kGThreads, 1)
    k_words_dimg_tc5(const __grid_constant__ CUtensorMap tm_u_hi, const __grid_constant__ CUtensorMap tm_u_lo,
                     const __grid_constant__ CUtensorMap tm_a2_hi, const __grid_constant__ CUtensorMap tm_a2_lo,
                     const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                     const __grid_constant__ CUtensorMap tm_qmn_hi, const __grid_constant__ CUtensorMap tm_qmn_lo,
                     const WtGemmParams p) {
    const int m = blockIdx.x, j = blockIdx.y;           // WORDS: m = column block, j = split
    if (WORDS && 2 * m >= p.plan->n_half) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kStages * kGStageBytes);
    unsigned long long* bar_full = bars;                  // [kStages]
    unsigned long long* bar_empty = bars + kStages;       // [kStages]
    unsigned long long* bar_d_full = bars + 2 * kStages;  // [1]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bar_d_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nef = p.nef;
    const int KCH = p.plan->n_half * (kHalf / kKC);       // d_img: stages per term
    // d_words: images [j0, j1) of this split, RKC region chunks each
    const int j0 = WORDS ? (int)(((long long)j * p.B_img) / gridDim.y) : 0;
    const int j1 = WORDS ? (int)(((long long)(j + 1) * p.B_img) / gridDim.y) : 0;
    const int NS = WORDS ? (j1 - j0) * p.RKC : 2 * KCH;

    if (tid == 0) {
        if (sbase & 1023u) __trap();
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(bar_d_full), 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_base_s), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_s, 0);

    if (warp == kProducerWarp) {
        for (int s = 0; s < NS; ++s) {
            const int st = s % kStages;
            if (s >= kStages) mbar_wait(smem_u32(&bar_empty[st]), (uint32_t)((s / kStages) - 1) & 1u);
            if (elect_one()) {
                const uint32_t full = smem_u32(&bar_full[st]), dst = sbase + st * kGStageBytes;
                mbar_expect_tx(full, (uint32_t)(2 * kTileBytes + 2 * nef * kKC * 4));
                if constexpr (WORDS) {
                    const int jj = j0 + s / p.RKC, kc = (s - (jj - j0) * p.RKC) * kKC;
                    tma_load_2d(dst, &tm_u_hi, kc, jj * p.ncols + m * 128, full);
                    tma_load_2d(dst + kTileBytes, &tm_u_lo, kc, jj * p.ncols + m * 128, full);
                    tma_load_2d(dst + 2 * kTileBytes, &tm_w_hi, kc, jj * nef, full);
                    tma_load_2d(dst + 2 * kTileBytes + kGBTileBytes, &tm_w_lo, kc, jj * nef, full);
                } else {
                    const bool second = s >= KCH;
                    const int kc = (second ? s - KCH : s) * kKC;
                    tma_load_2d(dst, second ? &tm_a2_hi : &tm_u_hi, kc, j * p.R + m * 128, full);
                    tma_load_2d(dst + kTileBytes, second ? &tm_a2_lo : &tm_u_lo, kc, j * p.R + m * 128, full);
                    if (!second) {
                        tma_load_2d(dst + 2 * kTileBytes, &tm_w_hi, kc, 0, full);
                        tma_load_2d(dst + 2 * kTileBytes + kGBTileBytes, &tm_w_lo, kc, 0, full);
                    } else {
                        // q^T rows [32 columns n][32 channels] per box: the MN-major form of the [channel][column] operand
                        for (int cb = 0; cb < nef / 32; ++cb) {
                            tma_load_2d(dst + 2 * kTileBytes + cb * 4096, &tm_qmn_hi, cb * 32, j * p.ncols + kc, full);
                            tma_load_2d(dst + 2 * kTileBytes + kGBTileBytes + cb * 4096, &tm_qmn_lo, cb * 32, j * p.ncols + kc, full);
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        constexpr uint32_t kHi = desc_hi(1024, kSwizzle128B), kMnHi = desc_hi(512, kSwizzle128B_Base32B);
        const uint32_t idesc = make_idesc(2, 0, 0, 128, nef), idesc_mn = make_idesc(2, 0, 1, 128, nef);
        for (int s = 0; s < NS; ++s) {
            const int st = s % kStages;
            mbar_wait(smem_u32(&bar_full[st]), (uint32_t)(s / kStages) & 1u);
            tc_fence_after();
            const uint32_t base = sbase + st * kGStageBytes;
            const uint32_t a_hi = desc_lo(base, 16), a_lo = desc_lo(base + kTileBytes, 16);
            // B: K-major [channel][32 columns] tiles (wcp; X in the d_words mode), or - second term of d_img - q^T boxes
            // [32 columns][32 channels] as an MN-major operand: 4096 bytes between the 32-channel boxes, 4-row atoms of
            // 512 bytes, 8 rows (one k-step) = 1024 bytes
            const bool mn = !WORDS && s >= KCH;
            const uint32_t lbo = mn ? 4096u : 16u;
            const uint32_t b_hi = desc_lo(base + 2 * kTileBytes, lbo), b_lo = desc_lo(base + 2 * kTileBytes + kGBTileBytes, lbo);
            const uint32_t kBh = mn ? kMnHi : kHi, id = mn ? idesc_mn : idesc, bstep = mn ? 64u : 2u;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kKC / 8; ++ks) {
                    const uint32_t o = (uint32_t)(ks * 2), ob = (uint32_t)ks * bstep;
                    umma_ss<true>(tmem_base, a_hi + o, kHi, b_hi + ob, kBh, id, (s > 0 || ks > 0) ? 1u : 0u);
                    umma_ss<true>(tmem_base, a_hi + o, kHi, b_lo + ob, kBh, id, 1u);
                    umma_ss<true>(tmem_base, a_lo + o, kHi, b_hi + ob, kBh, id, 1u);
                }
                umma_commit(smem_u32(&bar_empty[st]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(bar_d_full));
        __syncwarp();
    } else {
        const int q = warp & 3;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        const int r = m * 128 + q * 32 + lane;           // d_img: region; d_words: word column
        mbar_wait(smem_u32(bar_d_full), 0u);
        tc_fence_after();
        for (int c0 = 0; c0 < nef; c0 += 32) {
            uint32_t d[32];
            tmem_ld<32>(tl + c0, d);
            tmem_wait_ld();
            if constexpr (WORDS) {
                float4* out = reinterpret_cast<float4*>(p.part + ((size_t)j * p.ncols + r) * nef + c0);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    out[c4] = make_float4(__uint_as_float(d[4 * c4]), __uint_as_float(d[4 * c4 + 1]), __uint_as_float(d[4 * c4 + 2]),
                                          __uint_as_float(d[4 * c4 + 3]));
            } else {
                float* out = p.d_img + (size_t)j * nef * p.R + r;
                if (r < p.R) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) out[(size_t)(c0 + c) * p.R] = __uint_as_float(d[c]);
                }
            }
            __syncwarp();                            // (tcgen05.ld is warp-collective; the stores are lane-dependent)
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

constexpr size_t kWtGemmSmem = (size_t)kStages * kGStageBytes + (2 * kStages + 1 + 1) * 8;

// d_words[i][c][t] = sum of the split partials (in split order) + (sum_j kap[j][n]) / |w_n|^2 * w_n[c]  for the caption's
// T_i words, 0 beyond (losses.py:72-76 slices the caption; autograd leaves the padding's gradient at zero).
// block = (caption i), thread = channel c (strided)
__global__ void __launch_bounds__(256) k_wt_dwords(const float* __restrict__ part, const float* __restrict__ kap,
                                                   const float* __restrict__ wt_hi, const float* __restrict__ wt_lo,
                                                   const float* __restrict__ ww, const int* __restrict__ cap_col,
                                                   const int* __restrict__ cap_lens, float* __restrict__ d_words, int splits,
                                                   int ncols, int nef, int B_img, int Lw) {
    __shared__ float kn[32];
    const int i = blockIdx.x;
    int T = cap_lens[i];
    T = T < 0 ? 0 : (T > Lw ? Lw : T);
    const int n0 = cap_col[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < T; t += 8) {          // kappa_t: images added in order by one lane-strided tree per word
        float a = 0.f;
        for (int j = lane; j < B_img; j += 32) a += kap[(size_t)j * ncols + n0 + t];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const float w = ww[n0 + t];
        if (lane == 0) kn[t] = w > 0.f ? a / (w * w) : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nef; c += 256) {
        float* out = d_words + ((size_t)i * nef + c) * Lw;
        for (int t = 0; t < Lw; ++t) {
            float a = 0.f;
            if (t < T) {
                const size_t o = (size_t)(n0 + t) * nef + c;
                for (int s = 0; s < splits; ++s) a += part[(size_t)s * ncols * nef + o];
                a = fmaf(kn[t], wt_hi[o] + wt_lo[o], a);
            }
            out[t] = a;
        }
    }
}

constexpr size_t kWtBwdSmem = (size_t)kBStages * kBStageBytes + kNB * 16 + kMaxCapsPerHalf * 256 * 4 + 8 * kBStageF4 * 16 +
                              8 * 2 * 8 * 32 * 4 /* WORDS: u2 staging */ + (2 * kBStages + 4 + 1) * 8;
static_assert(kWtBwdSmem <= 232448, "phase B: operand ring + caption totals + output staging must fit the 227 KB of one CTA");

constexpr size_t kWtSmem = (size_t)kStages * kStageBytes + (size_t)kEBufs * kEBufBytes + (size_t)(kMaxChunks * kNB + kNB) * 4 +
                           (2 * kStages + 4 + kEBufs + 4 + 3) * 8;

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WtLayout {
    int ncols, n_half_max, MT, RKC, RKP, RMP;
    size_t plan, cap_col, col_cap, col_T, ww, wt_hi, wt_lo, wt_f, x_hi, x_lo, xt_hi, xt_lo, total;
    size_t wcp_hi, wcp_lo, scal, qt_hi, qt_lo, u_hi, u_lo, a2_hi, a2_lo;       // backward only
    size_t a1;                                                                               // backward only
    size_t kap, u2_hi, u2_lo, part;                                                          // word gradients only
    int splits;
};
WtLayout wt_layout(int B_img, int B_cap, int nef, int R, int Lw, bool bwd = false, bool dwords = false) {
    WtLayout w{};
    int per_half = kHalf / Lw < 1 ? 1 : kHalf / Lw;                  // captions per half block, worst case
    if (per_half > kMaxCapsPerHalf) per_half = kMaxCapsPerHalf;
    w.n_half_max = (B_cap + per_half - 1) / per_half;
    w.n_half_max += w.n_half_max & 1;                                // whole 128-column blocks
    w.ncols = w.n_half_max * kHalf;
    w.MT = (R + 127) / 128;
    w.RKC = (R + 31) / 32;
    w.RKP = w.RKC * 32;
    w.RMP = w.MT * 128;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += al256(bytes); return at; };
    w.plan = take(sizeof(WtPlan));
    w.cap_col = take((size_t)B_cap * 4);
    w.col_cap = take((size_t)w.ncols * 4);
    w.col_T = take((size_t)w.ncols * 4);
    w.ww = take((size_t)w.ncols * 4);
    w.wt_hi = take((size_t)w.ncols * nef * 4);
    w.wt_lo = take((size_t)w.ncols * nef * 4);
    w.wt_f = take((size_t)w.ncols * nef * 4);
    w.x_hi = take((size_t)B_img * nef * w.RKP * 4);
    w.x_lo = take((size_t)B_img * nef * w.RKP * 4);
    w.xt_hi = take((size_t)B_img * w.RMP * nef * 4);
    w.xt_lo = take((size_t)B_img * w.RMP * nef * 4);
    if (bwd) {
        w.wcp_hi = take((size_t)nef * w.ncols * 4);
        w.wcp_lo = take((size_t)nef * w.ncols * 4);
        w.scal = take((size_t)B_img * w.ncols * 16);
        w.qt_hi = take((size_t)B_img * w.ncols * nef * 4);
        w.qt_lo = take((size_t)B_img * w.ncols * nef * 4);
        w.u_hi = take((size_t)B_img * R * w.ncols * 4);
        w.u_lo = take((size_t)B_img * R * w.ncols * 4);
        w.a2_hi = take((size_t)B_img * R * w.ncols * 4);
        w.a2_lo = take((size_t)B_img * R * w.ncols * 4);
        w.a1 = take((size_t)B_img * w.ncols * w.RKP * 4);
    }
    w.splits = B_img < 16 ? B_img : 16;           // image splits of the d_words GEMM (partials added in order)
    if (bwd && dwords) {
        w.kap = take((size_t)B_img * w.ncols * 4);
        w.u2_hi = take((size_t)B_img * w.ncols * w.RKP * 4);
        w.u2_lo = take((size_t)B_img * w.ncols * w.RKP * 4);
        w.part = take((size_t)w.splits * w.ncols * nef * 4);
    }
    w.total = o;
    return w;
}

}  // namespace

namespace {
int make_k128_map(CUtensorMap* out, const void* base, long long rows, int cols, int box_rows, bool mn32) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
        return (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) ? reinterpret_cast<EncodeFn>(f) : (EncodeFn) nullptr;
    }();
    bind_context_to_thread();
    if (encode == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SBA_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a %lld x %d fp32 operand, box %d x %d", (int)r, rows, cols,
                  box_rows, kKC);
        return SBA_ERR_CUDA;
    }
    return SBA_OK;
}
}  // namespace

bool words_tc5_supports(int B_img, int B_cap, int nef, int R, int Lw) {
    return nef % 32 == 0 && nef >= 32 && nef <= 256 && R >= 1 && R <= 384 && Lw >= 1 && Lw <= 32 && B_img >= 1 && B_cap >= 1 &&
           B_cap <= kPlanMaxCaps &&
           (long long)B_img * 384 * nef < (1ll << 31);
}

size_t words_tc5_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw) {
    if (!words_tc5_supports(B_img, B_cap, nef, R, Lw)) return 0;
    return wt_layout(B_img, B_cap, nef, R, Lw).total;
}

int words_sim_fwd_tc5(const float* img, const float* words, const int* cap_lens, float* sim, void* workspace, size_t ws_bytes,
                      int B_img, int B_cap, int nef, int R, int Lw, float g1, float g2, float g3, float eps, cudaStream_t st) {
    if (!words_tc5_supports(B_img, B_cap, nef, R, Lw)) {
        set_error("words_sim_fwd(tcgen05): shape nef=%d R=%d Lw=%d not covered", nef, R, Lw);
        return SBA_ERR_UNSUPPORTED;
    }
    const WtLayout w = wt_layout(B_img, B_cap, nef, R, Lw);
    if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
        set_error("words_sim_fwd(tcgen05): workspace of %zu bytes given, %zu (256-byte aligned) needed", ws_bytes, w.total);
        return SBA_ERR_ARG;
    }
    int dev = 0, sms = 0;
    int rc = current_device(&dev, &sms, "words_sim_fwd(tcgen05)");
    if (rc) return rc;
    static std::atomic<unsigned long long> smem_done{0};
    rc = ensure_dynamic_smem(k_words_tc5<false>, kWtSmem, dev, smem_done, "words_sim_fwd(tcgen05)");
    if (rc) return rc;
    char* ws = static_cast<char*>(workspace);
    WtPlan* plan = reinterpret_cast<WtPlan*>(ws + w.plan);
    int* cap_col = reinterpret_cast<int*>(ws + w.cap_col);
    int* col_cap = reinterpret_cast<int*>(ws + w.col_cap);
    int* col_T = reinterpret_cast<int*>(ws + w.col_T);
    float* ww = reinterpret_cast<float*>(ws + w.ww);
    float* wt_hi = reinterpret_cast<float*>(ws + w.wt_hi);
    float* wt_lo = reinterpret_cast<float*>(ws + w.wt_lo);
    float* wt_f = reinterpret_cast<float*>(ws + w.wt_f);
    float* x_hi = reinterpret_cast<float*>(ws + w.x_hi);
    float* x_lo = reinterpret_cast<float*>(ws + w.x_lo);
    float* xt_hi = reinterpret_cast<float*>(ws + w.xt_hi);
    float* xt_lo = reinterpret_cast<float*>(ws + w.xt_lo);

    k_wt_plan<<<1, 256, 0, st>>>(cap_lens, B_cap, Lw, cap_col, plan, col_cap, col_T, w.ncols);
    k_wt_words<<<w.ncols / 8, 256, 0, st>>>(words, col_cap, cap_col, wt_hi, wt_lo, wt_f, ww, nullptr, nullptr, w.ncols, nef, Lw);
    k_wt_images<<<dim3(w.RMP / 32, nef / 32, B_img), 256, 0, st>>>(img, x_hi, x_lo, xt_hi, xt_lo, nef, R, w.RKP, w.RMP);
    add_launches(3);
    rc = check_launch("words_sim_fwd(tcgen05 pre-pass)");
    if (rc) return rc;

    CUtensorMap tm[6], tm_wf;
    rc = make_k128_map(&tm_wf, wt_f, w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[0], xt_hi, (long long)B_img * w.RMP, nef, 128);
    if (!rc) rc = make_k128_map(&tm[1], xt_lo, (long long)B_img * w.RMP, nef, 128);
    if (!rc) rc = make_k128_map(&tm[2], wt_hi, w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[3], wt_lo, w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[4], x_hi, (long long)B_img * nef, w.RKP, nef);
    if (!rc) rc = make_k128_map(&tm[5], x_lo, (long long)B_img * nef, w.RKP, nef);
    if (rc) return rc;
    WtParams p{};
    p.wt_hi = wt_hi; p.wt_lo = wt_lo; p.ww = ww; p.col_cap = col_cap; p.col_T = col_T; p.plan = plan; p.sim = sim;
    p.B_cap = B_cap; p.nef = nef; p.R = R; p.MT = w.MT; p.RKC = w.RKC;
    p.g1l2e = g1 * 1.4426950408889634f; p.g2 = g2; p.g3 = g3; p.eps = eps;
    k_words_tc5<false><<<dim3(w.n_half_max / 2, B_img), kWtThreads, kWtSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm_wf, tm_wf,
                                                                                  tm_wf, p);      // (q^T maps: backward only)
    add_launches(1);
    return check_launch("words_sim_fwd(tcgen05)");
}

size_t words_tc5_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw, bool need_words) {
    if (!words_tc5_supports(B_img, B_cap, nef, R, Lw)) return 0;
    const WtLayout w = wt_layout(B_img, B_cap, nef, R, Lw, true, need_words);
    if ((long long)B_img * w.ncols >= (1ll << 31) / 2) return 0;       // TMA row coordinates of q^T
    return w.total;
}

// d_img always; d_words (nullable: GAN training detaches the words, trainer_bert.py:257) for DAMSM pre-training
int words_sim_bwd_tc5(const float* img, const float* words, const int* cap_lens, const float* d_sim, float* d_img,
                      float* d_words, void* workspace, size_t ws_bytes, int B_img, int B_cap, int nef, int R, int Lw, float g1,
                      float g2, float g3, float eps, cudaStream_t st) {
    if (!words_tc5_supports(B_img, B_cap, nef, R, Lw)) {
        set_error("words_sim_bwd(tcgen05): shape nef=%d R=%d Lw=%d not covered", nef, R, Lw);
        return SBA_ERR_UNSUPPORTED;
    }
    const bool dwords = d_words != nullptr;
    const WtLayout w = wt_layout(B_img, B_cap, nef, R, Lw, true, dwords);
    if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
        set_error("words_sim_bwd(tcgen05): workspace of %zu bytes given, %zu (256-byte aligned) needed", ws_bytes, w.total);
        return SBA_ERR_ARG;
    }
    int dev = 0, sms = 0;
    int rc = current_device(&dev, &sms, "words_sim_bwd(tcgen05)");
    if (rc) return rc;
    static std::atomic<unsigned long long> smem_a{0}, smem_b{0}, smem_c{0}, smem_d{0}, smem_e{0};
    rc = ensure_dynamic_smem(k_words_tc5<true>, kWtSmem, dev, smem_a, "words_sim_bwd(tcgen05)");
    if (!rc) rc = ensure_dynamic_smem(k_words_bwd_tc5<false>, kWtBwdSmem, dev, smem_b, "words_sim_bwd(tcgen05)");
    if (!rc) rc = ensure_dynamic_smem(k_words_bwd_tc5<true>, kWtBwdSmem, dev, smem_e, "words_sim_bwd(tcgen05)");
    if (!rc) rc = ensure_dynamic_smem(k_words_dimg_tc5<false>, kWtGemmSmem, dev, smem_c, "words_sim_bwd(tcgen05)");
    if (!rc) rc = ensure_dynamic_smem(k_words_dimg_tc5<true>, kWtGemmSmem, dev, smem_d, "words_sim_bwd(tcgen05)");
    if (rc) return rc;
    char* ws = static_cast<char*>(workspace);
    WtPlan* plan = reinterpret_cast<WtPlan*>(ws + w.plan);
    int* cap_col = reinterpret_cast<int*>(ws + w.cap_col);
    int* col_cap = reinterpret_cast<int*>(ws + w.col_cap);
    int* col_T = reinterpret_cast<int*>(ws + w.col_T);
    float* ww = reinterpret_cast<float*>(ws + w.ww);
    float* wt_hi = reinterpret_cast<float*>(ws + w.wt_hi);
    float* wt_lo = reinterpret_cast<float*>(ws + w.wt_lo);
    float* wt_f = reinterpret_cast<float*>(ws + w.wt_f);
    float* x_hi = reinterpret_cast<float*>(ws + w.x_hi);
    float* x_lo = reinterpret_cast<float*>(ws + w.x_lo);
    float* xt_hi = reinterpret_cast<float*>(ws + w.xt_hi);
    float* xt_lo = reinterpret_cast<float*>(ws + w.xt_lo);
    float* wcp_hi = reinterpret_cast<float*>(ws + w.wcp_hi);
    float* wcp_lo = reinterpret_cast<float*>(ws + w.wcp_lo);
    float4* scal = reinterpret_cast<float4*>(ws + w.scal);
    float* qt_hi = reinterpret_cast<float*>(ws + w.qt_hi);
    float* qt_lo = reinterpret_cast<float*>(ws + w.qt_lo);
    float* u_hi = reinterpret_cast<float*>(ws + w.u_hi);
    float* u_lo = reinterpret_cast<float*>(ws + w.u_lo);
    float* a2_hi = reinterpret_cast<float*>(ws + w.a2_hi);
    float* a2_lo = reinterpret_cast<float*>(ws + w.a2_lo);
    float* a1 = reinterpret_cast<float*>(ws + w.a1);
    float* kap = dwords ? reinterpret_cast<float*>(ws + w.kap) : nullptr;
    float* u2_hi = dwords ? reinterpret_cast<float*>(ws + w.u2_hi) : nullptr;
    float* u2_lo = dwords ? reinterpret_cast<float*>(ws + w.u2_lo) : nullptr;
    float* part = dwords ? reinterpret_cast<float*>(ws + w.part) : nullptr;

    k_wt_plan<<<1, 256, 0, st>>>(cap_lens, B_cap, Lw, cap_col, plan, col_cap, col_T, w.ncols);
    k_wt_words<<<w.ncols / 8, 256, 0, st>>>(words, col_cap, cap_col, wt_hi, wt_lo, wt_f, ww, wcp_hi, wcp_lo, w.ncols, nef, Lw);
    k_wt_images<<<dim3(w.RMP / 32, nef / 32, B_img), 256, 0, st>>>(img, x_hi, x_lo, xt_hi, xt_lo, nef, R, w.RKP, w.RMP);
    add_launches(3);
    rc = check_launch("words_sim_bwd(tcgen05 pre-pass)");
    if (rc) return rc;

    CUtensorMap tm[18], tm_wf;
    rc = make_k128_map(&tm_wf, wt_f, w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[0], xt_hi, (long long)B_img * w.RMP, nef, 128);
    if (!rc) rc = make_k128_map(&tm[1], xt_lo, (long long)B_img * w.RMP, nef, 128);
    if (!rc) rc = make_k128_map(&tm[2], wt_hi, w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[3], wt_lo, w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[4], x_hi, (long long)B_img * nef, w.RKP, nef);
    if (!rc) rc = make_k128_map(&tm[5], x_lo, (long long)B_img * nef, w.RKP, nef);
    if (!rc) rc = make_k128_map(&tm[6], qt_hi, (long long)B_img * w.ncols, nef, 128);
    if (!rc) rc = make_k128_map(&tm[7], qt_lo, (long long)B_img * w.ncols, nef, 128);
    // the d_img GEMM: A = u^T / a2^T [B_img * R][ncols] (tiles of 128 regions; rows past the last image read as zero),
    // B = wcp [nef][ncols] and q^T [B_img * ncols][nef] (MN-major)
    if (!rc) rc = make_k128_map(&tm[8], u_hi, (long long)B_img * R, w.ncols, 128);
    if (!rc) rc = make_k128_map(&tm[9], u_lo, (long long)B_img * R, w.ncols, 128);
    if (!rc) rc = make_k128_map(&tm[10], a2_hi, (long long)B_img * R, w.ncols, 128);
    if (!rc) rc = make_k128_map(&tm[11], a2_lo, (long long)B_img * R, w.ncols, 128);
    if (!rc) rc = make_k128_map(&tm[12], wcp_hi, nef, w.ncols, nef);
    if (!rc) rc = make_k128_map(&tm[13], wcp_lo, nef, w.ncols, nef);
    if (!rc) rc = make_k128_map(&tm[14], qt_hi, (long long)B_img * w.ncols, nef, 32, true);      // q^T again, MN-major boxes
    if (!rc) rc = make_k128_map(&tm[15], qt_lo, (long long)B_img * w.ncols, nef, 32, true);
    if (dwords) {
        if (!rc) rc = make_k128_map(&tm[16], u2_hi, (long long)B_img * w.ncols, w.RKP, 128);
        if (!rc) rc = make_k128_map(&tm[17], u2_lo, (long long)B_img * w.ncols, w.RKP, 128);
    }
    if (rc) return rc;
    WtParams p{};
    p.wt_hi = wt_hi; p.wt_lo = wt_lo; p.ww = ww; p.col_cap = col_cap; p.col_T = col_T; p.plan = plan; p.sim = nullptr;
    p.B_cap = B_cap; p.nef = nef; p.R = R; p.MT = w.MT; p.RKC = w.RKC;
    p.g1l2e = g1 * 1.4426950408889634f; p.g2 = g2; p.g3 = g3; p.eps = eps;
    p.d_sim = d_sim; p.cap_col = cap_col; p.scal = scal; p.qt_hi = qt_hi; p.qt_lo = qt_lo;
    p.ncols = w.ncols; p.g1 = g1; p.kap = kap; p.a1 = a1; p.RKP = w.RKP;
    const dim3 grid(w.n_half_max / 2, B_img);
    k_words_tc5<true><<<grid, kWtThreads, kWtSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm_wf, tm[6], tm[7], p);
    rc = check_launch("words_sim_bwd(tcgen05 phase A)");
    if (rc) return rc;
    WtBwdParams b{};
    b.col_cap = col_cap; b.plan = plan; b.scal = scal; b.a1 = a1; b.u_hi = u_hi; b.u_lo = u_lo; b.a2_hi = a2_hi; b.a2_lo = a2_lo;
    b.u2_hi = u2_hi; b.u2_lo = u2_lo; b.RKP = w.RKP;
    b.nef = nef; b.R = R; b.MT = w.MT; b.RKC = w.RKC; b.ncols = w.ncols; b.g1 = g1; b.g1l2e = p.g1l2e;
    if (dwords)
        k_words_bwd_tc5<true><<<grid, kWtThreads, kWtBwdSmem, st>>>(tm[0], tm[1], tm[6], tm[7], b);
    else
        k_words_bwd_tc5<false><<<grid, kWtThreads, kWtBwdSmem, st>>>(tm[0], tm[1], tm[6], tm[7], b);
    rc = check_launch("words_sim_bwd(tcgen05 phase B)");
    if (rc) return rc;
    WtGemmParams g{};
    g.plan = plan; g.d_img = d_img; g.nef = nef; g.R = R; g.ncols = w.ncols;
    k_words_dimg_tc5<false><<<dim3(w.MT, B_img), kGThreads, kWtGemmSmem, st>>>(tm[8], tm[9], tm[10], tm[11], tm[12], tm[13],
                                                                                tm[14], tm[15], g);
    add_launches(3);
    rc = check_launch("words_sim_bwd(tcgen05 d_img)");
    if (rc || !dwords) return rc;
    g.part = part; g.ncols = w.ncols; g.B_img = B_img; g.RKC = w.RKC;
    // (maps 2, 3, 6, 7 are unused in this mode; 4, 5 = X hi / lo as in the forward's second GEMM)
    k_words_dimg_tc5<true><<<dim3(w.n_half_max / 2, w.splits), kGThreads, kWtGemmSmem, st>>>(tm[16], tm[17], tm[16], tm[17], tm[4],
                                                                                              tm[5], tm[4], tm[5], g);
    k_wt_dwords<<<B_cap, 256, 0, st>>>(part, kap, wt_hi, wt_lo, ww, cap_col, cap_lens, d_words, w.splits, w.ncols, nef, B_img, Lw);
    add_launches(2);
    return check_launch("words_sim_bwd(tcgen05 d_words)");
}

}  // namespace sba

#ifdef SBA_DEV_AIDS
// per-CTA phase times of k_words_tc5 summed over CTAs: [count, tile loop, wait for the last G2, final epilogue] x {fwd, phase A}
extern "C" __attribute__((visibility("default"))) void sba_dev_words_times(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, sba::g_wt_ns, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(sba::g_wt_ns, z, sizeof(z));
    }
}
// progress / timeout counters in mapped host memory: readable while a kernel hangs
extern "C" __attribute__((visibility("default"))) unsigned* sba_dev_words_progress(void) {
    static unsigned* host = nullptr;
    if (host == nullptr) {
        cudaHostAlloc(&host, 32 * sizeof(unsigned), cudaHostAllocMapped);
        for (int i = 0; i < 32; ++i) host[i] = 0;
        unsigned* dev = nullptr;
        cudaHostGetDevicePointer(&dev, host, 0);
        cudaMemcpyToSymbol(sba::g_wt_host, &dev, sizeof(dev));
    }
    return host;
}
#endif

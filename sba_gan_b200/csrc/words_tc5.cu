// Kernel (c) forward on the 5th-generation tensor cores: the B x B DAMSM region-word similarity of words_loss
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

constexpr int kWtThreads = 64 + 256;      // producer warp, MMA warp, 8 epilogue warps
constexpr int kNB = 128;                  // word columns per CTA
constexpr int kHalf = 64;                 // columns per half block (one epilogue warp per lane quarter each)
constexpr int kKC = 32;                   // floats per 128-byte operand row = K per pipeline stage
constexpr int kTileBytes = 128 * kKC * 4; // one [128 rows][32 floats] operand tile: 16 KB
constexpr int kStageBytes = 4 * kTileBytes;       // G1: A hi, A lo, B hi, B lo;  G2: X hi, X lo ([256][32] each)
constexpr int kStages = 2;
constexpr int kEBufs = 2;
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
    float* sim;                // [B_img][B_cap]
    int B_cap, nef, R, MT, RKC;
    float g1l2e, g2, g3, eps;  // gamma1 * log2(e), gamma2, gamma3
};

// ---- pre-pass 1: pack the captions into half blocks (sequential greedy, one thread; B_cap is a few hundred) ---------
__global__ void k_wt_plan(const int* __restrict__ cap_lens, int B_cap, int Lw, int* __restrict__ cap_col, WtPlan* plan,
                          int* __restrict__ col_cap, int* __restrict__ col_T, int ncols) {
    for (int n = threadIdx.x; n < ncols; n += blockDim.x) { col_cap[n] = -1; col_T[n] = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int h = 0, off = 0;
        for (int i = 0; i < B_cap; ++i) {
            int T = cap_lens[i];
            T = T < 0 ? 0 : (T > Lw ? Lw : T);
            if (off + T > kHalf) { ++h; off = 0; }
            cap_col[i] = h * kHalf + off;
            if (T > 0) col_T[h * kHalf + off] = T;
            for (int t = 0; t < T; ++t) col_cap[h * kHalf + off + t] = i;
            off += T;
        }
        plan->n_half = h + 1;
    }
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = tf32_rna(v);
    lo = tf32_rna(v - hi);
}

// ---- pre-pass 2: WT hi / lo [ncols][nef] and |w_n|; one block per 8 columns, thread = channel --------------------------
__global__ void __launch_bounds__(256) k_wt_words(const float* __restrict__ words, const int* __restrict__ col_cap,
                                                  const int* __restrict__ cap_col, float* __restrict__ wt_hi,
                                                  float* __restrict__ wt_lo, float* __restrict__ ww, int nef, int Lw) {
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

// fp32 K-major operand tiles: rows of 32 floats (128 bytes), standard 128-byte swizzle
int make_k128_map(CUtensorMap* out, const void* base, long long rows, int cols, int box_rows);

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

__global__ void __launch_bounds__(kWtThreads, 1)
    k_words_tc5(const __grid_constant__ CUtensorMap tm_xt_hi, const __grid_constant__ CUtensorMap tm_xt_lo,
                const __grid_constant__ CUtensorMap tm_wt_hi, const __grid_constant__ CUtensorMap tm_wt_lo,
                const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo, const WtParams p) {
    const int nb = blockIdx.x, j = blockIdx.y;
    if (2 * nb >= p.plan->n_half) return;                      // the grid is sized for the worst packing

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
    unsigned long long* bar_e_free = bar_e_ready + kEBufs;  // [kEBufs]
    unsigned long long* bar_d_full = bar_e_free + kEBufs;   // [1]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bar_d_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nef = p.nef, MT = p.MT, RKC = p.RKC, KCH = nef / kKC;

    if (tid == 0) {
        if (sbase & 1023u) __trap();
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_s_full[s]), 1); mbar_init(smem_u32(&bar_s_free[s]), 8); }
        for (int s = 0; s < kEBufs; ++s) { mbar_init(smem_u32(&bar_e_ready[s]), 2); mbar_init(smem_u32(&bar_e_free[s]), 1); }
        mbar_init(smem_u32(bar_d_full), 1);
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
            if (s >= kStages) mbar_wait(smem_u32(&bar_empty[st]), (uint32_t)((s / kStages) - 1) & 1u);
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
            if (s >= kStages) mbar_wait(smem_u32(&bar_empty[st]), (uint32_t)((s / kStages) - 1) & 1u);
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
    } else if (warp == kMmaWarp) {
        // ------------------------------- MMA issuer ------------------------------------------------------------------
        constexpr uint32_t kHi = desc_hi(1024, kSwizzle128B);
        const uint32_t idesc1 = make_idesc(2, 0, 0, 128, kNB);
        const uint32_t idesc2 = make_idesc(2, 0, 0, 128, nef);
        int s = 0;
        bool d_started = false;
        auto g1 = [&](int m) {
            const int buf = m & 1;
            if (m >= 2) mbar_wait(smem_u32(&bar_s_free[buf]), (uint32_t)((m >> 1) - 1) & 1u);
            const uint32_t d = tmem_base + COL_S + 128 * buf;
            for (int kc = 0; kc < KCH; ++kc, ++s) {
                const int st = s % kStages;
                mbar_wait(smem_u32(&bar_full[st]), (uint32_t)(s / kStages) & 1u);
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
                mbar_wait(smem_u32(&bar_e_ready[eb]), (uint32_t)(k / kEBufs) & 1u);
                mbar_wait(smem_u32(&bar_full[st]), (uint32_t)(s / kStages) & 1u);
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
                    umma_commit(smem_u32(&bar_e_free[eb]));
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
            mbar_wait(smem_u32(&bar_s_full[buf]), (uint32_t)(m >> 1) & 1u);
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
#pragma unroll
            for (int c = kHalf - 1; c >= 0; --c) {
                seg = ((last >> c) & 1ull) ? mma::rcp_approx(__uint_as_float(sr[c])) : seg;
                const float a1 = v[c] * seg;
                // e = exp(gamma1 (a1 - 1)); padding columns and regions beyond R contribute nothing
                v[c] = (rv && ((valid >> c) & 1ull)) ? mma::ex2_approx((a1 - 1.f) * p.g1l2e) : 0.f;
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
            if (k >= kEBufs) mbar_wait(smem_u32(&bar_e_free[eb]), (uint32_t)((k / kEBufs) - 1) & 1u);
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
        if (ew < 4) {
            const int n = 32 * q + lane, ng = nb * kNB + n;
            mbar_wait(smem_u32(bar_d_full), 0u);
            tc_fence_after();
            float num = 0.f, wn2 = 0.f;
            const float4* wh = reinterpret_cast<const float4*>(p.wt_hi + (size_t)ng * nef);
            const float4* wl = reinterpret_cast<const float4*>(p.wt_lo + (size_t)ng * nef);
            for (int c0 = 0; c0 < nef; c0 += 32) {
                uint32_t d[32];
                tmem_ld<32>(tl + COL_D + c0, d);
                tmem_wait_ld();
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 a = __ldg(wh + (c0 >> 2) + c4), b = __ldg(wl + (c0 >> 2) + c4);
                    const float w[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
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
            const bool cv = p.col_cap[ng] >= 0;
            float ex = 0.f;
            if (cv) {
                const float invZ = 1.0f / Z;
                const float den = fmaxf(p.ww[ng] * sqrtf(wn2) * invZ, p.eps);        // losses.py:17
                ex = expf(p.g2 * (num * invZ) / den);                                 // losses.py:106
            }
            exs[n] = ex;
            mma::named_bar_sync(2, 128);
            const int T = p.col_T[ng];
            if (T > 0) {
                float E = 0.f;
                for (int t = 0; t < T; ++t) E += exs[n + t];
                p.sim[(size_t)j * p.B_cap + p.col_cap[ng]] = p.g3 * logf(E);        // losses.py:107-108, 123
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

constexpr size_t kWtSmem = (size_t)kStages * kStageBytes + (size_t)kEBufs * kEBufBytes + (size_t)(kMaxChunks * kNB + kNB) * 4 +
                           (2 * kStages + 4 + 2 * kEBufs + 2) * 8;

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WtLayout {
    int ncols, n_half_max, MT, RKC, RKP, RMP;
    size_t plan, cap_col, col_cap, col_T, ww, wt_hi, wt_lo, x_hi, x_lo, xt_hi, xt_lo, total;
};
WtLayout wt_layout(int B_img, int B_cap, int nef, int R, int Lw) {
    WtLayout w{};
    const int per_half = kHalf / Lw < 1 ? 1 : kHalf / Lw;            // captions per half block, worst case
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
    w.x_hi = take((size_t)B_img * nef * w.RKP * 4);
    w.x_lo = take((size_t)B_img * nef * w.RKP * 4);
    w.xt_hi = take((size_t)B_img * w.RMP * nef * 4);
    w.xt_lo = take((size_t)B_img * w.RMP * nef * 4);
    w.total = o;
    return w;
}

}  // namespace

namespace {
int make_k128_map(CUtensorMap* out, const void* base, long long rows, int cols, int box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
        return (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) ? reinterpret_cast<EncodeFn>(f) : (EncodeFn) nullptr;
    }();
    if (encode == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SBA_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
    rc = ensure_dynamic_smem(k_words_tc5, kWtSmem, dev, smem_done, "words_sim_fwd(tcgen05)");
    if (rc) return rc;
    char* ws = static_cast<char*>(workspace);
    WtPlan* plan = reinterpret_cast<WtPlan*>(ws + w.plan);
    int* cap_col = reinterpret_cast<int*>(ws + w.cap_col);
    int* col_cap = reinterpret_cast<int*>(ws + w.col_cap);
    int* col_T = reinterpret_cast<int*>(ws + w.col_T);
    float* ww = reinterpret_cast<float*>(ws + w.ww);
    float* wt_hi = reinterpret_cast<float*>(ws + w.wt_hi);
    float* wt_lo = reinterpret_cast<float*>(ws + w.wt_lo);
    float* x_hi = reinterpret_cast<float*>(ws + w.x_hi);
    float* x_lo = reinterpret_cast<float*>(ws + w.x_lo);
    float* xt_hi = reinterpret_cast<float*>(ws + w.xt_hi);
    float* xt_lo = reinterpret_cast<float*>(ws + w.xt_lo);

    k_wt_plan<<<1, 256, 0, st>>>(cap_lens, B_cap, Lw, cap_col, plan, col_cap, col_T, w.ncols);
    k_wt_words<<<w.ncols / 8, 256, 0, st>>>(words, col_cap, cap_col, wt_hi, wt_lo, ww, nef, Lw);
    k_wt_images<<<dim3(w.RMP / 32, nef / 32, B_img), 256, 0, st>>>(img, x_hi, x_lo, xt_hi, xt_lo, nef, R, w.RKP, w.RMP);
    add_launches(3);
    rc = check_launch("words_sim_fwd(tcgen05 pre-pass)");
    if (rc) return rc;

    CUtensorMap tm[6];
    rc = make_k128_map(&tm[0], xt_hi, (long long)B_img * w.RMP, nef, 128);
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
    k_words_tc5<<<dim3(w.n_half_max / 2, B_img), kWtThreads, kWtSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    add_launches(1);
    return check_launch("words_sim_fwd(tcgen05)");
}

}  // namespace sba

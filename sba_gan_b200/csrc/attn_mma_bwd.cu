// Kernel (b): fused backward of GlobalAttentionGeneral on tensor cores (SBA_ALGO_MMA).
//
// Same skeleton as attn_mma_fwd.cu: one persistent CTA = 8 consumer warps + 1 TMA producer
// warp; the producer streams [idf x 128-pixel] tiles of x AND g_c into a shared-memory ring
// (1-D bulk copies, mbarrier completion); consumer warp w owns pixels [16w, 16w+16) of every
// tile.  Per 16-pixel m-tile, entirely in registers (formulas: SURVEY.md §8a-4,
// oracle/attention.py:attn_backward):
//     S  = x^T . srcT,  dP = g^T . srcT          (shared B fragments, two accumulators)
//     P  = masked softmax(S)                     (recomputed, not re-read from HBM)
//     dS = P * (dP [+ g_attn] - sum_l P dP)
//     dX = dS . srcT^T                           (dS re-used as A operand from the accumulators)
//     dSrc[ch][l] += g[ch][px] P[px][l] + x[ch][px] dS[px][l]
// The last contraction runs over the warp's 16 pixels (K = 16): its A operands are the
// channel-major view of the same x / g fragments (movmatrix transposes on the fp32 path,
// a second non-transposed ldmatrix on the bf16 path), its B operands are movmatrix
// transposes of the packed P / dS accumulators.  dSrc accumulates in 24 registers per thread
// across all tiles of a sample and is flushed with fp32 atomics once per (warp, sample).
// fp32 tensors: every operand is split into fp16 hi+lo with a power-of-two scale (three
// MMAs per product, ~2^-22 relative); the split sourceT fragments do not fit in registers
// next to the accumulators, so the CTA keeps them in a shared-memory fragment table
// (one conflict-free LDS.64 per B operand).  bf16 tensors: single bf16 MMAs, sourceT
// fragments register resident.
//
// dW / dContext are formed from the complete dSrc by attn_bwd_post (attn_tc5_bwd.cu).
#include "kernels.h"
#include "mma_common.cuh"

namespace sba {
namespace {
using namespace mma;

constexpr int kMaxStages = 4;

struct BwdParams {
    const void* x;
    const void* g;
    const void* ga;
    const float* srcT;
    const uint8_t* mask;
    void* dX;
    float* dSrc;
    const float* ctx;     // [B, cdf, L]   (epilogue)
    const float* W;       // [idf, cdf]    (epilogue, dCtx only)
    float* dW;            // [idf, cdf]    nullable
    float* dCtx;          // [B, cdf, L]   nullable
    uint32_t* cnt;        // [B] per-sample completion counters, [B] "dW is zeroed" flag; zero on entry
    int B, L, Q, cdf, mask_mode;
    int tiles_per_sample;
    int n_tiles;
    int nst;
};

template <typename T, int IDF, int NT>
struct BwdCfg {
    static constexpr int KS = IDF / 16;    // k-steps over channels (S, dP) == m-tiles over channels (dSrc)
    static constexpr int NC8 = IDF / 8;    // n-tiles over channels (dX)
    static constexpr int NK16 = NT / 2;    // k16 steps over words (dX)
    static constexpr bool HAS_K8 = (NT % 2) != 0;
    static constexpr int RS = TileStride<T>::value;
    static constexpr bool HALF = sizeof(T) == 4;   // fp32 tensors -> split fp16 MMAs
    static constexpr int NSPLIT = HALF ? 2 : 1;
    static constexpr int TILE_BYTES = IDF * RS * (int)sizeof(T);
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES;          // x tile, then g tile
    // B-operand fragments of sourceT (uint2 per lane), per split:
    static constexpr int NF_S = KS * NT;                        // S / dP phase: k = channel, n = word
    static constexpr int NF_C16 = NK16 * NC8;                   // dX phase: k = word (16), n = channel
    static constexpr int NF_C8 = HAS_K8 ? NC8 / 2 : 0;          // dX phase: k = last 8 words, two n-tiles per entry
    static constexpr int NFB = NF_S + NF_C16 + NF_C8;
    static constexpr int TAB_BYTES = HALF ? NSPLIT * NFB * 32 * 8 : 0;   // fragment table in shared memory (fp32 only)
    static constexpr int EPI_BYTES = IDF * 32 * 4;                       // dSrc[b] staged for the dW / dCtx epilogue
    static constexpr int SCR_BYTES = TAB_BYTES > EPI_BYTES ? TAB_BYTES : EPI_BYTES;   // the two never live together
};

// One fragment entry (the four sourceT values lane (g, c) contributes to fragment fb).
template <typename T, int IDF, int NT>
__device__ __forceinline__ void frag_values(int fb, int g, int c, const float* __restrict__ src, int L, float (&v)[4]) {
    using C = BwdCfg<T, IDF, NT>;
    auto ld = [&](int ch, int l) -> float { return (l < L) ? __ldcg(src + ch * L + l) : 0.f; };
    if (fb < C::NF_S) {
        // m16n8k16 B fragment: b0 = (k 2c,2c+1; n g), b1 = (k 2c+8,2c+9; n g); k = channel, n = word
        const int ks = fb / NT, nt = fb - ks * NT, l = nt * 8 + g;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int ch0, ch1;
            if (C::HALF) { ch0 = chan_of(ks, c, 2 * h); ch1 = chan_of(ks, c, 2 * h + 1); }
            else { ch0 = 16 * ks + 8 * h + 2 * c; ch1 = ch0 + 1; }
            v[2 * h] = ld(ch0, l);
            v[2 * h + 1] = ld(ch1, l);
        }
    } else if (fb < C::NF_S + C::NF_C16) {
        // k = word, n = channel
        const int e = fb - C::NF_S, j = e / C::NC8, nc = e - j * C::NC8, ch = nc * 8 + g;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int l0 = 16 * j + 8 * h + 2 * c;
            v[2 * h] = ld(ch, l0);
            v[2 * h + 1] = ld(ch, l0 + 1);
        }
    } else {
        // m16n8k8 B fragment b0 = (k 2c,2c+1; n g) for n-tiles 2e (x) and 2e+1 (y)
        const int e = fb - C::NF_S - C::NF_C16, l0 = (NT - 1) * 8 + 2 * c;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ch = (2 * e + h) * 8 + g;
            v[2 * h] = ld(ch, l0);
            v[2 * h + 1] = ld(ch, l0 + 1);
        }
    }
}

template <typename T, int IDF, int NT, bool HAS_GA>
__global__ void __launch_bounds__(kThreads, 2) k_attn_bwd_mma(const BwdParams p) {
    using C = BwdCfg<T, IDF, NT>;
    constexpr int KS = C::KS, NC8 = C::NC8, NK16 = C::NK16, RS = C::RS, NFB = C::NFB;
    constexpr bool HAS_K8 = C::HAS_K8, HALF = C::HALF;
    constexpr int NSPLIT = C::NSPLIT;
    constexpr float kLog2e = 1.4426950408889634f;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NST = p.nst;
    uint2* tab = reinterpret_cast<uint2*>(smem_raw + (size_t)NST * C::STAGE_BYTES);                        // [NSPLIT*NFB][32]
    uint32_t* mb_s = reinterpret_cast<uint32_t*>(smem_raw + (size_t)NST * C::STAGE_BYTES + C::SCR_BYTES);  // [B]
    __shared__ __align__(8) unsigned long long bar_full[kMaxStages], bar_empty[kMaxStages];
    __shared__ float red_s[kConsumerWarps];
    __shared__ int fin_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    const int L = p.L, Q = p.Q, TPS = p.tiles_per_sample;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), kConsumerWarps);
        }
        fence_barrier_init();
    }
    if (p.mask != nullptr) {
        for (int cap = tid; cap < p.B; cap += kThreads) {
            uint32_t bits = 0;
            for (int l = 0; l < L; ++l) bits |= (p.mask[(size_t)cap * L + l] ? 1u : 0u) << l;
            mb_s[cap] = bits;
        }
    }
    __syncthreads();

    const int w_begin = (int)(((long long)blockIdx.x * p.n_tiles) / gridDim.x);
    const int w_end = (int)(((long long)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    int b = w_begin / TPS, t = w_begin - b * TPS;

    if (warp == kConsumerWarps) {
        // ------------------------------ TMA producer warp ------------------------------------
        const T* xg = static_cast<const T*>(p.x);
        const T* gg = static_cast<const T*>(p.g);
        int stage = 0, phase = 0;
        for (int w = w_begin; w < w_end; ++w) {
            if (w - w_begin >= NST) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
            const uint32_t full = smem_u32(&bar_full[stage]);
            if (lane == 0) mbar_expect_tx(full, 2u * IDF * TQ * (uint32_t)sizeof(T));
            __syncwarp();
            const uint32_t dst0 = smem_u32(smem_raw) + stage * C::STAGE_BYTES;
            for (int r = lane; r < 2 * IDF; r += 32) {
                const bool isg = r >= IDF;
                const int ch = isg ? r - IDF : r;
                const T* srcp = (isg ? gg : xg) + ((size_t)b * IDF + ch) * Q + (size_t)t * TQ;
                tma_load_1d(dst0 + r * RS * (uint32_t)sizeof(T), srcp, TQ * (uint32_t)sizeof(T), full);
            }
            if (++t == TPS) { t = 0; ++b; }
            if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        return;
    }

    // ---------------------------------- consumer warps ---------------------------------------
    uint2 fr[HALF ? 1 : NFB];              // bf16: sourceT fragments register resident
    float dacc[KS][NT][4];                 // dSrc: m-tile = 16 channels, n-tile = 8 words
#pragma unroll
    for (int mt = 0; mt < KS; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { dacc[mt][nt][0] = 0.f; dacc[mt][nt][1] = 0.f; dacc[mt][nt][2] = 0.f; dacc[mt][nt][3] = 0.f; }

    auto flush_dsrc = [&](int bb) {
        float* db = p.dSrc + (size_t)bb * IDF * L;
#pragma unroll
        for (int mt = 0; mt < KS; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // accumulator row m = g + 8h of m-tile mt; fp32 path rows are in chan_of order
                const int ch = HALF ? (16 * mt + (g >> 1) + 4 * (g & 1) + 8 * h) : (16 * mt + g + 8 * h);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int l = nt * 8 + 2 * c + j;
                        if (l < L) atomicAdd(db + ch * L + l, dacc[mt][nt][2 * h + j]);
                        dacc[mt][nt][2 * h + j] = 0.f;
                    }
            }
    };
    // dSrc[bb] gets this CTA's share once its tile range leaves the sample; dW / dCtx are formed from the
    // complete dSrc by attn_bwd_post (a programmatic dependent of this kernel): doing that here, per sample,
    // puts 64-way contended atomics on dW right at the end of the stream.
    bool waited_zero = false;
    auto finish_sample = [&](int bb) {
        if (!waited_zero) {
            asm volatile("griddepcontrol.wait;" ::: "memory");      // the zero-fill grid in front has cleared dSrc / dW
            waited_zero = true;
        }
        flush_dsrc(bb);
    };
    // B-operand fragment fb of split sp (0 = hi, 1 = lo)
    auto frag = [&](int sp, int fb) -> uint2 {
        if constexpr (HALF) return tab[(sp * NFB + fb) * 32 + lane];
        else return fr[fb];
    };

    float inv_src = 1.f;
    int cur_b = -1;
    const uint32_t pad_bits = (L < 32) ? ~((1u << L) - 1u) : 0u;
    const uint32_t Bu = (uint32_t)p.B;
    const uint32_t step_mod = (uint32_t)TQ % Bu;
    // reference mask order: pixel n = b*Q + q uses caption n mod B (GlobalAttention.py:104-108)
    uint32_t cap0 = (uint32_t)(((unsigned long long)w_begin * TQ + warp * 16 + g) % Bu);
    int stage = 0, phase = 0;

    for (int w = w_begin; w < w_end; ++w) {
        if (b != cur_b) {
            if (cur_b >= 0) finish_sample(cur_b);
            cur_b = b;
            const float* sb = p.srcT + (size_t)b * IDF * L;
            if constexpr (HALF) {
                named_bar_sync(1, kConsumers);          // every warp is done with the old table
                float lm = 0.f;
                for (int o = tid; o < IDF * L; o += kConsumers) lm = fmaxf(lm, fabsf(__ldcg(sb + o)));
                lm = warp_absmax_redux(lm);
                if (lane == 0) red_s[warp] = lm;
                named_bar_sync(1, kConsumers);
                float m = red_s[0];
#pragma unroll
                for (int i = 1; i < kConsumerWarps; ++i) m = fmaxf(m, red_s[i]);
                float sc_src;
                pow2_scale(m, sc_src, inv_src);
                for (int i = tid; i < NFB * 32; i += kConsumers) {
                    const int fb = i >> 5, ln = i & 31;
                    float v[4];
                    frag_values<T, IDF, NT>(fb, ln >> 2, ln & 3, sb, L, v);
                    uint2 hi, lo;
                    split2(v[0], v[1], sc_src, hi.x, lo.x);
                    split2(v[2], v[3], sc_src, hi.y, lo.y);
                    tab[fb * 32 + ln] = hi;
                    tab[(NFB + fb) * 32 + ln] = lo;
                }
                named_bar_sync(1, kConsumers);
            } else {
#pragma unroll
                for (int fb = 0; fb < NFB; ++fb) {
                    float v[4];
                    frag_values<T, IDF, NT>(fb, g, c, sb, L, v);
                    fr[fb].x = pack_bf16(v[0], v[1]);
                    fr[fb].y = pack_bf16(v[2], v[3]);
                }
            }
        }

        // ---- px-major A fragments of x and g for this warp's 16 pixels ---------------------------
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        const T* xs = reinterpret_cast<const T*>(smem_raw + (size_t)stage * C::STAGE_BYTES) + warp * 16;
        const T* gs = xs + IDF * RS;
        uint32_t ax[NSPLIT][KS][4], ag[NSPLIT][KS][4];
        float inv_x = 1.f, inv_g = 1.f;
        if constexpr (HALF) {
            float xv[KS][8], gv[KS][8];
            float amx = 0.f, amg = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int off = chan_of(ks, c, j) * RS + g;
                    const float* xr = reinterpret_cast<const float*>(xs) + off;
                    const float* gr = reinterpret_cast<const float*>(gs) + off;
                    xv[ks][2 * j] = xr[0];        // pixel g
                    xv[ks][2 * j + 1] = xr[8];    // pixel g + 8
                    gv[ks][2 * j] = gr[0];
                    gv[ks][2 * j + 1] = gr[8];
                    amx = fmaxf(amx, fmaxf(fabsf(xv[ks][2 * j]), fabsf(xv[ks][2 * j + 1])));
                    amg = fmaxf(amg, fmaxf(fabsf(gv[ks][2 * j]), fabsf(gv[ks][2 * j + 1])));
                }
            float sc_x, sc_g;
            pow2_scale(warp_absmax_redux(amx), sc_x, inv_x);
            pow2_scale(warp_absmax_redux(amg), sc_g, inv_g);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                // a0 = (row g; k 2c,2c+1) a1 = (row g+8; same k) a2 = (row g; k 2c+8,2c+9) a3 = (row g+8; ...)
                split2(xv[ks][0], xv[ks][2], sc_x, ax[0][ks][0], ax[NSPLIT - 1][ks][0]);
                split2(xv[ks][1], xv[ks][3], sc_x, ax[0][ks][1], ax[NSPLIT - 1][ks][1]);
                split2(xv[ks][4], xv[ks][6], sc_x, ax[0][ks][2], ax[NSPLIT - 1][ks][2]);
                split2(xv[ks][5], xv[ks][7], sc_x, ax[0][ks][3], ax[NSPLIT - 1][ks][3]);
                split2(gv[ks][0], gv[ks][2], sc_g, ag[0][ks][0], ag[NSPLIT - 1][ks][0]);
                split2(gv[ks][1], gv[ks][3], sc_g, ag[0][ks][1], ag[NSPLIT - 1][ks][1]);
                split2(gv[ks][4], gv[ks][6], sc_g, ag[0][ks][2], ag[NSPLIT - 1][ks][2]);
                split2(gv[ks][5], gv[ks][7], sc_g, ag[0][ks][3], ag[NSPLIT - 1][ks][3]);
            }
        } else {
            // ldmatrix.trans turns the [channel][pixel] tile into (pixel-row, channel-k) fragments:
            // matrices 0 = ch 0-7 / px 0-7, 1 = ch 0-7 / px 8-15, 2 = ch 8-15 / px 0-7, 3 = ch 8-15 / px 8-15
            const int mi = lane >> 3, r = lane & 7;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int off = (16 * ks + 8 * (mi >> 1) + r) * RS + 8 * (mi & 1);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(ax[0][ks][0]), "=r"(ax[0][ks][1]), "=r"(ax[0][ks][2]), "=r"(ax[0][ks][3])
                             : "r"(smem_u32(xs + off)));
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(ag[0][ks][0]), "=r"(ag[0][ks][1]), "=r"(ag[0][ks][2]), "=r"(ag[0][ks][3])
                             : "r"(smem_u32(gs + off)));
            }
        }

        // ---- S = x^T . srcT and dP = g^T . srcT ------------------------------------------------
        float s[NT][4], dp[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f;
            dp[nt][0] = 0.f; dp[nt][1] = 0.f; dp[nt][2] = 0.f; dp[nt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint2 bh2 = frag(0, ks * NT + nt);
                const uint32_t bh[2] = {bh2.x, bh2.y};
                if constexpr (HALF) {
                    const uint2 bl2 = frag(1, ks * NT + nt);
                    const uint32_t bl[2] = {bl2.x, bl2.y};
                    mma16816<true>(s[nt], ax[1][ks], bh);
                    mma16816<true>(dp[nt], ag[1][ks], bh);
                    mma16816<true>(s[nt], ax[0][ks], bl);
                    mma16816<true>(dp[nt], ag[0][ks], bl);
                }
                mma16816<HALF>(s[nt], ax[0][ks], bh);
                mma16816<HALF>(dp[nt], ag[0][ks], bh);
            }

        // ---- P = masked softmax over words (recomputed; GlobalAttention.py:104-109) ---------------
        const int q0 = t * TQ + warp * 16 + g;     // rows: pixel q0 (regs 0,1) and q0 + 8 (regs 2,3)
        uint32_t mb0 = pad_bits, mb1 = pad_bits;
        if (p.mask != nullptr) {
            if (p.mask_mode == SBA_MASK_PER_SAMPLE) {
                mb0 |= mb_s[b];
                mb1 = mb0;
            } else {
                uint32_t cap1 = cap0 + (8u % Bu);
                if (cap1 >= Bu) cap1 -= Bu;
                mb0 |= mb_s[cap0];
                mb1 |= mb_s[cap1];
            }
        }
        mb0 >>= 2 * c;
        mb1 >>= 2 * c;
        {
            const float unscale = inv_x * inv_src * kLog2e;
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[nt][j] = ((mb0 >> (8 * nt + j)) & 1u) ? -INFINITY : s[nt][j] * unscale;
                    s[nt][2 + j] = ((mb1 >> (8 * nt + j)) & 1u) ? -INFINITY : s[nt][2 + j] * unscale;
                    m0 = fmaxf(m0, s[nt][j]);
                    m1 = fmaxf(m1, s[nt][2 + j]);
                }
            m0 = quad_max(m0);
            m1 = quad_max(m1);
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[nt][j] = ex2_approx(s[nt][j] - m0);      // all-masked row: NaN, as the reference
                    s[nt][2 + j] = ex2_approx(s[nt][2 + j] - m1);
                    sum0 += s[nt][j];
                    sum1 += s[nt][2 + j];
                }
            const float inv0 = rcp_approx(quad_sum(sum0)), inv1 = rcp_approx(quad_sum(sum1));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[nt][j] *= inv0;
                    s[nt][2 + j] *= inv1;
                }
        }

        // ---- dS = P * (dP - sum_l P dP)  (softmax backward; masked / padded words have P = 0) -----
        float inv_d = 1.f;
        {
            const float unscale_g = inv_g * inv_src;
            float dot0 = 0.f, dot1 = 0.f;
            const T* ga_b = HAS_GA ? static_cast<const T*>(p.ga) + (size_t)b * L * Q + q0 : nullptr;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float d0 = dp[nt][j] * unscale_g, d1 = dp[nt][2 + j] * unscale_g;
                    if constexpr (HAS_GA) {
                        const int l = nt * 8 + 2 * c + j;
                        if (l < L) {
                            const T* gp = ga_b + (size_t)l * Q;
                            if constexpr (HALF) { d0 += __ldg(gp); d1 += __ldg(gp + 8); }
                            else { d0 += __bfloat162float(gp[0]); d1 += __bfloat162float(gp[8]); }
                        }
                    }
                    dp[nt][j] = d0;
                    dp[nt][2 + j] = d1;
                    dot0 = fmaf(s[nt][j], d0, dot0);
                    dot1 = fmaf(s[nt][2 + j], d1, dot1);
                }
            dot0 = quad_sum(dot0);
            dot1 = quad_sum(dot1);
            float amd = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dp[nt][j] = s[nt][j] * (dp[nt][j] - dot0);
                    dp[nt][2 + j] = s[nt][2 + j] * (dp[nt][2 + j] - dot1);
                    if constexpr (HALF) amd = fmaxf(amd, fmaxf(fabsf(dp[nt][j]), fabsf(dp[nt][2 + j])));
                }
            if constexpr (HALF) {
                // NaN rows (fully masked caption) must not poison the scale: fmaxf drops NaN
                float sc_d;
                pow2_scale(warp_absmax_redux(amd), sc_d, inv_d);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dp[nt][j] *= sc_d;
            }
        }
        // packed A-operand views of dS (px-major) and P
        uint32_t dsA[NSPLIT][NT][2], pA[NSPLIT][NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if constexpr (HALF) {
                split2(dp[nt][0], dp[nt][1], 1.f, dsA[0][nt][0], dsA[1][nt][0]);
                split2(dp[nt][2], dp[nt][3], 1.f, dsA[0][nt][1], dsA[1][nt][1]);
                split2(s[nt][0], s[nt][1], 1.f, pA[0][nt][0], pA[1][nt][0]);
                split2(s[nt][2], s[nt][3], 1.f, pA[0][nt][1], pA[1][nt][1]);
            } else {
                dsA[0][nt][0] = pack_bf16(dp[nt][0], dp[nt][1]);
                dsA[0][nt][1] = pack_bf16(dp[nt][2], dp[nt][3]);
                pA[0][nt][0] = pack_bf16(s[nt][0], s[nt][1]);
                pA[0][nt][1] = pack_bf16(s[nt][2], s[nt][3]);
            }
        }

        // ---- dX = dS . srcT^T --------------------------------------------------------------------
        {
            float cc[NC8][4];
#pragma unroll
            for (int nc = 0; nc < NC8; ++nc) { cc[nc][0] = 0.f; cc[nc][1] = 0.f; cc[nc][2] = 0.f; cc[nc][3] = 0.f; }
#pragma unroll
            for (int j = 0; j < NK16; ++j) {
                uint32_t ah[4] = {dsA[0][2 * j][0], dsA[0][2 * j][1], dsA[0][2 * j + 1][0], dsA[0][2 * j + 1][1]};
#pragma unroll
                for (int nc = 0; nc < NC8; ++nc) {
                    const uint2 bh2 = frag(0, C::NF_S + j * NC8 + nc);
                    const uint32_t bh[2] = {bh2.x, bh2.y};
                    if constexpr (HALF) {
                        uint32_t al[4] = {dsA[1][2 * j][0], dsA[1][2 * j][1], dsA[1][2 * j + 1][0], dsA[1][2 * j + 1][1]};
                        const uint2 bl2 = frag(1, C::NF_S + j * NC8 + nc);
                        const uint32_t bl[2] = {bl2.x, bl2.y};
                        mma16816<true>(cc[nc], al, bh);
                        mma16816<true>(cc[nc], ah, bl);
                    }
                    mma16816<HALF>(cc[nc], ah, bh);
                }
            }
            if constexpr (HAS_K8) {
#pragma unroll
                for (int e = 0; e < NC8 / 2; ++e) {
                    const uint2 bh2 = frag(0, C::NF_S + C::NF_C16 + e);
                    if constexpr (HALF) {
                        const uint2 bl2 = frag(1, C::NF_S + C::NF_C16 + e);
                        mma1688<true>(cc[2 * e], dsA[1][NT - 1][0], dsA[1][NT - 1][1], bh2.x);
                        mma1688<true>(cc[2 * e + 1], dsA[1][NT - 1][0], dsA[1][NT - 1][1], bh2.y);
                        mma1688<true>(cc[2 * e], dsA[0][NT - 1][0], dsA[0][NT - 1][1], bl2.x);
                        mma1688<true>(cc[2 * e + 1], dsA[0][NT - 1][0], dsA[0][NT - 1][1], bl2.y);
                    }
                    mma1688<HALF>(cc[2 * e], dsA[0][NT - 1][0], dsA[0][NT - 1][1], bh2.x);
                    mma1688<HALF>(cc[2 * e + 1], dsA[0][NT - 1][0], dsA[0][NT - 1][1], bh2.y);
                }
            }
            T* dx_b = static_cast<T*>(p.dX) + (size_t)b * IDF * Q + q0 + (unsigned)(2 * c) * (unsigned)Q;
            const float un = inv_d * inv_src;
            unsigned off = 0;
#pragma unroll
            for (int nc = 0; nc < NC8; ++nc) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if constexpr (HALF) {
                        dx_b[off + j * (unsigned)Q] = cc[nc][j] * un;
                        dx_b[off + j * (unsigned)Q + 8] = cc[nc][2 + j] * un;
                    } else {
                        dx_b[off + j * (unsigned)Q] = __float2bfloat16_rn(cc[nc][j]);
                        dx_b[off + j * (unsigned)Q + 8] = __float2bfloat16_rn(cc[nc][2 + j]);
                    }
                }
                off += 8u * (unsigned)Q;
            }
        }

        // ---- dSrc += g . P + x . dS  over this warp's 16 pixels (K = 16) ---------------------------
        // B operands: b0 = (k px 2c,2c+1; n word g) = transpose of the packed (px g; words 2c,2c+1) block
        uint32_t pT[NSPLIT][NT][2], dT[NSPLIT][NT][2];
#pragma unroll
        for (int sp = 0; sp < NSPLIT; ++sp)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                pT[sp][nt][0] = movmatrix_trans(pA[sp][nt][0]);
                pT[sp][nt][1] = movmatrix_trans(pA[sp][nt][1]);
                dT[sp][nt][0] = movmatrix_trans(dsA[sp][nt][0]);
                dT[sp][nt][1] = movmatrix_trans(dsA[sp][nt][1]);
            }
#pragma unroll
        for (int mt = 0; mt < KS; ++mt) {
            // channel-major A fragments (rows = 16 channels of m-tile mt, k = 16 pixels)
            uint32_t gT[NSPLIT][4], xT[NSPLIT][4];
            if constexpr (HALF) {
#pragma unroll
                for (int sp = 0; sp < NSPLIT; ++sp) {
                    gT[sp][0] = movmatrix_trans(ag[sp][mt][0]);
                    gT[sp][1] = movmatrix_trans(ag[sp][mt][2]);
                    gT[sp][2] = movmatrix_trans(ag[sp][mt][1]);
                    gT[sp][3] = movmatrix_trans(ag[sp][mt][3]);
                    xT[sp][0] = movmatrix_trans(ax[sp][mt][0]);
                    xT[sp][1] = movmatrix_trans(ax[sp][mt][2]);
                    xT[sp][2] = movmatrix_trans(ax[sp][mt][1]);
                    xT[sp][3] = movmatrix_trans(ax[sp][mt][3]);
                }
                float tmp[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    tmp[nt][0] = 0.f; tmp[nt][1] = 0.f; tmp[nt][2] = 0.f; tmp[nt][3] = 0.f;
                    mma16816<true>(tmp[nt], gT[1], pT[0][nt]);
                    mma16816<true>(tmp[nt], gT[0], pT[1][nt]);
                    mma16816<true>(tmp[nt], gT[0], pT[0][nt]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dacc[mt][nt][j] = fmaf(tmp[nt][j], inv_g, dacc[mt][nt][j]);
                const float un = inv_x * inv_d;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    tmp[nt][0] = 0.f; tmp[nt][1] = 0.f; tmp[nt][2] = 0.f; tmp[nt][3] = 0.f;
                    mma16816<true>(tmp[nt], xT[1], dT[0][nt]);
                    mma16816<true>(tmp[nt], xT[0], dT[1][nt]);
                    mma16816<true>(tmp[nt], xT[0], dT[0][nt]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dacc[mt][nt][j] = fmaf(tmp[nt][j], un, dacc[mt][nt][j]);
            } else {
                // non-transposed ldmatrix: matrices 0 = ch 0-7 / px 0-7, 1 = ch 8-15 / px 0-7,
                // 2 = ch 0-7 / px 8-15, 3 = ch 8-15 / px 8-15  (= a0..a3 of the channel-major tile)
                const int mi = lane >> 3, r = lane & 7;
                const int off = (16 * mt + 8 * (mi & 1) + r) * RS + 8 * (mi >> 1);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(gT[0][0]), "=r"(gT[0][1]), "=r"(gT[0][2]), "=r"(gT[0][3])
                             : "r"(smem_u32(gs + off)));
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(xT[0][0]), "=r"(xT[0][1]), "=r"(xT[0][2]), "=r"(xT[0][3])
                             : "r"(smem_u32(xs + off)));
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma16816<false>(dacc[mt][nt], gT[0], pT[0][nt]);
                    mma16816<false>(dacc[mt][nt], xT[0], dT[0][nt]);
                }
            }
        }
        // Release the stage behind the dX stores (a release cannot be hoisted above them and they
        // depend on every fragment load of the tile); see the note in attn_mma_fwd.cu.
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));

        if (++t == TPS) { t = 0; ++b; }
        if (++stage == NST) { stage = 0; phase ^= 1; }
        cap0 += step_mod;
        if (cap0 >= Bu) cap0 -= Bu;
    }
    if (cur_b >= 0) finish_sample(cur_b);
}

template <typename T, int IDF, int NT, bool HAS_GA>
int launch_bwd_mma(const BwdParams& p0, cudaStream_t st) {
    using C = BwdCfg<T, IDF, NT>;
    BwdParams p = p0;
    auto kern = k_attn_bwd_mma<T, IDF, NT, HAS_GA>;
    constexpr int kSmemCap = 112 * 1024;     // two CTAs per SM
    static int max_ctas = 0;
    if (max_ctas == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, kSmemCap);
        if (e != cudaSuccess || per_sm < 1 || sms < 1) {
            set_error("attn_bwd(mma): occupancy query failed: %s", cudaGetErrorString(e));
            return SBA_ERR_CUDA;
        }
        max_ctas = sms * (per_sm > 2 ? 2 : per_sm);
    }
    const size_t fixed = (size_t)C::SCR_BYTES + (size_t)p.B * 4 + 16;
    int nst = fixed < (size_t)kSmemCap ? (int)(((size_t)kSmemCap - fixed) / C::STAGE_BYTES) : 0;
    if (nst > kMaxStages) nst = kMaxStages;
    if (nst < 2) {
        set_error("attn_bwd(mma): shared memory budget exceeded (B=%d idf=%d)", p.B, IDF);
        return SBA_ERR_UNSUPPORTED;
    }
    p.nst = nst;
    int rc = attn_bwd_zero(p.dSrc, (size_t)p.B * IDF * p.L + p.B + 1, p.dW, p.dW ? (size_t)IDF * p.cdf : 0, st);
    if (rc) return rc;
    const int grid = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemCap;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) {
        set_error("attn_bwd(mma): launch: %s", cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    add_launches(1);
    rc = check_launch("attn_bwd(mma)");
    if (rc) return rc;
    return attn_bwd_post(p.dSrc, p.ctx, p.W, p.dW, p.dCtx, p.B, IDF, p.cdf, p.L, st);
}

template <typename T, int IDF, bool HAS_GA>
int dispatch_nt(const BwdParams& p, int NT, cudaStream_t st) {
    switch (NT) {
        case 2: return launch_bwd_mma<T, IDF, 2, HAS_GA>(p, st);
        case 3: return launch_bwd_mma<T, IDF, 3, HAS_GA>(p, st);
        case 4: return launch_bwd_mma<T, IDF, 4, HAS_GA>(p, st);
        default: return -1;
    }
}

template <typename T, int IDF>
int dispatch_ga(const BwdParams& p, int NT, cudaStream_t st) {
    return p.ga != nullptr ? dispatch_nt<T, IDF, true>(p, NT, st) : dispatch_nt<T, IDF, false>(p, NT, st);
}

}  // namespace

int mma_attn_bwd(const void* x, const float* ctx, const float* W, const float* srcT, const uint8_t* mask, const void* g_c,
                 const void* g_attn, void* dX, float* dSrc, float* dW, float* dCtx, const AttnShape& s, cudaStream_t st) {
    BwdParams p{};
    p.x = x; p.g = g_c; p.ga = g_attn; p.srcT = srcT; p.mask = mask; p.dX = dX; p.dSrc = dSrc;
    p.ctx = ctx; p.W = W; p.dW = dW; p.dCtx = dCtx;
    p.cnt = reinterpret_cast<uint32_t*>(dSrc + (size_t)s.B * s.idf * s.L);
    p.B = s.B; p.L = s.L; p.Q = s.Q; p.cdf = s.cdf; p.mask_mode = s.mask_mode;
    p.tiles_per_sample = s.Q / mma::TQ;
    p.n_tiles = s.B * p.tiles_per_sample;
    const int NT = (s.L + 7) / 8;
    int rc = -1;
    if (s.dtype == SBA_F32) {
        if (s.idf == 32) rc = dispatch_ga<float, 32>(p, NT, st);
        else if (s.idf == 48) rc = dispatch_ga<float, 48>(p, NT, st);
    } else {
        if (s.idf == 32) rc = dispatch_ga<__nv_bfloat16, 32>(p, NT, st);
        else if (s.idf == 48) rc = dispatch_ga<__nv_bfloat16, 48>(p, NT, st);
    }
    if (rc == -1) {
        set_error("attn_bwd(mma): unsupported shape idf=%d L=%d", s.idf, s.L);
        return SBA_ERR_UNSUPPORTED;
    }
    return rc;
}

}  // namespace sba

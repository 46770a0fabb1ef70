// Kernel (a): fused GlobalAttentionGeneral forward on tensor cores (SBA_ALGO_MMA).
//
// One persistent CTA = 8 consumer warps + 1 TMA producer warp.  The producer streams
// [idf x 128-pixel] tiles of x into a 4-stage shared-memory ring with 1-D bulk copies (one per
// channel row, 512 B fp32 / 256 B bf16, completion on an mbarrier).  Each consumer warp owns a
// 16-pixel m-tile of the current tile and does, entirely in registers:
//     S = x^T . srcT          (m16n8k16 MMAs, B operand = sourceT fragments kept in registers)
//     masked softmax over words (quad shuffles), store attn
//     c = P . srcT^T          (P re-used as the A operand straight from the accumulator layout)
// The 1x1 conv_context projection srcT = W.ctx[b] is computed by the CTA itself whenever its
// tile range enters a new sample (and written out once per sample for the backward).
// fp32 tensors: operands are split into fp16 hi+lo with a per-tile power-of-two scale and
// multiplied with three MMAs; bf16 tensors: single bf16 MMAs.
//
// Reference semantics: AttnGAN2/code/GlobalAttention.py:82-121 (oracle/attention.py).
#include "kernels.h"
#include "mma_common.cuh"

namespace sba {
namespace {
using namespace mma;

struct FwdParams {
    const void* x;
    const float* ctx;
    const float* W;
    const uint8_t* mask;
    void* c_code;
    void* attn;
    float* srcT;
    uint32_t* mask_bits;
    int B, cdf, L, Q, mask_mode;
    int tiles_per_sample;
    int n_tiles;
};

template <typename T, int IDF, int NT>
struct FwdCfg {
    static constexpr int KS = IDF / 16;    // k-steps over channels
    static constexpr int NC8 = IDF / 8;    // n-tiles over channels (second contraction)
    static constexpr int NK16 = NT / 2;    // k16 steps over words (second contraction)
    static constexpr bool HAS_K8 = (NT % 2) != 0;
    static constexpr int LP = NT * 8;      // padded word count
    static constexpr int RS = TileStride<T>::value;
    static constexpr int NST = IDF <= 32 ? 4 : 3;   // ring depth: keep two CTAs per SM
    static constexpr int STAGE_BYTES = IDF * RS * (int)sizeof(T);
    static constexpr bool HALF = sizeof(T) == 4;   // fp32 tensors -> split fp16 MMAs
};

// sourceT fragments of one sample, resident in registers for the whole sample
template <int KS, int NT, int NC8, int NK16, bool HAS_K8, int NSPLIT>
struct SrcFrags {
    uint32_t bs[NSPLIT][KS][NT][2];                      // S-phase  B: k = channel, n = word
    uint32_t bc16[NSPLIT][NK16 > 0 ? NK16 : 1][NC8][2];  // c-phase  B: k = word (16), n = channel
    uint32_t bc8[NSPLIT][NC8];                           // c-phase  B: k = word (last 8)
};

// Build the register-resident sourceT fragments of sample b from global memory (written by
// phase 0 of this launch, possibly by another SM: L1 is bypassed with __ldcg).
template <typename T, int IDF, int NT, class Frags>
__device__ __forceinline__ void load_src_frags(Frags& f, const float* __restrict__ srcT_b, int L, float sc_src, int g, int c) {
    using C = FwdCfg<T, IDF, NT>;
    constexpr int KS = C::KS, NC8 = C::NC8, NK16 = C::NK16;
    constexpr bool HAS_K8 = C::HAS_K8, HALF = C::HALF;
    constexpr int NSPLIT = HALF ? 2 : 1;
    auto ld = [&](int ch, int l) -> float { return (l < L) ? __ldcg(srcT_b + ch * L + l) : 0.f; };
    // m16n8k16 B fragment: b0 = (k 2c,2c+1; n g), b1 = (k 2c+8,2c+9; n g)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int ch0, ch1;
                if (HALF) { ch0 = chan_of(ks, c, 2 * h); ch1 = chan_of(ks, c, 2 * h + 1); }
                else { ch0 = 16 * ks + 8 * h + 2 * c; ch1 = ch0 + 1; }
                const float v0 = ld(ch0, nt * 8 + g), v1 = ld(ch1, nt * 8 + g);
                if (HALF) split2(v0, v1, sc_src, f.bs[0][ks][nt][h], f.bs[NSPLIT - 1][ks][nt][h]);
                else f.bs[0][ks][nt][h] = pack_bf16(v0, v1);
            }
#pragma unroll
    for (int nc = 0; nc < NC8; ++nc) {
        const int ch = nc * 8 + g;
#pragma unroll
        for (int j = 0; j < NK16; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l0 = 16 * j + 8 * h + 2 * c;
                const float v0 = ld(ch, l0), v1 = ld(ch, l0 + 1);
                if (HALF) split2(v0, v1, sc_src, f.bc16[0][j][nc][h], f.bc16[NSPLIT - 1][j][nc][h]);
                else f.bc16[0][j][nc][h] = pack_bf16(v0, v1);
            }
        if (HAS_K8) {
            const int l0 = (NT - 1) * 8 + 2 * c;
            const float v0 = ld(ch, l0), v1 = ld(ch, l0 + 1);
            if (HALF) split2(v0, v1, sc_src, f.bc8[0][nc], f.bc8[NSPLIT - 1][nc]);
            else f.bc8[0][nc] = pack_bf16(v0, v1);
        }
    }
}

// srcT = W . ctx (the bias-free 1x1 conv_context, GlobalAttention.py:95-97) as its own small
// grid, launched in front of the streaming kernel with programmatic dependent launch: block u =
// (sample, group of 8 output channels), one warp per channel, lanes = 8 words x 4 quarters of
// the cdf reduction.  The streaming kernel starts prefetching x tiles while this runs and only
// its consumers wait (griddepcontrol.wait) before they read srcT.
template <int IDF, int NT>
__global__ void __launch_bounds__(256) k_project_mma(const float* __restrict__ ctx, const float* __restrict__ W,
                                                     float* __restrict__ srcT, int cdf, int L) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lg = lane & 7, kq = lane >> 3;
    constexpr int RG = IDF / 8;
    const int u = blockIdx.x, b = u / RG, i = (u - b * RG) * 8 + warp;
    const float* wrow = W + (size_t)i * cdf;
    const float* cb = ctx + (size_t)b * cdf * L;
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
    const int c_lo = (cdf * kq) >> 2, c_hi = (cdf * (kq + 1)) >> 2;
#pragma unroll 8
    for (int cc = c_lo; cc < c_hi; ++cc) {
        const float wv = __ldg(wrow + cc);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int l = lg + 8 * n;
            const float v = (l < L) ? __ldg(cb + (size_t)cc * L + l) : 0.f;
            acc[n] = fmaf(wv, v, acc[n]);
        }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 8);
        acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 16);
        const int l = lg + 8 * n;
        if (kq == 0 && l < L) srcT[((size_t)b * IDF + i) * L + l] = acc[n];
    }
}

template <typename T, int IDF, int NT>
__global__ void __launch_bounds__(kThreads, 2) k_attn_fwd_mma(const FwdParams p) {
    using C = FwdCfg<T, IDF, NT>;
    constexpr int KS = C::KS, NC8 = C::NC8, NK16 = C::NK16, RS = C::RS, NST = C::NST;
    constexpr bool HAS_K8 = C::HAS_K8, HALF = C::HALF;
    constexpr int NSPLIT = HALF ? 2 : 1;
    constexpr float kLog2e = 1.4426950408889634f;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* stages = reinterpret_cast<T*>(smem_raw);
    uint32_t* mb_s = reinterpret_cast<uint32_t*>(smem_raw + (size_t)NST * C::STAGE_BYTES);   // [B]
    __shared__ __align__(8) unsigned long long bar_full[NST], bar_empty[NST];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    const int L = p.L, Q = p.Q, TPS = p.tiles_per_sample;

    if (tid == 0) {
#pragma unroll
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
            if (blockIdx.x == 0) p.mask_bits[cap] = bits;
        }
    }
    __syncthreads();

    const int w_begin = (int)(((long long)blockIdx.x * p.n_tiles) / gridDim.x);
    const int w_end = (int)(((long long)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    int b = w_begin / TPS, t = w_begin - b * TPS;

    if (warp == kConsumerWarps) {
        // ------------------------------ TMA producer warp ------------------------------------
        const T* xg = static_cast<const T*>(p.x);
        int stage = 0, phase = 0;
        for (int w = w_begin; w < w_end; ++w) {
            if (w - w_begin >= NST) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
            const uint32_t full = smem_u32(&bar_full[stage]);
            if (lane == 0) mbar_expect_tx(full, IDF * TQ * (uint32_t)sizeof(T));
            __syncwarp();
            const uint32_t dst0 = smem_u32(stages) + stage * C::STAGE_BYTES;
            for (int r = lane; r < IDF; r += 32)
                tma_load_1d(dst0 + r * RS * (uint32_t)sizeof(T), xg + ((size_t)b * IDF + r) * Q + (size_t)t * TQ,
                            TQ * (uint32_t)sizeof(T), full);
            if (++t == TPS) { t = 0; ++b; }
            if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        return;
    }

    // ---------------------------------- consumer warps ---------------------------------------
    asm volatile("griddepcontrol.wait;" ::: "memory");      // srcT of k_project_mma is complete and visible

    SrcFrags<KS, NT, NC8, NK16, HAS_K8, NSPLIT> f;
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
            cur_b = b;
            float sc_src = 1.f;
            if (HALF) {
                const float* sb = p.srcT + (size_t)b * IDF * L;
                float lm = 0.f;
                for (int o = lane; o < IDF * L; o += 32) lm = fmaxf(lm, fabsf(__ldcg(sb + o)));
                pow2_scale(warp_absmax_redux(lm), sc_src, inv_src);
            }
            load_src_frags<T, IDF, NT>(f, p.srcT + (size_t)b * IDF * L, L, sc_src, g, c);
        }

        // ---- A fragments of this warp's 16 pixels from the staged tile ------------------------
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        const T* xs = reinterpret_cast<const T*>(smem_raw + (size_t)stage * C::STAGE_BYTES) + warp * 16;
        uint32_t a_hi[KS][4], a_lo[HALF ? KS : 1][4];
        float inv_x = 1.f;
        if constexpr (HALF) {
            float xv[KS][8];
            float amax = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* rowp = reinterpret_cast<const float*>(xs) + chan_of(ks, c, j) * RS + g;
                    xv[ks][2 * j] = rowp[0];        // pixel g
                    xv[ks][2 * j + 1] = rowp[8];    // pixel g + 8
                    amax = fmaxf(amax, fmaxf(fabsf(xv[ks][2 * j]), fabsf(xv[ks][2 * j + 1])));
                }
            amax = warp_absmax_redux(amax);
            float sc_x;
            pow2_scale(amax, sc_x, inv_x);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                // a0 = (row g; k 2c,2c+1) a1 = (row g+8; same k) a2 = (row g; k 2c+8,2c+9) a3 = (row g+8; ...)
                split2(xv[ks][0], xv[ks][2], sc_x, a_hi[ks][0], a_lo[ks][0]);
                split2(xv[ks][1], xv[ks][3], sc_x, a_hi[ks][1], a_lo[ks][1]);
                split2(xv[ks][4], xv[ks][6], sc_x, a_hi[ks][2], a_lo[ks][2]);
                split2(xv[ks][5], xv[ks][7], sc_x, a_hi[ks][3], a_lo[ks][3]);
            }
        } else {
            // bf16: ldmatrix.trans turns the [channel][pixel] tile into (pixel-row, channel-k) fragments
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                // matrices: 0 = ch 0-7 / px 0-7, 1 = ch 0-7 / px 8-15, 2 = ch 8-15 / px 0-7, 3 = ch 8-15 / px 8-15
                const int mi = lane >> 3, r = lane & 7;
                const T* rowp = xs + (16 * ks + 8 * (mi >> 1) + r) * RS + 8 * (mi & 1);
                const uint32_t addr = smem_u32(rowp);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a_hi[ks][0]), "=r"(a_hi[ks][1]), "=r"(a_hi[ks][2]), "=r"(a_hi[ks][3])
                             : "r"(addr));
            }
        }

        // ---- S = x^T . srcT  (GlobalAttention.py:102) ------------------------------------------
        float s[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
        if constexpr (HALF) {
            // small cross terms first, every accumulator touched once per round (keeps 3 MMAs in flight)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma16816<true>(s[nt], a_lo[ks], f.bs[0][ks][nt]);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma16816<true>(s[nt], a_hi[ks], f.bs[1][ks][nt]);
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma16816<HALF>(s[nt], a_hi[ks], f.bs[0][ks][nt]);

        // ---- mask (GlobalAttention.py:104-108) + softmax over words (:109), in the log2 domain ----
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
                s[nt][j] = ex2_approx(s[nt][j] - m0);      // all-masked row: -inf - -inf = NaN, as the reference
                s[nt][2 + j] = ex2_approx(s[nt][2 + j] - m1);
                sum0 += s[nt][j];
                sum1 += s[nt][2 + j];
            }
        const float inv0 = rcp_approx(quad_sum(sum0)), inv1 = rcp_approx(quad_sum(sum1));
        // row offsets as 32-bit element indices (mma_supports guarantees L*Q and idf*Q < 2^31)
        T* attn_b = static_cast<T*>(p.attn) + (size_t)b * L * Q + q0 + (unsigned)(2 * c) * (unsigned)Q;
        {
            unsigned off = 0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[nt][j] *= inv0;
                    s[nt][2 + j] *= inv1;
                    const int l = nt * 8 + 2 * c + j;
                    if (nt < NT - 1 || l < L) {
                        if constexpr (HALF) {
                            attn_b[off + j * (unsigned)Q] = s[nt][j];
                            attn_b[off + j * (unsigned)Q + 8] = s[nt][2 + j];
                        } else {
                            attn_b[off + j * (unsigned)Q] = __float2bfloat16_rn(s[nt][j]);
                            attn_b[off + j * (unsigned)Q + 8] = __float2bfloat16_rn(s[nt][2 + j]);
                        }
                    }
                }
                off += 8u * (unsigned)Q;
            }
        }

        // ---- c = P . srcT^T  (GlobalAttention.py:117): accumulator layout == A-fragment layout ----
        uint32_t pa_hi[NT][2], pa_lo[HALF ? NT : 1][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if constexpr (HALF) {
                split2(s[nt][0], s[nt][1], 1.f, pa_hi[nt][0], pa_lo[nt][0]);
                split2(s[nt][2], s[nt][3], 1.f, pa_hi[nt][1], pa_lo[nt][1]);
            } else {
                pa_hi[nt][0] = pack_bf16(s[nt][0], s[nt][1]);
                pa_hi[nt][1] = pack_bf16(s[nt][2], s[nt][3]);
            }
        }
        float cc[NC8][4];
#pragma unroll
        for (int nc = 0; nc < NC8; ++nc) { cc[nc][0] = 0.f; cc[nc][1] = 0.f; cc[nc][2] = 0.f; cc[nc][3] = 0.f; }
#pragma unroll
        for (int sp = (HALF ? 0 : 2); sp < 3; ++sp) {
            // sp 0: lo x hi, 1: hi x lo, 2: hi x hi
            const int bi = (sp == 1) ? NSPLIT - 1 : 0;
#pragma unroll
            for (int j = 0; j < NK16; ++j) {
                uint32_t a[4];
                if (sp == 0 && HALF) { a[0] = pa_lo[2 * j][0]; a[1] = pa_lo[2 * j][1]; a[2] = pa_lo[2 * j + 1][0]; a[3] = pa_lo[2 * j + 1][1]; }
                else { a[0] = pa_hi[2 * j][0]; a[1] = pa_hi[2 * j][1]; a[2] = pa_hi[2 * j + 1][0]; a[3] = pa_hi[2 * j + 1][1]; }
#pragma unroll
                for (int nc = 0; nc < NC8; ++nc) mma16816<HALF>(cc[nc], a, f.bc16[bi][j][nc]);
            }
            if constexpr (HAS_K8) {
                const uint32_t a0 = (sp == 0 && HALF) ? pa_lo[HALF ? NT - 1 : 0][0] : pa_hi[NT - 1][0];
                const uint32_t a1 = (sp == 0 && HALF) ? pa_lo[HALF ? NT - 1 : 0][1] : pa_hi[NT - 1][1];
#pragma unroll
                for (int nc = 0; nc < NC8; ++nc) mma1688<HALF>(cc[nc], a0, a1, f.bc8[bi][nc]);
            }
        }
        T* c_b = static_cast<T*>(p.c_code) + (size_t)b * IDF * Q + q0 + (unsigned)(2 * c) * (unsigned)Q;
        {
            unsigned off = 0;
#pragma unroll
            for (int nc = 0; nc < NC8; ++nc) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if constexpr (HALF) {
                        c_b[off + j * (unsigned)Q] = cc[nc][j] * inv_src;
                        c_b[off + j * (unsigned)Q + 8] = cc[nc][2 + j] * inv_src;
                    } else {
                        c_b[off + j * (unsigned)Q] = __float2bfloat16_rn(cc[nc][j]);
                        c_b[off + j * (unsigned)Q + 8] = __float2bfloat16_rn(cc[nc][2 + j]);
                    }
                }
                off += 8u * (unsigned)Q;
            }
        }

        // Release the stage only here, behind the output stores: mbarrier.arrive is a release, so
        // ptxas cannot hoist it above them, and they depend on every fragment load of the tile.
        // (Arriving right after the ldmatrix / LDS instructions were ISSUED let the producer's next
        // bulk copy overwrite the stage while loads of a slow warp were still queued.)
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));

        if (++t == TPS) { t = 0; ++b; }
        if (++stage == NST) { stage = 0; phase ^= 1; }
        cap0 += step_mod;
        if (cap0 >= Bu) cap0 -= Bu;
    }
}

template <typename T, int IDF, int NT>
int launch_fwd_mma(const FwdParams& p, cudaStream_t st) {
    using C = FwdCfg<T, IDF, NT>;
    const size_t smem = (size_t)C::NST * C::STAGE_BYTES + (size_t)p.B * 4 + 16;
    auto kern = k_attn_fwd_mma<T, IDF, NT>;
    static int max_ctas = 0;     // one persistent wave: two CTAs per SM
    if (max_ctas == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 100 * 1024);
        if (e != cudaSuccess || per_sm < 1 || sms < 1) {
            set_error("attn_fwd(mma): occupancy query failed: %s", cudaGetErrorString(e));
            return SBA_ERR_CUDA;
        }
        max_ctas = sms * (per_sm > 2 ? 2 : per_sm);
    }
    if (smem > 100 * 1024) {
        set_error("attn_fwd(mma): %zu bytes of shared memory needed (B=%d)", smem, p.B);
        return SBA_ERR_UNSUPPORTED;
    }
    k_project_mma<IDF, NT><<<p.B * (IDF / 8), 256, 0, st>>>(p.ctx, p.W, p.srcT, p.cdf, p.L);
    int rc = check_launch("project(mma)");
    if (rc) return rc;
    const int grid = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 100 * 1024;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) {
        set_error("attn_fwd(mma): launch: %s", cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    add_launches(2);
    return check_launch("attn_fwd(mma)");
}

template <typename T, int IDF>
int dispatch_nt(const FwdParams& p, int NT, cudaStream_t st) {
    switch (NT) {
        case 2: return launch_fwd_mma<T, IDF, 2>(p, st);
        case 3: return launch_fwd_mma<T, IDF, 3>(p, st);
        case 4: return launch_fwd_mma<T, IDF, 4>(p, st);
        default: return -1;
    }
}

}  // namespace

bool mma_supports(const AttnShape& s) {
    if (s.idf != 32 && s.idf != 48) return false;
    if (s.L < 9 || s.L > 32) return false;
    if (s.Q % mma::TQ != 0) return false;
    if (s.cdf % 4 != 0) return false;
    if (s.B > 2048 || (unsigned long long)s.B * s.Q >= (1ull << 31)) return false;
    if ((unsigned long long)s.Q * 64 >= (1ull << 31)) return false;   // 32-bit row offsets inside one sample
    return true;
}

int mma_attn_fwd(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                 float* srcT, uint32_t* mask_bits, const AttnShape& s, cudaStream_t st) {
    FwdParams p{};
    p.x = x; p.ctx = ctx; p.W = W; p.mask = mask; p.c_code = c_code; p.attn = attn; p.srcT = srcT; p.mask_bits = mask_bits;
    p.B = s.B; p.cdf = s.cdf; p.L = s.L; p.Q = s.Q; p.mask_mode = s.mask_mode;
    p.tiles_per_sample = s.Q / mma::TQ;
    p.n_tiles = s.B * p.tiles_per_sample;
    const int NT = (s.L + 7) / 8;
    int rc = -1;
    if (s.dtype == SBA_F32) {
        if (s.idf == 32) rc = dispatch_nt<float, 32>(p, NT, st);
        else if (s.idf == 48) rc = dispatch_nt<float, 48>(p, NT, st);
    } else {
        if (s.idf == 32) rc = dispatch_nt<__nv_bfloat16, 32>(p, NT, st);
        else if (s.idf == 48) rc = dispatch_nt<__nv_bfloat16, 48>(p, NT, st);
    }
    if (rc == -1) {
        set_error("attn_fwd(mma): unsupported shape idf=%d L=%d", s.idf, s.L);
        return SBA_ERR_UNSUPPORTED;
    }
    return rc;
}

}  // namespace sba

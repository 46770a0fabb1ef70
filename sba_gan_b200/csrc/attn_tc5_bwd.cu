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
//   second stage: dX row of the pixel -> staged -> TMA box store; when the CTA leaves a sample, the two
//                 diagonal blocks of the TMEM accumulator are added (g.P + x.dS, through shared memory)
//                 and stored to this CTA's PARTIAL SLOT of the sample: slot = CTA index + sample index
//                 (unique because the CTAs' tile ranges are contiguous and ordered).
// Two {S, dP} TMEM buffers and two PB buffers form a software pipeline: MMA1(j+1) is issued before
// MMA2(j), so the first stage of tile j+1 overlaps the tensor core and the second stage of tile j.
// At a sample boundary the pipeline drains (the operands are rebuilt only after the last MMA of the
// old sample has retired, the accumulator is handed back after it has been read).
// A small kernel behind it (k_bwd_finish_tc5, a programmatic dependent) adds the slots of every sample in CTA
// order -> dSrc[b], forms dW = sum_b dSrc[b] . ctx[b]^T
// over 16 sample groups whose partial products the last-arriving block of each column slice adds in group
// order, and dCtx[b] = W^T . dSrc[b].  Nothing is zero-filled and no value is accumulated with atomics
// (counters only decide WHO adds, every sum has a fixed order): the whole backward is two kernels and
// bit-reproducible run to run (data-parallel replicas stay identical).
#include "host_util.h"
#include "kernels.h"
#include "tc5_common.cuh"

namespace sba {
namespace {
using namespace tc5;

struct Tc5BwdParams {
    const float* srcT;
    const uint8_t* mask;
    const uint32_t* mask_bits;   // [B] caption mask words written by the forward (scratch)
    const void* ga;       // [B, L, Q] nullable
    float* part;          // [n_ctas + B][idf][LP] partial dSrc sums: slot = CTA index + sample index
    uint32_t* counters;   // completion counters of the finish kernel, zeroed here by CTA 0
    int n_counters;
    int B, L, Q, mask_mode;
    int g_rows, g_row0;   // g_c lives in rows [g_row0, g_row0 + idf) of a [B, g_rows, Q] buffer
    int tiles_per_sample;
    int n_tiles;
    unsigned long long* tl;   // development timeline stamps (tc5_common.cuh), NULL in the product library
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
    static constexpr int XCHG_BYTES = IDF * LP * 4;             // x.dS block on its way to the g.P rows (sample flush)
    static constexpr int SMEM_BYTES =
        NST * STAGE_BYTES + 2 * PB_BYTES + B1_BYTES + B2_BYTES + 4 * OUT_WARP_BYTES + XCHG_BYTES;
    static constexpr uint32_t IDESC1 = make_idesc(1, 1, 0, TQ, NS);      // A = tile, MN-major
    static constexpr uint32_t IDESC2 = make_idesc(1, 1, 0, TQ, IDF);     // A = dS rows of PB, MN-major
    static constexpr uint32_t IDESC3 = make_idesc(1, 0, 0, MD, ND);      // A = [g ; x], B = PB, both K-major
    static constexpr int CTAS_PER_SM = 512 / TMEM_COLS;
    static_assert(IDF % 16 == 0 && 2 * IDF <= 128, "idf must be a multiple of 16, at most 64");
    static_assert(LP <= 32 && IDF <= 64, "at most 32 words");
};

// ------------------------------------------------------------------------------------------------
// Finish kernel: dSrc[b] = sum of the sample's partial slots (CTA order), dW, dCtx.
//   blocks [0, n_dw)        : (slice cg of 32 input channels c, sample group grp), 1024 threads, rounds of up to 4 samples.
//                             thread = (sample kq of the round, output channel i, 4 input channels c):
//                               before griddepcontrol.wait (ctx does not depend on the streaming kernel):
//                                 cs [(sample, word)][32 c] = ctx^T of the first round, one 16-byte load per thread;
//                               after it: ds [sample][i][LP] = slot sums, one batch of up to 8 independent 16-byte loads
//                                 per (sample, i, 4 words) item; L FMAs x 4 per thread (one broadcast LDS + one LDS.128
//                                 per 4 FMAs); the 4 samples of a round meet in shared memory in sample order.
//                             The group's partial [idf x 32] goes to dwp[cg][grp]; the block that arrives last at the
//                             slice's counter adds the `groups` partials in group order (4 loads per thread).
//   blocks [n_dw, n_dw + B) : dCtx of one sample (only when words need a gradient)
// This kernel runs once per call on an otherwise idle GPU, so nothing hides latency: every dependent step (global
// round trip, barrier, cold instruction fetch) is exposed and instructions cost ~2.5 cycles each per thread at 8
// warps per SM.  Hence many threads with a few dozen instructions each and exactly one batch of loads per phase.
// (Measured on the way here at B=64: 256-thread blocks with the same phases 9.7 us; rolled loops 15 us.)
// ------------------------------------------------------------------------------------------------
struct Tc5FinishParams {
    const float* part;    // [n_ctas + B][idf][LP]
    const float* ctx;     // [B, cdf, L]
    const float* W;       // [idf, cdf]   (dCtx only)
    float* dSrc;          // [B, idf, L] out
    float* dW;            // [idf, cdf] out, nullable
    float* dCtx;          // [B, cdf, L] out, nullable
    float* dwp;           // [cslices][groups][idf][32] group partials of dW
    uint32_t* counters;   // [1 + cslices], zeroed by the streaming kernel
    int B, cdf, L, LP;
    int tiles_per_sample, n_tiles, n_ctas;    // tile -> CTA map of the streaming kernel
    int groups, n_dw;
    int ctx_aligned;      // ctx is 16-byte aligned (vector loads)
    unsigned long long* tl;
};

// CTA k of the streaming kernel owns tiles [floor(k n / G), floor((k + 1) n / G)): the CTA that owns tile t
__device__ __forceinline__ int tile_owner(int t, int n_tiles, int n_ctas) {
    // 32-bit when (t + 1) * n_ctas cannot overflow (n_tiles * n_ctas < 2^32: always, short of 14 M tiles)
    if ((unsigned long long)n_tiles * (unsigned)n_ctas < (1ull << 32))
        return (int)((((unsigned)t + 1u) * (unsigned)n_ctas - 1u) / (unsigned)n_tiles);
    return (int)((((long long)t + 1) * n_ctas - 1) / n_tiles);
}

constexpr int kFinThreads = 1024;
constexpr int kFinRound = 4;       // samples per round = thread groups of 256
constexpr int kFinCS = 36;         // row stride (floats) of cs: 16-byte aligned rows, 4-bank skew
constexpr int kFinSlots = 8;       // slots loaded in one batch (more: further batches)
constexpr int kFinGroups = 64;     // at most this many sample groups (their partials are added by the last block)

template <int IDF>
__global__ void __launch_bounds__(kFinThreads) k_bwd_finish_tc5(const Tc5FinishParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ int s_last;
    constexpr int NH = IDF / 32 + (IDF % 32 != 0);     // output channels per thread: i0, i0 + 32
    const int tid = threadIdx.x;
    const int B = p.B, cdf = p.cdf, L = p.L, LP = p.LP, TPS = p.tiles_per_sample;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // (the head kernel of the next call)
    if (tid == 0) { SBA_TL(p.tl, 8); SBA_TL(p.tl, 9); }
    if ((int)blockIdx.x < p.n_dw) {
        float* cs = sm;                                   // [kFinRound * 32][kFinCS]  ctx^T of the round: row = (sample, word)
        float* ds = sm + kFinRound * 32 * kFinCS;         // [kFinRound][IDF][LP]      slot sums; later the 4 samples' products
        const int cg = blockIdx.x / p.groups, grp = blockIdx.x - cg * p.groups;
        const int c0 = cg * 32, nc = cdf - c0 < 32 ? cdf - c0 : 32;
        const int b_lo = B * grp / p.groups, b_hi = B * (grp + 1) / p.groups;
        const int kq = tid >> 8, i0 = (tid >> 3) & 31, cq = (tid & 7) * 4;
        const int LP4 = LP >> 2;
        const float inv_L = 1.0f / (float)L;
        const int cvec = (nc * L) >> 2;                   // float4s of one sample's [nc][L] ctx block
        const bool cvec_ok = ((nc * L) & 3) == 0 && p.ctx_aligned;
        float acc[NH][4];
#pragma unroll
        for (int h = 0; h < NH; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;

        auto stage_ctx = [&](int bb, int nb) {            // cs[(s, l)][c] = ctx[bb + s][c0 + c][l]
            if (cvec_ok) {
                for (int f = tid; f < nb * cvec; f += kFinThreads) {
                    const int sidx = f / cvec, r = f - sidx * cvec;
                    const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.ctx + ((size_t)(bb + sidx) * cdf + c0) * L) + r);
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int el = 4 * r + e;                     // element of the [nc][L] block
                        const int c = (int)(((float)el + 0.5f) * inv_L), l = el - c * L;
                        cs[(sidx * L + l) * kFinCS + c] = v[e];
                    }
                }
            } else {
                for (int el = tid; el < nb * nc * L; el += kFinThreads) {
                    const int sidx = el / (nc * L), r = el - sidx * nc * L;
                    const int c = r / L, l = r - c * L;
                    cs[(sidx * L + l) * kFinCS + c] = __ldg(p.ctx + ((size_t)(bb + sidx) * cdf + c0) * L + r);
                }
            }
        };
        stage_ctx(b_lo, b_hi - b_lo < kFinRound ? b_hi - b_lo : kFinRound);      // ahead of the grid dependency
        asm volatile("griddepcontrol.wait;" ::: "memory");          // every slot is complete and visible
        if (tid == 0) { SBA_TL(p.tl, 4); SBA_TL(p.tl, 10); SBA_TL(p.tl, 11); }
        for (int bb = b_lo; bb < b_hi; bb += kFinRound) {
            const int nb = b_hi - bb < kFinRound ? b_hi - bb : kFinRound;
            if (bb > b_lo) {
                __syncthreads();                         // the previous round is done with cs / ds
                stage_ctx(bb, nb);
            }
            // slot sums: item = (sample s, channel i, float4 q), one batch of up to 8 loads each
            for (int item = tid; item < nb * IDF * LP4; item += kFinThreads) {
                const int row = item / LP4, q = item - row * LP4;
                const int sidx = row / IDF, i = row - sidx * IDF;
                const int b = bb + sidx;
                const int k_lo = tile_owner(b * TPS, p.n_tiles, p.n_ctas), k_hi = tile_owner((b + 1) * TPS - 1, p.n_tiles, p.n_ctas);
                const float4* sp = reinterpret_cast<const float4*>(p.part + ((size_t)(k_lo + b) * IDF + i) * LP) + q;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k0 = k_lo; k0 <= k_hi; k0 += kFinSlots, sp += (size_t)kFinSlots * IDF * LP4) {
                    float4 v[kFinSlots];
#pragma unroll
                    for (int u = 0; u < kFinSlots; ++u)
                        v[u] = (k0 + u <= k_hi) ? __ldcg(sp + (size_t)u * IDF * LP4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < kFinSlots; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
                }
                reinterpret_cast<float4*>(ds)[item] = a;              // ds[(s * IDF + i) * LP + 4 q]
                if (cg == 0) {
                    float* o = p.dSrc + ((size_t)b * IDF + i) * L + 4 * q;
                    if (4 * q < L) o[0] = a.x;
                    if (4 * q + 1 < L) o[1] = a.y;
                    if (4 * q + 2 < L) o[2] = a.z;
                    if (4 * q + 3 < L) o[3] = a.w;
                }
            }
            __syncthreads();
            if (tid == 0 && bb == b_lo) { SBA_TL(p.tl, 12); SBA_TL(p.tl, 13); }
            if (kq < nb) {
                const float* cc = cs + kq * L * kFinCS + cq;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = i0 + 32 * h;
                    if (i < IDF) {
                        const float* d0p = ds + (kq * IDF + i) * LP;
#pragma unroll 6
                        for (int l = 0; l < L; ++l) {
                            const float4 c4 = *reinterpret_cast<const float4*>(cc + l * kFinCS);
                            const float d0 = d0p[l];
                            acc[h][0] = fmaf(d0, c4.x, acc[h][0]); acc[h][1] = fmaf(d0, c4.y, acc[h][1]);
                            acc[h][2] = fmaf(d0, c4.z, acc[h][2]); acc[h][3] = fmaf(d0, c4.w, acc[h][3]);
                        }
                    }
                }
            }
        }
        if (tid == 0) { SBA_TL(p.tl, 14); SBA_TL(p.tl, 15); }
        // the 4 thread groups (samples of a round) meet in shared memory: red[kq][i][32 c], added in kq order
        __syncthreads();                                 // everybody is done reading ds
        float* red = ds;                                 // (kFinRound * IDF * 32 floats <= kFinRound * IDF * LP? no: sized below)
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int i = i0 + 32 * h;
            if (i < IDF) *reinterpret_cast<float4*>(red + ((size_t)kq * IDF + i) * 32 + cq) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
        }
        __syncthreads();
        float* mine = p.dwp + ((size_t)cg * p.groups + grp) * IDF * 32;
        for (int o = tid; o < IDF * 8; o += kFinThreads) {            // group partial -> dwp[cg][grp][i][32]
            float4 a = reinterpret_cast<const float4*>(red)[o];
#pragma unroll
            for (int k = 1; k < kFinRound; ++k) {
                const float4 v = reinterpret_cast<const float4*>(red)[k * IDF * 8 + o];
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            reinterpret_cast<float4*>(mine)[o] = a;
        }
        __syncthreads();
        if (tid == 0) {
            // release (cumulative over the block's stores, ordered by the barrier) / acquire on the slice counter
            uint32_t old;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + 1 + cg) : "memory");
            s_last = old == (uint32_t)(p.groups - 1);
        }
        __syncthreads();
        if (tid == 0) { SBA_TL(p.tl, 0); SBA_TL(p.tl, 1); }
        if (s_last) {
            // every group's partial of this channel slice is complete: thread = (quarter gq of the groups, output float4 o);
            // each thread adds its groups in order (batches of 8 independent loads), the quarters meet in shared memory
            const float4* base = reinterpret_cast<const float4*>(p.dwp + (size_t)cg * p.groups * IDF * 32);
            float4* red4 = reinterpret_cast<float4*>(red);
            const int gper = (p.groups + 3) >> 2;
            for (int o0 = 0; o0 < IDF * 8; o0 += 256) {
                const int o = o0 + (tid & 255), gq = tid >> 8;
                const int g_lo = gq * gper, g_hi = (g_lo + gper < p.groups ? g_lo + gper : p.groups);
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int g0 = g_lo; g0 < g_hi; g0 += 8) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        v[u] = (o < IDF * 8 && g0 + u < g_hi) ? __ldcg(base + (size_t)(g0 + u) * IDF * 8 + o)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
                }
                __syncthreads();
                red4[tid] = a;
                __syncthreads();
                if (tid < 256 && o < IDF * 8) {
                    float4 r = red4[tid];
#pragma unroll
                    for (int k = 1; k < 4; ++k) {
                        const float4 w = red4[k * 256 + tid];
                        r.x += w.x; r.y += w.y; r.z += w.z; r.w += w.w;
                    }
                    const int i = o >> 3, c4 = (o & 7) * 4;
                    float* o4 = p.dW + (size_t)i * cdf + c0 + c4;
                    if (c4 + 3 < nc && (cdf & 3) == 0) *reinterpret_cast<float4*>(o4) = r;
                    else {
                        if (c4 < nc) o4[0] = r.x;
                        if (c4 + 1 < nc) o4[1] = r.y;
                        if (c4 + 2 < nc) o4[2] = r.z;
                        if (c4 + 3 < nc) o4[3] = r.w;
                    }
                }
            }
        }
    } else {
        float* ds = sm;                  // [idf][L]
        const int b = blockIdx.x - p.n_dw;
        const int k_lo = tile_owner(b * TPS, p.n_tiles, p.n_ctas), k_hi = tile_owner((b + 1) * TPS - 1, p.n_tiles, p.n_ctas);
        asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll 1
        for (int o = tid; o < IDF * LP; o += blockDim.x) {
            const int i = o / LP, l = o - i * LP;
            float a = 0.f;
#pragma unroll 4
            for (int k = k_lo; k <= k_hi; ++k) a += __ldcg(p.part + ((size_t)(k + b) * IDF + i) * LP + l);
            if (l < L) {
                ds[i * L + l] = a;
                if (p.n_dw == 0) p.dSrc[((size_t)b * IDF + i) * L + l] = a;
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int o = tid; o < cdf * L; o += blockDim.x) {
            const int c = o / L, l = o - c * L;
            float a = 0.f;
#pragma unroll 4
            for (int i = 0; i < IDF; ++i) a = fmaf(__ldg(p.W + (size_t)i * cdf + c), ds[i * L + l], a);
            p.dCtx[(size_t)b * cdf * L + o] = a;
        }
    }
    if (tid == 0) SBA_TL(p.tl, 5);
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
    float* g_xchg = reinterpret_cast<float*>(g_out + 4 * C::OUT_WARP_BYTES);      // [idf][LP] sample flush exchange

    unsigned long long* bars = reinterpret_cast<unsigned long long*>(g_out + 4 * C::OUT_WARP_BYTES + C::XCHG_BYTES);
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
    // let the finish kernel's blocks become resident wherever there is room; they park in griddepcontrol.wait
    // until this grid has completed, which takes its launch latency off the critical path
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) SBA_TL(p.tl, 2);

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
        // This kernel heads the call's launch chain (programmatic dependent of whatever precedes it in the stream):
        // everything above ran under the predecessor's tail; nothing of the caller's memory is touched before this.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        SBA_TL(p.tl, 6);
        if (blockIdx.x == 0)
            for (int c = 0; c < p.n_counters; ++c) p.counters[c] = 0u;      // the finish kernel's completion counters
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
                    tma_load_2d(dst + (2 * bx) * C::BOX_BYTES, &tm_g, t * TQ + bx * C::BOX_PX, b * p.g_rows + p.g_row0, full);
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
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (thread 0 already passed it: returns at once)


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
                    tma_load_2d(dst + (2 * bx) * C::BOX_BYTES, &tm_g, t * TQ + bx * C::BOX_PX, b * p.g_rows + p.g_row0, full);
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
                        sv[g][k] = (ch < IDF && l < L) ? __ldcg(sb + ch * L + l) : 0.f;
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
            const int q = t * TQ + px;
            uint32_t mb = pad_bits;
            if (p.mask != nullptr) mb |= __ldcg(p.mask_bits + (p.mask_mode == SBA_MASK_PER_SAMPLE ? (uint32_t)b : cap));
            float ga[HAS_GA ? LP : 1];
            if constexpr (HAS_GA) {
                const T* gp = static_cast<const T*>(p.ga) + (size_t)b * L * Q + q;
#pragma unroll
                for (int l = 0; l < LP; ++l) ga[l] = (l < L) ? __bfloat162float(gp[(size_t)l * Q]) : 0.f;
            }
            // ---- S and dP rows of this pixel ---------------------------------------------------------
            mbar_wait(smem_u32(&bar_s_full[buf]), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
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

            if (++t == TPS) { t = 0; ++b; }
            cap += step_mod;
            if (cap >= Bu) cap -= Bu;
        }
    } else {
        // --------------------------------- second-stage warps: thread = pixel -------------------
        // dX row of the pixel -> staged [channel][32 px] per warp -> one TMA box store; when the CTA leaves a sample
        // the diagonal blocks of the TMEM accumulator are added and stored to the (CTA, sample) partial slot.
        const int cw = warp & 3;
        const uint32_t tl = tmem_base + ((uint32_t)(cw * 32) << 16);
        const uint32_t so = s_out + cw * C::OUT_WARP_BYTES;
        T* go = reinterpret_cast<T*>(g_out + cw * C::OUT_WARP_BYTES) + lane;
        int b = b0, t = t0;
        for (int j = 0; j < n_local; ++j) {
            mbar_wait(smem_u32(&bar_c_full), (uint32_t)j & 1u);
            tc_fence_after();
            uint32_t cr[IDF];
            tmem_ld<IDF>(tl + C::COL_DX, cr);
            tmem_wait_ld();
            tc_fence_before();
            warp_arrive(smem_u32(&bar_dx_free), lane);
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

            const bool last_of_sample = (t + 1 == TPS) || (j + 1 == n_local);
            if (last_of_sample) {
                // every MMA of sample b issued by this CTA has completed (c_full(j)): flush its share of dSrc[b].
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
                // dSrc[b][ch] share = (g.P)[ch] + (x.dS)[ch]: the x rows travel through shared memory to the lanes
                // that hold the g rows, which add (always in this order) and store the slot row
                mma::named_bar_sync(1, 128);                   // the previous flush's readers are done
                float* xs = g_xchg + ch * LP;
                if (valid && is_x) {
#pragma unroll
                    for (int q = 0; q < LP / 4; ++q)
                        reinterpret_cast<uint4*>(xs)[q] = make_uint4(a1[4 * q], a1[4 * q + 1], a1[4 * q + 2], a1[4 * q + 3]);
                }
                mma::named_bar_sync(1, 128);
                if (valid && !is_x) {
                    float4* out = reinterpret_cast<float4*>(p.part + ((size_t)(blockIdx.x + b) * IDF + ch) * LP);
#pragma unroll
                    for (int q = 0; q < LP / 4; ++q) {
                        const float4 xv = reinterpret_cast<const float4*>(xs)[q];
                        out[q] = make_float4(__uint_as_float(a0[4 * q]) + xv.x, __uint_as_float(a0[4 * q + 1]) + xv.y,
                                             __uint_as_float(a0[4 * q + 2]) + xv.z, __uint_as_float(a0[4 * q + 3]) + xv.w);
                    }
                }
            }
            if (++t == TPS) { t = 0; ++b; }
        }
        if (warp == kFirstEpilogueWarp && lane == 0) SBA_TL(p.tl, 7);
        if (lane == 0) bulk_wait<0>();
        tc_fence_before();
    }

    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (tid == 0) SBA_TL(p.tl, 3);
}

// workspace layout behind the B*idf*L + B + 1 words the other kernel families use (sizes in floats)
struct Tc5BwdWs {
    size_t part, dwp, counters, total;
    int slots, groups, lp, n_counters;
};
Tc5BwdWs tc5_bwd_ws(int B, int idf, int cdf, int L, int sms) {
    Tc5BwdWs w{};
    w.lp = (L + 3) / 4 * 4;
    w.slots = 2 * sms + B;                                   // at most two CTAs per SM (TMEM), slot = CTA + sample
    // All blocks (cslices x groups) must be resident at once, and a 1024-thread block with 64 registers per thread fills
    // an SM: B = 64 -> 16 groups of 4 samples = 128 blocks, one round each; larger batches take several rounds per group.
    // (Measured: two 32-register blocks per SM instead cost 3 us at B = 64 and gain nothing at B = 128 / 256.)
    const int cslices = (cdf + 31) / 32;
    int gmax = sms / cslices;
    if (gmax > kFinGroups) gmax = kFinGroups;
    if (gmax < 1) gmax = 1;
    w.groups = (B + kFinRound - 1) / kFinRound < gmax ? (B + kFinRound - 1) / kFinRound : gmax;
    w.n_counters = 1 + (cdf + 31) / 32;
    const size_t head = ((size_t)B * idf * L + B + 1 + 3) / 4 * 4;
    w.part = head;
    w.dwp = w.part + (size_t)w.slots * idf * w.lp;
    w.counters = w.dwp + (size_t)((cdf + 31) / 32) * w.groups * idf * 32;
    w.total = w.counters + (size_t)(w.n_counters + 3) / 4 * 4;
    return w;
}

template <int IDF, int NQ, bool HAS_GA, int NST_>
int launch_bwd_tc5(const void* x, const void* g, void* dX, const Tc5BwdParams& p_in, const Tc5FinishParams& f_in, int phase,
                   cudaStream_t st) {
    using C = Tc5BwdCfg<IDF, NQ, NST_>;
    auto kern = k_attn_bwd_tc5<IDF, NQ, HAS_GA, NST_>;
    constexpr size_t smem = (size_t)C::SMEM_BYTES + (2 * C::NST + 14) * 8;
    static_assert(smem <= 220 * 1024, "shared memory budget of the tcgen05 backward exceeded");
    int dev = 0, sms = 0;
    int rc = current_device(&dev, &sms, "attn_bwd(tcgen05)");
    if (rc) return rc;
    static std::atomic<unsigned long long> smem_done{0};
    rc = ensure_dynamic_smem(kern, smem, dev, smem_done, "attn_bwd(tcgen05)");
    if (rc) return rc;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > C::CTAS_PER_SM) per_sm = C::CTAS_PER_SM;
    if (per_sm > 2) per_sm = 2;                              // the slot workspace is sized for two CTAs per SM
    if (per_sm < 1) per_sm = 1;
#ifdef SBA_DEV_AIDS
    if (dev_tuning().ctas_per_sm > 0 && dev_tuning().ctas_per_sm < per_sm) per_sm = dev_tuning().ctas_per_sm;
#endif
    const int max_ctas = sms * per_sm;
    Tc5BwdParams p = p_in;
    Tc5FinishParams f = f_in;
    const int grid = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;      // (a function of the shape and the device only:
    p.tl = f.tl = SBA_TL_SLOT();                                       //  the finish phase recomputes the same value)

    if (phase != SBA_PHASE_SECOND) {
        CUtensorMap tm_x, tm_g, tm_dx;
        rc = make_tile_map(&tm_x, x, SBA_BF16, p.B * IDF, p.Q, IDF, 64, true);
        if (!rc) rc = make_tile_map(&tm_g, g, SBA_BF16, p.B * p.g_rows, p.Q, IDF, 64, true);
        if (!rc) rc = make_tile_map(&tm_dx, dX, SBA_BF16, p.B * IDF, p.Q, IDF, 32, false);
        if (rc) return rc;
        PdlLaunch ml(dim3(grid), dim3(kBwdThreads), smem, st);
        cudaError_t e = cudaLaunchKernelEx(&ml.cfg, kern, tm_x, tm_g, tm_dx, p);
        if (e != cudaSuccess) {
            set_error("attn_bwd(tcgen05): launch: %s", cudaGetErrorString(e));
            return SBA_ERR_CUDA;
        }
        add_launches(1);
        rc = check_launch("attn_bwd(tcgen05)");
        if (rc || phase == SBA_PHASE_FIRST) return rc;
    }
    // finish kernel: slots -> dSrc, dW, dCtx
    f.n_ctas = grid;
    f.n_dw = f.dW != nullptr ? f.groups * ((f.cdf + 31) / 32) : 0;
    const int fgrid = f.n_dw + (f.dCtx != nullptr ? f.B : 0);
    if (fgrid == 0) return SBA_OK;
    // cs [128][36] + ds / red [4][IDF][32]: 34 KB at idf 32
    constexpr size_t fsmem = (size_t)(kFinRound * 32 * kFinCS + kFinRound * IDF * 32) * sizeof(float);
    static std::atomic<unsigned long long> fsmem_done{0};
    rc = ensure_dynamic_smem(k_bwd_finish_tc5<IDF>, fsmem, dev, fsmem_done, "attn_bwd(finish)");
    if (rc) return rc;
    PdlLaunch fl(dim3(fgrid), dim3(kFinThreads), fsmem, st);
    cudaError_t e = cudaLaunchKernelEx(&fl.cfg, k_bwd_finish_tc5<IDF>, f);
    if (e != cudaSuccess) {
        set_error("attn_bwd(finish): launch: %s", cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    add_launches(1);
    return check_launch("attn_bwd(finish)");
}

template <int IDF, bool HAS_GA>
int dispatch_nq(const void* x, const void* g, void* dX, const Tc5BwdParams& p, const Tc5FinishParams& f, int phase,
                cudaStream_t st) {
    // long streams at idf 32 (>= 16 tiles per CTA of a 2-per-SM grid, e.g. 128x128 at B >= 40) take the 4-deep ring:
    // it still fits two CTAs per SM there and measured 2-4 % faster; short streams lose to its longer prologue
    constexpr bool kDeep = IDF == 32;
    const bool deep = kDeep && p.n_tiles >= 16 * 2 * 148;
#define SBA_BWD_CASE(n)                                                                     \
    case n:                                                                                 \
        if constexpr (kDeep) {                                                              \
            if (deep) return launch_bwd_tc5<IDF, n, HAS_GA, 4>(x, g, dX, p, f, phase, st);         \
        }                                                                                   \
        return launch_bwd_tc5<IDF, n, HAS_GA, 3>(x, g, dX, p, f, phase, st);
    switch ((p.L + 3) / 4) {
        SBA_BWD_CASE(1) SBA_BWD_CASE(2) SBA_BWD_CASE(3) SBA_BWD_CASE(4)
        SBA_BWD_CASE(5) SBA_BWD_CASE(6) SBA_BWD_CASE(7) SBA_BWD_CASE(8)
        default: return -1;
    }
#undef SBA_BWD_CASE
}

template <int IDF>
int dispatch_ga(const void* x, const void* g, void* dX, const Tc5BwdParams& p, const Tc5FinishParams& f, int phase,
                cudaStream_t st) {
    return p.ga != nullptr ? dispatch_nq<IDF, true>(x, g, dX, p, f, phase, st) : dispatch_nq<IDF, false>(x, g, dX, p, f, phase, st);
}

}  // namespace

bool tc5_bwd_supports(const AttnShape& s) { return tc5_supports(s) && s.dtype == SBA_BF16; }

size_t attn_bwd_workspace_floats(int B, int idf, int cdf, int L) {
    int dev = 0, sms = 0;
    if (current_device(&dev, &sms, "attn_bwd_workspace") != SBA_OK) sms = 160;     // no device (build check): an upper bound
    return tc5_bwd_ws(B, idf, cdf, L, sms).total;
}

int tc5_attn_bwd(const void* x, const float* ctx, const float* W, const float* srcT, const uint8_t* mask,
                 uint32_t* mask_bits, const void* g_c,
                 const void* g_attn, void* dX, float* ws, size_t ws_floats, float* dW, float* dCtx, const AttnShape& s,
                 cudaStream_t st) {
    int dev = 0, sms = 0;
    int rc = current_device(&dev, &sms, "attn_bwd(tcgen05)");
    if (rc) return rc;
    const Tc5BwdWs w = tc5_bwd_ws(s.B, s.idf, s.cdf, s.L, sms);
    if (ws_floats < w.total) {
        set_error("attn_bwd(tcgen05): workspace of %zu floats given, %zu needed (sba_attn_bwd_workspace_floats)", ws_floats,
                  w.total);
        return SBA_ERR_ARG;
    }
    Tc5BwdParams p{};
    p.srcT = srcT; p.mask = mask; p.mask_bits = mask_bits; p.ga = g_attn;
    p.part = ws + w.part;
    p.counters = reinterpret_cast<uint32_t*>(ws + w.counters);
    p.n_counters = w.n_counters;
    p.B = s.B; p.L = s.L; p.Q = s.Q; p.mask_mode = s.mask_mode;
    p.g_rows = s.c_rows > 0 ? s.c_rows : s.idf;
    p.g_row0 = s.c_rows > 0 ? s.c_row0 : 0;
    p.tiles_per_sample = s.Q / tc5::TQ;
    p.n_tiles = s.B * p.tiles_per_sample;
    Tc5FinishParams f{};
    f.part = p.part; f.ctx = ctx; f.W = W; f.dSrc = ws; f.dW = dW; f.dCtx = dCtx; f.dwp = ws + w.dwp; f.counters = p.counters;
    f.B = s.B; f.cdf = s.cdf; f.L = s.L; f.LP = w.lp; f.groups = w.groups;
    f.tiles_per_sample = p.tiles_per_sample; f.n_tiles = p.n_tiles;
    f.ctx_aligned = (reinterpret_cast<uintptr_t>(ctx) & 15u) == 0;
    rc = -1;
    if (s.idf == 32) rc = dispatch_ga<32>(x, g_c, dX, p, f, s.phase, st);
    else if (s.idf == 48) rc = dispatch_ga<48>(x, g_c, dX, p, f, s.phase, st);
    else if (s.idf == 64) rc = dispatch_ga<64>(x, g_c, dX, p, f, s.phase, st);
    if (rc == -1) {
        set_error("attn_bwd(tcgen05): unsupported shape idf=%d L=%d", s.idf, s.L);
        return SBA_ERR_UNSUPPORTED;
    }
    return rc;
}

}  // namespace sba

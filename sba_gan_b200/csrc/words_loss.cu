// Kernel (c): fused DAMSM region-word similarity (the B x B inner loop of words_loss).
//
// Reference: AttnGAN2/code/miscc/losses.py:62-123 calling func_attention
// (AttnGAN2/code/GlobalAttention.py:31-69) and cosine_similarity (losses.py:11-17) once per
// caption in a Python loop.  Here one CTA owns image j and a block of NC captions and runs
//   P1  s[r,t]   = sum_c X[c,r] W[c,t]                 (thread = RT regions x all words)
//       a1       = softmax_t(s);  a2[t,:] = softmax_r(gamma1 * a1[:,t])
//   P2  wc[c,t]  = sum_r X[c,r] a2[t,r]                (thread = 4 channels x all words)
//       cos_t    = <W_t, wc_t> / max(|W_t||wc_t|, eps); sim = gamma3 log sum_t exp(gamma2 cos_t)
// with the image tile streamed through shared memory twice (channel-major for P1,
// region-major for P2) by cp.async into double buffers, the next slice landing under the FMAs of
// the current one.  The backward (template BWD) recomputes the above, then
//   P3  da2[r,t] = sum_c X[c,r] dwc[c,t],  softmax backward twice -> ds
// and materialises per-pair  u = ds + alpha a2,  a2,  v = beta wc  in HBM so that the two
// dense gradient contractions become plain batched GEMMs (k_gemm below):
//   d_img[j]   = sum_i  W_i u_ji + v_ji a2_ji
//   d_words[i] = sum_j  X_j u_ji^T + kappa_i W_i
// Formulas: oracle/attention.py::words_loss_backward (checked against reference autograd).
#include "kernels.h"

namespace sba {
namespace {

template <int TP_, int RT_, int TPC_, int NC_>
struct WL {
    static constexpr int TP = TP_;    // words per caption held in registers (>= Lw, even)
    static constexpr int RT = RT_;    // regions per thread in P1/P3
    static constexpr int TPC = TPC_;  // threads per caption
    static constexpr int NC = NC_;    // captions per CTA
    static constexpr int kThreads = TPC * NC;
    static constexpr int kWarps = kThreads / 32;
    static constexpr int WPC = TPC / 32;
    static constexpr int RP = RT * TPC;  // padded region count
    static constexpr int KC = 32;        // channels per P1/P3 chunk
    static constexpr int RC = 32;        // regions per P2 chunk
    static constexpr int CPT = 4;        // channels per thread in P2
    static constexpr int CMAX = CPT * TPC;
    static constexpr int XT_STRIDE = CMAX + 1;
    static constexpr int XBUF = (KC * RP > RC * XT_STRIDE) ? KC * RP : RC * XT_STRIDE;
    static constexpr int WBUF = KC * NC * TP;
    static constexpr int A2T = NC * RP * TP;
    static constexpr int SCR = NC * WPC * 32;
    static constexpr int ALPHA = NC * 32;
    // x / w tiles double buffered where that fits in 220 KB (all but the 32-word variant)
    static constexpr int NBUF = (size_t)(2 * XBUF + 2 * WBUF + A2T + SCR + ALPHA) * sizeof(float) <= 220 * 1024 ? 2 : 1;
    static constexpr size_t kSmemBytes = (size_t)(NBUF * XBUF + NBUF * WBUF + A2T + SCR + ALPHA) * sizeof(float);
};

struct OpSum { __device__ __forceinline__ float operator()(float a, float b) const { return a + b; } };
struct OpMax { __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); } };

// lane l ends with op over lanes of v[l] in v[0]
template <class Op>
__device__ __forceinline__ void reduce_scatter32(float (&v)[32], int lane, Op op) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = upper ? v[k] : v[k + s];
            const float keep = upper ? v[k + s] : v[k];
            v[k] = op(keep, __shfl_xor_sync(0xffffffffu, send, s));
        }
    }
}

// All-reduce TP per-thread values over the TPC threads of one caption.  Every thread of the
// CTA must call it (two __syncthreads inside).
template <class C, class Op>
__device__ __forceinline__ void caption_allreduce(float (&v)[C::TP], float* scratch, int cap, int wic, int lane, Op op,
                                                  float identity) {
    float w[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) w[l] = (l < C::TP) ? v[l < C::TP ? l : 0] : identity;
    reduce_scatter32(w, lane, op);
    scratch[(cap * C::WPC + wic) * 32 + lane] = w[0];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < C::TP; ++t) {
        float r = scratch[(cap * C::WPC) * 32 + t];
#pragma unroll
        for (int k = 1; k < C::WPC; ++k) r = op(r, scratch[(cap * C::WPC + k) * 32 + t]);
        v[t] = r;
    }
    __syncthreads();
}


// 4-byte asynchronous global -> shared copy (zero fill when !valid): the next image tile streams in while the
// FMAs of the current one run
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct WordsArgs {
    const float* img;       // [B_img][nef][R]
    const float* words;     // [B_cap][nef][Lw]
    const int* cap_lens;    // [B_cap]
    float* sim;             // [B_img][B_cap]
    float* att_diag;        // [B_cap][Lw][R] nullable
    float* wc_out;          // paired mode only: [B][nef][Lw]
    const float* d_sim;     // BWD
    float* ws_u;            // BWD [B_img][B_cap][Lw][R]
    float* ws_a2;           // BWD [B_img][B_cap][Lw][R]
    float* ws_v;            // BWD [B_img][nef][B_cap][Lw]
    float* kappa;           // BWD [B_cap][Lw] (zeroed by the caller), nullable
    int B_img, B_cap, row_offset, nef, R, Lw;
    float g1, g2, g3, eps;
    int paired;             // func_attention mode: image j against caption j + pair_shift only; length fixed_T, or
    int fixed_T;            // cap_lens[caption] when cap_lens is given (the diagonal attention maps of words_loss)
    int pair_shift;
};

// [KC x RP] channel-major tile of image j -> shared memory, asynchronously (cp.async, zero fill)
template <class C>
__device__ __forceinline__ void load_x_p1_async(const WordsArgs& a, int j, int c0, float* xbuf, int tid) {
    const float* xj = a.img + (size_t)j * a.nef * a.R;
    for (int idx = tid; idx < C::KC * C::RP; idx += C::kThreads) {
        const int kk = idx / C::RP, r = idx - kk * C::RP;
        const int c = c0 + kk;
        const bool ok = c < a.nef && r < a.R;
        cp_async4(xbuf + idx, ok ? xj + (size_t)c * a.R + r : xj, ok);
    }
}
// [KC x NC x TP] tile of the (optionally alpha-scaled + v-shifted) words: loaded into registers, stored later
template <class C>
struct WTile {
    static constexpr int N = (C::KC * C::NC * C::TP + C::kThreads - 1) / C::kThreads;
    float v[N];
};
template <class C, bool DWC>
__device__ __forceinline__ void load_w_regs(const WordsArgs& a, int j, int cap0, int c0, const int* T_s, const float* alpha_s,
                                            WTile<C>& w, int tid) {
#pragma unroll
    for (int s = 0; s < WTile<C>::N; ++s) {
        const int idx = tid + s * C::kThreads;
        float val = 0.f;
        if (idx < C::KC * C::NC * C::TP) {
            const int kk = idx / (C::NC * C::TP);
            const int rem = idx - kk * (C::NC * C::TP);
            const int cp = rem / C::TP, t = rem - cp * C::TP;
            const int c = c0 + kk;
            const int ii = a.paired ? j + a.pair_shift : cap0 + cp;
            if (c < a.nef && t < T_s[cp]) {
                val = __ldg(a.words + ((size_t)ii * a.nef + c) * a.Lw + t);
                if (DWC) {
                    // dwc[c,t] = alpha_t W[c,t] + v[c,t]; v was written by this CTA before the barrier
                    val = alpha_s[cp * 32 + t] * val + a.ws_v[(((size_t)j * a.nef + c) * a.B_cap + ii) * a.Lw + t];
                }
            }
        }
        w.v[s] = val;
    }
}
template <class C>
__device__ __forceinline__ void store_w_regs(const WTile<C>& w, float* wbuf, int tid) {
#pragma unroll
    for (int s = 0; s < WTile<C>::N; ++s) {
        const int idx = tid + s * C::kThreads;
        if (idx < C::KC * C::NC * C::TP) wbuf[idx] = w.v[s];
    }
}
// P1 / P3 contraction over all channels: acc[q][t] = sum_c X[c, r0+q] * Wc[c, cap, t], tiles double buffered
template <class C, bool DWC, bool STAGE>
__device__ __forceinline__ void p1_all(const WordsArgs& a, int j, int cap0, const int* T_s, const float* alpha_s, float* xbuf,
                                       float* wbuf, int r0, int cap, int tid, float (&acc)[C::RT][C::TP]);

// acc[a][t] += sum over the chunk's channels of X[c, r0+a] * Wc[c, cap, t]
template <class C>
__device__ __forceinline__ void p1_chunk(const float* xbuf, const float* wbuf, int r0, int cap, float (&acc)[C::RT][C::TP]) {
#pragma unroll 2
    for (int kk = 0; kk < C::KC; ++kk) {
        float x[C::RT];
        if constexpr (C::RT == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(xbuf + kk * C::RP + r0);
            x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w;
        } else {
#pragma unroll
            for (int q = 0; q < C::RT; ++q) x[q] = xbuf[kk * C::RP + r0 + q];
        }
        const float2* w2 = reinterpret_cast<const float2*>(wbuf + (kk * C::NC + cap) * C::TP);
#pragma unroll
        for (int t2 = 0; t2 < C::TP / 2; ++t2) {
            const float2 w = w2[t2];
#pragma unroll
            for (int q = 0; q < C::RT; ++q) {
                acc[q][2 * t2] = fmaf(x[q], w.x, acc[q][2 * t2]);
                acc[q][2 * t2 + 1] = fmaf(x[q], w.y, acc[q][2 * t2 + 1]);
            }
        }
    }
}

template <class C, bool DWC, bool STAGE>
__device__ __forceinline__ void p1_all(const WordsArgs& a, int j, int cap0, const int* T_s, const float* alpha_s, float* xbuf,
                                       float* wbuf, int r0, int cap, int tid, float (&acc)[C::RT][C::TP]) {
    WTile<C> w;
    __syncthreads();                      // the tile buffers are free (and, DWC: ws_v / alpha_s writes are visible)
    load_x_p1_async<C>(a, j, 0, xbuf, tid);
    cp_async_commit();
    load_w_regs<C, DWC>(a, j, cap0, 0, T_s, alpha_s, w, tid);
    store_w_regs<C>(w, wbuf, tid);
    cp_async_wait_all();
    __syncthreads();
    int buf = 0;
    for (int c0 = 0; c0 < a.nef; c0 += C::KC) {
        const bool more = c0 + C::KC < a.nef;
        if constexpr (C::NBUF == 2) {
            if (more) {
                load_x_p1_async<C>(a, j, c0 + C::KC, xbuf + (buf ^ 1) * C::XBUF, tid);
                cp_async_commit();
                load_w_regs<C, DWC>(a, j, cap0, c0 + C::KC, T_s, alpha_s, w, tid);
                // the backward kernel runs at the register limit: its small word tile is stored at once instead of being
                // held across the FMA block (one exposed L2 round trip per chunk; the image tile stays asynchronous)
                if (!STAGE) store_w_regs<C>(w, wbuf + (buf ^ 1) * C::WBUF, tid);
            }
            p1_chunk<C>(xbuf + buf * C::XBUF, wbuf + buf * C::WBUF, r0, cap, acc);
            if (more) {
                if (STAGE) store_w_regs<C>(w, wbuf + (buf ^ 1) * C::WBUF, tid);
                cp_async_wait_all();
            }
            __syncthreads();
            buf ^= 1;
        } else {                                  // single buffers: load the next tile after the FMAs of this one
            p1_chunk<C>(xbuf, wbuf, r0, cap, acc);
            __syncthreads();
            if (more) {
                load_x_p1_async<C>(a, j, c0 + C::KC, xbuf, tid);
                cp_async_commit();
                load_w_regs<C, DWC>(a, j, cap0, c0 + C::KC, T_s, alpha_s, w, tid);
                store_w_regs<C>(w, wbuf, tid);
                cp_async_wait_all();
                __syncthreads();
            }
        }
    }
}

template <class C, bool BWD>
__global__ void __launch_bounds__(C::kThreads, 1) k_words(const WordsArgs a) {
    constexpr int TP = C::TP, RT = C::RT, TPC = C::TPC, NC = C::NC, RP = C::RP, CPT = C::CPT;
    extern __shared__ __align__(16) float smem[];
    float* xbuf = smem;                       // [NBUF][XBUF]
    float* wbuf = xbuf + C::NBUF * C::XBUF;   // [NBUF][WBUF]
    float* a2t = wbuf + C::NBUF * C::WBUF;    // [NC][RP][TP]
    float* scratch = a2t + C::A2T;   // [NC][WPC][32]
    float* alpha_s = scratch + C::SCR;  // [NC][32]
    __shared__ int T_s[NC];

    const int tid = threadIdx.x, lane = tid & 31;
    const int cap = tid / TPC, tc = tid - cap * TPC, wic = tc >> 5;
    const int j = blockIdx.y;
    const int cap0 = blockIdx.x * NC;
    const int i = a.paired ? j + a.pair_shift : cap0 + cap;
    const bool valid = a.paired ? (cap == 0 && i < a.B_cap) : (i < a.B_cap);
    if (tid < NC) {
        const int ii = a.paired ? j + a.pair_shift : cap0 + tid;
        const bool v = a.paired ? (tid == 0 && ii < a.B_cap) : (ii < a.B_cap);
        int T = v ? ((a.paired && a.cap_lens == nullptr) ? a.fixed_T : a.cap_lens[ii]) : 0;
        T = max(0, min(T, min(a.Lw, TP)));
        T_s[tid] = T;
    }
    __syncthreads();
    const int T = T_s[cap];
    const int r0 = tc * RT;

    // ---- P1: region x word scores --------------------------------------------------------
    float s[RT][TP];
#pragma unroll
    for (int q = 0; q < RT; ++q)
#pragma unroll
        for (int t = 0; t < TP; ++t) s[q][t] = 0.f;
    p1_all<C, false, !BWD>(a, j, cap0, T_s, alpha_s, xbuf, wbuf, r0, cap, tid, s);

    // ---- a1 = softmax over words (GlobalAttention.py:50-51), thread local ----------------
#pragma unroll
    for (int q = 0; q < RT; ++q) {
        float m = -INFINITY;
#pragma unroll
        for (int t = 0; t < TP; ++t)
            if (t < T) m = fmaxf(m, s[q][t]);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const float e = (t < T) ? __expf(s[q][t] - m) : 0.f;
            s[q][t] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int t = 0; t < TP; ++t) s[q][t] = (t < T) ? s[q][t] * inv : 0.f;
    }
    const size_t pair_base = valid ? (((size_t)j * a.B_cap + i) * a.Lw) * a.R : 0;
    if (BWD && valid) {
        // stash a1 in this pair's u slot; the same thread reads it back after P3
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            if (t < T) {
#pragma unroll
                for (int q = 0; q < RT; ++q)
                    if (r0 + q < a.R) a.ws_u[pair_base + (size_t)t * a.R + r0 + q] = s[q][t];
            }
        }
    }

    // ---- a2[t,:] = softmax over regions of gamma1 * a1[:,t] (GlobalAttention.py:56-60) ----
    {
        float red[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            float m = -INFINITY;
#pragma unroll
            for (int q = 0; q < RT; ++q)
                if (r0 + q < a.R) m = fmaxf(m, a.g1 * s[q][t]);
            red[t] = m;
        }
        caption_allreduce<C>(red, scratch, cap, wic, lane, OpMax(), -INFINITY);
        float sum[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < RT; ++q) {
                const float e = (t < T && r0 + q < a.R) ? __expf(a.g1 * s[q][t] - red[t]) : 0.f;
                s[q][t] = e;
                acc += e;
            }
            sum[t] = acc;
        }
        caption_allreduce<C>(sum, scratch, cap, wic, lane, OpSum(), 0.f);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const float inv = (t < T) ? 1.0f / sum[t] : 0.f;
#pragma unroll
            for (int q = 0; q < RT; ++q) s[q][t] = (t < T) ? s[q][t] * inv : 0.f;
        }
    }
    // a2 -> shared [cap][r][t] (P2 operand), and the diagonal pair's attention map
#pragma unroll
    for (int q = 0; q < RT; ++q) {
        float2* dst = reinterpret_cast<float2*>(a2t + ((size_t)cap * RP + r0 + q) * TP);
#pragma unroll
        for (int t2 = 0; t2 < TP / 2; ++t2) dst[t2] = make_float2(s[q][2 * t2], s[q][2 * t2 + 1]);
    }
    if (!BWD && valid && a.att_diag != nullptr && (a.paired || a.row_offset + j == i)) {
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            if (t < T) {
#pragma unroll
                for (int q = 0; q < RT; ++q)
                    if (r0 + q < a.R) a.att_diag[((size_t)i * a.Lw + t) * a.R + r0 + q] = s[q][t];
            }
        }
    }

    if (a.paired && a.sim == nullptr && a.wc_out == nullptr) return;       // only the attention maps were asked for (block-uniform)

    // ---- P2: wc[c,t] = sum_r X[c,r] a2[t,r] (GlobalAttention.py:67) ------------------------
    float wc[CPT][TP];
#pragma unroll
    for (int k = 0; k < CPT; ++k)
#pragma unroll
        for (int t = 0; t < TP; ++t) wc[k][t] = 0.f;
    {
        const float* xj = a.img + (size_t)j * a.nef * a.R;
        const int warp = tid >> 5;
        auto load_x_p2_async = [&](int rc0, float* xb) {
            const int r = rc0 + lane;
            for (int c = warp; c < a.nef; c += C::kWarps)
                cp_async4(xb + lane * C::XT_STRIDE + c, r < a.R ? xj + (size_t)c * a.R + r : xj, r < a.R);
        };
        __syncthreads();                  // P1's last tile has been consumed
        load_x_p2_async(0, xbuf);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        int buf = 0;
        for (int rc0 = 0; rc0 < a.R; rc0 += C::RC) {
            const bool more = rc0 + C::RC < a.R;
            if (C::NBUF == 2 && more) {
                load_x_p2_async(rc0 + C::RC, xbuf + (buf ^ 1) * C::XBUF);
                cp_async_commit();
            }
            const float* xb = xbuf + buf * C::XBUF;
#pragma unroll 2
            for (int rr = 0; rr < C::RC; ++rr) {
                float xv[CPT];
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const int ch = tc + k * TPC;
                    xv[k] = (ch < a.nef) ? xb[rr * C::XT_STRIDE + ch] : 0.f;
                }
                const float2* a2 = reinterpret_cast<const float2*>(a2t + ((size_t)cap * RP + rc0 + rr) * TP);
#pragma unroll
                for (int t2 = 0; t2 < TP / 2; ++t2) {
                    const float2 w = a2[t2];
#pragma unroll
                    for (int k = 0; k < CPT; ++k) {
                        wc[k][2 * t2] = fmaf(xv[k], w.x, wc[k][2 * t2]);
                        wc[k][2 * t2 + 1] = fmaf(xv[k], w.y, wc[k][2 * t2 + 1]);
                    }
                }
            }
            if constexpr (C::NBUF == 2) {
                if (more) cp_async_wait_all();
                __syncthreads();
                buf ^= 1;
            } else {
                __syncthreads();
                if (more) {
                    load_x_p2_async(rc0 + C::RC, xbuf);
                    cp_async_commit();
                    cp_async_wait_all();
                    __syncthreads();
                }
            }
        }
    }
    if (a.paired && valid && a.wc_out != nullptr) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int ch = tc + k * TPC;
            if (ch < a.nef) {
#pragma unroll
                for (int t = 0; t < TP; ++t)
                    if (t < T) a.wc_out[((size_t)j * a.nef + ch) * a.Lw + t] = wc[k][t];
            }
        }
    }

    // ---- cosine similarity per word (losses.py:11-17) and LSE (losses.py:106-108) --------
    float num[TP], wn2[TP], ww2[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) { num[t] = 0.f; wn2[t] = 0.f; ww2[t] = 0.f; }
    if (valid) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int ch = tc + k * TPC;
            if (ch < a.nef) {
                const float* wrow = a.words + ((size_t)i * a.nef + ch) * a.Lw;
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    if (t < T) {
                        const float w = __ldg(wrow + t);
                        num[t] = fmaf(w, wc[k][t], num[t]);
                        wn2[t] = fmaf(wc[k][t], wc[k][t], wn2[t]);
                        ww2[t] = fmaf(w, w, ww2[t]);
                    }
                }
            }
        }
    }
    caption_allreduce<C>(num, scratch, cap, wic, lane, OpSum(), 0.f);
    caption_allreduce<C>(wn2, scratch, cap, wic, lane, OpSum(), 0.f);
    caption_allreduce<C>(ww2, scratch, cap, wic, lane, OpSum(), 0.f);
    float E = 0.f;
    float ex[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) {
        const float den = fmaxf(sqrtf(ww2[t]) * sqrtf(wn2[t]), a.eps);
        const float cosv = num[t] / den;
        ex[t] = (t < T) ? expf(a.g2 * cosv) : 0.f;
        E += ex[t];
    }
    if (!BWD) {
        if (valid && tc == 0 && a.sim != nullptr) a.sim[(size_t)j * a.B_cap + i] = a.g3 * logf(E);
        return;
    }

    // ======================= backward part ================================================
    if constexpr (BWD) {
        // per-word scalars: alpha = d num, beta = d|wc| / |wc|, kappa += d|W| / |W|
        float beta[TP];
        {
            const float g = valid ? a.d_sim[(size_t)j * a.B_cap + i] : 0.f;
#pragma unroll
            for (int t = 0; t < TP; ++t) {
                const float ww = sqrtf(ww2[t]), wn = sqrtf(wn2[t]);
                const float prod = ww * wn;
                const float den = fmaxf(prod, a.eps);
                const float gcos = (t < T) ? g * a.g3 / E * a.g2 * ex[t] : 0.f;
                const float d_num = gcos / den;
                const float d_den = (prod > a.eps) ? -gcos * num[t] / (den * den) : 0.f;
                const float al = (t < T) ? d_num : 0.f;
                beta[t] = (t < T && wn > 0.f) ? d_den * ww / wn : 0.f;
                if (tc == 0) {
                    alpha_s[cap * 32 + t] = al;
                    if (valid && t < T && a.kappa != nullptr && ww > 0.f) atomicAdd(a.kappa + (size_t)i * a.Lw + t, d_den * wn / ww);
                }
                num[t] = al;  // keep alpha in registers
            }
        }
        // v = beta * wc  -> HBM [j][c][i][t]  (zero beyond the caption's length)
        if (valid) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int ch = tc + k * TPC;
                if (ch < a.nef) {
                    float* vrow = a.ws_v + (((size_t)j * a.nef + ch) * a.B_cap + i) * a.Lw;
#pragma unroll
                    for (int t = 0; t < TP; ++t)
                        if (t < a.Lw) vrow[t] = (t < T) ? beta[t] * wc[k][t] : 0.f;
                }
            }
        }
        // ---- P3: da2[r,t] = sum_c X[c,r] (alpha_t W[c,t] + v[c,t]) --------------------------
#pragma unroll
        for (int q = 0; q < RT; ++q)
#pragma unroll
            for (int t = 0; t < TP; ++t) s[q][t] = 0.f;
        p1_all<C, true, false>(a, j, cap0, T_s, alpha_s, xbuf, wbuf, r0, cap, tid, s);   // (its first barrier orders the ws_v / alpha_s writes)
        // softmax-over-regions backward: dz = a2 * (da2 - sum_r a2 da2)
        float dotr[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < RT; ++q) acc = fmaf(a2t[((size_t)cap * RP + r0 + q) * TP + t], s[q][t], acc);
            dotr[t] = acc;
        }
        caption_allreduce<C>(dotr, scratch, cap, wic, lane, OpSum(), 0.f);
        if (valid) {
#pragma unroll
            for (int q = 0; q < RT; ++q) {
                if (r0 + q >= a.R) continue;
                float a1[TP], a2v[TP], da1[TP];
                float dott = 0.f;
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    a2v[t] = a2t[((size_t)cap * RP + r0 + q) * TP + t];
                    a1[t] = (t < T) ? a.ws_u[pair_base + (size_t)t * a.R + r0 + q] : 0.f;
                    da1[t] = a.g1 * a2v[t] * (s[q][t] - dotr[t]);
                    dott = fmaf(a1[t], da1[t], dott);
                }
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    if (t < a.Lw) {
                        const float ds = a1[t] * (da1[t] - dott);
                        const size_t o = pair_base + (size_t)t * a.R + r0 + q;
                        a.ws_u[o] = (t < T) ? ds + num[t] * a2v[t] : 0.f;
                        a.ws_a2[o] = (t < T) ? a2v[t] : 0.f;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// batched fp32 GEMM with two-level K (and N) addressing for the gradient contractions
//   A(m,k) = A + b*sA_batch + (k / kbA)*sA_kb + m*lda + (k % kbA)
//   B(k,n) = transB ? B + b*sB_batch + (k / kbB)*sB_kb + n*ldb + (k % kbB)
//                   : B + b*sB_batch + k*ldb + n
//   C(m,n) = C + b*sC_batch + (n / nbC)*sC_nb + m*ldc + (n % nbC)      (+= when accumulate)
// grid = (N tiles, M tiles, batch * splits); with splits > 1 every z-slice owns a K range and adds its
// partial product with red.global.add (C must then hold the initial value; accumulate is implied).
// 128 x BN tile, 256 threads, 8 x (BN/16) outputs per thread as 4-wide groups (LDS.128 operands),
// shared tiles double-buffered, the next K tile's global loads in flight during the FMA block.
// ---------------------------------------------------------------------------------------
struct GemmArgs {
    const float* A; long long sA_batch, sA_kb; int kbA, lda;
    const float* B; long long sB_batch, sB_kb; int kbB, ldb, transB;
    float* C; long long sC_batch, sC_nb; int nbC, ldc, accumulate;
    int M, N, K, splits;
};

template <int BN>
__global__ void __launch_bounds__(256, 2) k_gemm(const GemmArgs g) {
    constexpr int BM = 128, BK = 16, TM = 8, TN = BN / 16, NV = TN / 4;
    constexpr int LA = BM * BK / 256, LB = BN * BK / 256;
    static_assert(TN % 4 == 0, "BN must be a multiple of 64");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int b = blockIdx.z / g.splits, z = blockIdx.z - b * g.splits;
    const int K = g.K;
    const int kper = ceil_div(ceil_div(K, g.splits), BK) * BK;
    const int kbeg = z * kper, kend = min(K, kbeg + kper);
    const float* A = g.A + (size_t)b * g.sA_batch;
    const float* B = g.B + (size_t)b * g.sB_batch;

    // per-thread load slots: A element (am, ak) and B element (bn, bk) of the current K tile;
    // the two-level K position (block, remainder) is advanced incrementally, no division per tile
    const int ak = tid & (BK - 1), am = tid >> 4;                  // + 16 rows per slot
    int a_kb = (kbeg + ak) / g.kbA, a_kr = (kbeg + ak) - a_kb * g.kbA;
    const int bk_t = tid & (BK - 1), bn_t = tid >> 4;              // transB: k fastest
    const int bn_n = tid % BN, bk_n = tid / BN;                    // non-trans: n fastest, + 256/BN k per slot
    int b_kb = 0, b_kr = 0;
    if (g.transB) { b_kb = (kbeg + bk_t) / g.kbB; b_kr = (kbeg + bk_t) - b_kb * g.kbB; }

    float ra[LA], rb[LB];
    auto gload = [&](int k0) {
        const bool kok = k0 + ak < kend;
        const float* ap = A + (size_t)a_kb * g.sA_kb + a_kr;
#pragma unroll
        for (int s = 0; s < LA; ++s) {
            const int m = m0 + am + s * 16;
            ra[s] = (kok && m < g.M) ? __ldg(ap + (size_t)m * g.lda) : 0.f;
        }
        a_kr += BK;
        while (a_kr >= g.kbA) { a_kr -= g.kbA; ++a_kb; }
        if (g.transB) {
            const bool bok = k0 + bk_t < kend;
            const float* bp = B + (size_t)b_kb * g.sB_kb + b_kr;
#pragma unroll
            for (int s = 0; s < LB; ++s) {
                const int n = n0 + bn_t + s * 16;
                rb[s] = (bok && n < g.N) ? __ldg(bp + (size_t)n * g.ldb) : 0.f;
            }
            b_kr += BK;
            while (b_kr >= g.kbB) { b_kr -= g.kbB; ++b_kb; }
        } else {
            const int n = n0 + bn_n;
#pragma unroll
            for (int s = 0; s < LB; ++s) {
                const int k = k0 + bk_n + s * (256 / BN);
                rb[s] = (k < kend && n < g.N) ? __ldg(B + (size_t)k * g.ldb + n) : 0.f;
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int s = 0; s < LA; ++s) As[buf][ak][am + s * 16] = ra[s];
        if (g.transB) {
#pragma unroll
            for (int s = 0; s < LB; ++s) Bs[buf][bk_t][bn_t + s * 16] = rb[s];
        } else {
#pragma unroll
            for (int s = 0; s < LB; ++s) Bs[buf][bk_n + s * (256 / BN)][bn_n] = rb[s];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (kbeg < kend) {
        gload(kbeg);
        sstore(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            const bool more = k0 + BK < kend;
            if (more) gload(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float av[TM], bv[TN];
#pragma unroll
                for (int v = 0; v < 2; ++v)
                    *reinterpret_cast<float4*>(&av[4 * v]) = *reinterpret_cast<const float4*>(&As[buf][kk][v * 64 + ty * 4]);
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    *reinterpret_cast<float4*>(&bv[4 * v]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][v * 64 + tx * 4]);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (more) {
                sstore(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }
    float* Cb = g.C + (size_t)b * g.sC_batch;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
            if (n >= g.N) continue;
            const int nb = n / g.nbC;
            float* p = Cb + (size_t)nb * g.sC_nb + (size_t)m * g.ldc + (n - nb * g.nbC);
            if (g.splits > 1) atomicAdd(p, acc[i][j]);
            else *p = g.accumulate ? *p + acc[i][j] : acc[i][j];
        }
    }
}

// d_words[i][c][t] = kappa[i][t] * words[i][c][t]   (the direct |W| term; GEMM accumulates on top)
__global__ void k_kappa_init(const float* __restrict__ words, const float* __restrict__ kappa, float* __restrict__ d_words,
                             int B_cap, int nef, int Lw) {
    const size_t n = (size_t)B_cap * nef * Lw;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(o % Lw);
        const int i = (int)(o / ((size_t)nef * Lw));
        d_words[o] = kappa[(size_t)i * Lw + t] * words[o];
    }
}

template <class C, bool BWD>
int launch_words(const WordsArgs& a, cudaStream_t st) {
    if (a.R > C::RP || a.nef > C::CMAX || a.Lw > C::TP) {
        set_error("words_sim: shape R=%d nef=%d Lw=%d exceeds kernel limits (R<=%d nef<=%d Lw<=%d)", a.R, a.nef, a.Lw, C::RP,
                  C::CMAX, C::TP);
        return SBA_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaFuncSetAttribute(k_words<C, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    if (e != cudaSuccess) {
        set_error("words_sim: cudaFuncSetAttribute(%zu bytes): %s", C::kSmemBytes, cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    dim3 grid(a.paired ? 1 : ceil_div(a.B_cap, C::NC), a.B_img);
    k_words<C, BWD><<<grid, C::kThreads, C::kSmemBytes, st>>>(a);
    add_launches(1);
    return check_launch(BWD ? "words_sim_bwd" : "words_sim_fwd");
}

template <bool BWD>
int dispatch_words(const WordsArgs& a, cudaStream_t st) {
    if (a.Lw <= 18) return launch_words<WL<18, 5, 64, 4>, BWD>(a, st);
    if (a.Lw <= 24) return launch_words<WL<24, 4, 96, 2>, BWD>(a, st);
    return launch_words<WL<32, 3, 128, 2>, BWD>(a, st);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t words_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw) {
    const size_t pair = align256((size_t)B_img * B_cap * Lw * R * sizeof(float));
    const size_t v = align256((size_t)B_img * nef * B_cap * Lw * sizeof(float));
    const size_t kap = align256((size_t)B_cap * Lw * sizeof(float));
    return 2 * pair + v + kap;
}

// att_diag rows [row_offset, row_offset + B_img): the region attention of pair (local image j, caption row_offset + j) at the
// caption's true length - all that the tensor-core forward (words_tc5.cu) leaves to this kernel
int words_att_diag(const float* img, const float* words, const int* cap_lens, float* att_diag, int B_img, int B_cap,
                   int row_offset, int nef, int R, int Lw, float g1, cudaStream_t st) {
    WordsArgs a{};
    a.img = img; a.words = words; a.cap_lens = cap_lens; a.att_diag = att_diag;
    a.B_img = B_img; a.B_cap = B_cap; a.row_offset = row_offset; a.nef = nef; a.R = R; a.Lw = Lw;
    a.g1 = g1; a.g2 = 1.f; a.g3 = 1.f; a.eps = 1e-8f; a.paired = 1; a.fixed_T = Lw; a.pair_shift = row_offset;
    return dispatch_words<false>(a, st);
}

int words_sim_fwd(const float* img, const float* words, const int* cap_lens, float* sim, float* att_diag, float* wc_out,
                  int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float g1, float g2, float g3, float eps,
                  int paired, cudaStream_t st) {
    WordsArgs a{};
    a.img = img; a.words = words; a.cap_lens = cap_lens; a.sim = sim; a.att_diag = att_diag; a.wc_out = wc_out;
    a.B_img = B_img; a.B_cap = B_cap; a.row_offset = row_offset; a.nef = nef; a.R = R; a.Lw = Lw;
    a.g1 = g1; a.g2 = g2; a.g3 = g3; a.eps = eps; a.paired = paired; a.fixed_T = Lw;
    return dispatch_words<false>(a, st);
}

int words_sim_bwd(const float* img, const float* words, const int* cap_lens, const float* d_sim, float* d_img,
                  float* d_words, void* workspace, int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float g1,
                  float g2, float g3, float eps, cudaStream_t st) {
    const size_t pair = align256((size_t)B_img * B_cap * Lw * R * sizeof(float));
    const size_t vb = align256((size_t)B_img * nef * B_cap * Lw * sizeof(float));
    char* ws = static_cast<char*>(workspace);
    WordsArgs a{};
    a.img = img; a.words = words; a.cap_lens = cap_lens; a.d_sim = d_sim;
    a.ws_u = reinterpret_cast<float*>(ws);
    a.ws_a2 = reinterpret_cast<float*>(ws + pair);
    a.ws_v = reinterpret_cast<float*>(ws + 2 * pair);
    a.kappa = d_words ? reinterpret_cast<float*>(ws + 2 * pair + vb) : nullptr;
    a.B_img = B_img; a.B_cap = B_cap; a.row_offset = row_offset; a.nef = nef; a.R = R; a.Lw = Lw;
    a.g1 = g1; a.g2 = g2; a.g3 = g3; a.eps = eps; a.paired = 0;
    if (a.kappa) {
        cudaError_t e = cudaMemsetAsync(a.kappa, 0, (size_t)B_cap * Lw * sizeof(float), st);
        if (e != cudaSuccess) { set_error("words_sim_bwd: memset: %s", cudaGetErrorString(e)); return SBA_ERR_CUDA; }
    }
    int rc = dispatch_words<true>(a, st);
    if (rc) return rc;
    const int KL = B_cap * Lw;
    const int big = 0x7fffffff;
    // d_img[j] = W2 [nef x KL] . u_j [KL x R]  +  v_j [nef x KL] . a2_j [KL x R]
    GemmArgs g{};
    g.A = words; g.sA_batch = 0; g.sA_kb = (long long)nef * Lw; g.kbA = Lw; g.lda = Lw;
    g.B = a.ws_u; g.sB_batch = (long long)KL * R; g.sB_kb = 0; g.kbB = big; g.ldb = R; g.transB = 0;
    g.C = d_img; g.sC_batch = (long long)nef * R; g.sC_nb = 0; g.nbC = big; g.ldc = R; g.accumulate = 0;
    g.M = nef; g.N = R; g.K = KL; g.splits = 1;
    dim3 grid2(ceil_div(R, 64), ceil_div(nef, 128), B_img);
    k_gemm<64><<<grid2, 256, 0, st>>>(g);
    g.A = a.ws_v; g.sA_batch = (long long)nef * KL; g.sA_kb = 0; g.kbA = big; g.lda = KL;
    g.B = a.ws_a2; g.accumulate = 1;
    k_gemm<64><<<grid2, 256, 0, st>>>(g);
    add_launches(2);
    rc = check_launch("words_sim_bwd(d_img gemm)");
    if (rc) return rc;
    if (d_words != nullptr) {
        // d_words[i][c][t] = kappa_i[t] W_i[c][t] + sum_{j,r} X_j[c][r] u_j[(i,t)][r]:
        // one GEMM with M = nef, N = (i,t), K = (j,r), split over K so that the grid fills the GPU
        k_kappa_init<<<ceil_div(B_cap * nef * Lw, 256) < 1024 ? ceil_div(B_cap * nef * Lw, 256) : 1024, 256, 0, st>>>(
            words, a.kappa, d_words, B_cap, nef, Lw);
        GemmArgs h{};
        h.A = img; h.sA_batch = 0; h.sA_kb = (long long)nef * R; h.kbA = R; h.lda = R;
        h.B = a.ws_u; h.sB_batch = 0; h.sB_kb = (long long)KL * R; h.kbB = R; h.ldb = R; h.transB = 1;
        h.C = d_words; h.sC_batch = 0; h.sC_nb = (long long)nef * Lw; h.nbC = Lw; h.ldc = Lw; h.accumulate = 1;
        h.M = nef; h.N = KL; h.K = B_img * R;
        const int tiles = ceil_div(KL, 128) * ceil_div(nef, 128);
        int splits = ceil_div(2 * 148 * 2, tiles);                 // about two waves of 2 CTAs per SM
        if (splits > B_img) splits = B_img;
        if (splits < 1) splits = 1;
        h.splits = splits;
        dim3 grid3(ceil_div(KL, 128), ceil_div(nef, 128), splits);
        k_gemm<128><<<grid3, 256, 0, st>>>(h);
        add_launches(2);
        rc = check_launch("words_sim_bwd(d_words gemm)");
        if (rc) return rc;
    }
    return SBA_OK;
}

}  // namespace sba

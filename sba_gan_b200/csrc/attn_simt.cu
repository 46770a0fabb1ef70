// CUDA-core (FFMA + warp-shuffle) kernels for GlobalAttentionGeneral forward/backward.
//
// This is the exact-fp32 family (SBA_ALGO_SIMT): one thread owns PX consecutive pixels and
// keeps all L word scores of those pixels in registers, so the masked softmax over words
// needs no cross-lane traffic; the only cross-lane step is the dSrc reduction of the
// backward (a 32-lane butterfly reduce-scatter).  It covers every shape with L <= 32 and
// is the fallback of the tensor-core family (attn_mma.cu) for shapes that one rejects.
//
// Reference semantics: AttnGAN2/code/GlobalAttention.py:82-121 and its autograd backward
// (formulas in SURVEY.md §8a-3/a-4, restated in oracle/attention.py).
#include "kernels.h"

namespace sba {
namespace {

constexpr int kThreads = 128;

// ---------------------------------------------------------------------------------------
// srcT[b] = W . ctx[b]  (the bias-free 1x1 conv_context, GlobalAttention.py:95-97), plus
// the caption padding mask packed to one bit word per caption.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_project(const float* __restrict__ ctx, const float* __restrict__ W, const uint8_t* __restrict__ mask,
          float* __restrict__ srcT, uint32_t* __restrict__ mask_bits, int idf, int cdf, int L) {
    extern __shared__ __align__(16) float ctx_s[];  // [cdf][L]
    const int b = blockIdx.x;
    const float* cb = ctx + (size_t)b * cdf * L;
    for (int o = threadIdx.x; o < cdf * L; o += blockDim.x) ctx_s[o] = __ldg(cb + o);
    if (mask != nullptr && threadIdx.x == 0) {
        uint32_t bits = 0;
        for (int l = 0; l < L; ++l) bits |= (mask[(size_t)b * L + l] ? 1u : 0u) << l;
        mask_bits[b] = bits;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < idf * L; o += blockDim.x) {
        const int i = o / L, l = o - i * L;
        const float* w = W + (size_t)i * cdf;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int c = 0;
        for (; c + 3 < cdf; c += 4) {
            a0 = fmaf(__ldg(w + c), ctx_s[c * L + l], a0);
            a1 = fmaf(__ldg(w + c + 1), ctx_s[(c + 1) * L + l], a1);
            a2 = fmaf(__ldg(w + c + 2), ctx_s[(c + 2) * L + l], a2);
            a3 = fmaf(__ldg(w + c + 3), ctx_s[(c + 3) * L + l], a3);
        }
        for (; c < cdf; ++c) a0 = fmaf(__ldg(w + c), ctx_s[c * L + l], a0);
        srcT[(size_t)b * idf * L + o] = (a0 + a1) + (a2 + a3);
    }
}

// scores of PX pixels against all words: S[l][p] = sum_i x[i][p] * src[i][l]
template <typename T, int LP, int PX>
__device__ __forceinline__ void scores(const T* __restrict__ xb, const float* __restrict__ src_s, int idf, int Q,
                                       float (&S)[LP][PX]) {
#pragma unroll
    for (int l = 0; l < LP; ++l)
#pragma unroll
        for (int p = 0; p < PX; ++p) S[l][p] = 0.f;
#pragma unroll 2
    for (int i = 0; i < idf; ++i) {
        float xv[PX];
        PixIO<T, PX>::load(xb + (size_t)i * Q, xv);
        const float4* s4 = reinterpret_cast<const float4*>(src_s + i * LP);
#pragma unroll
        for (int l4 = 0; l4 < LP / 4; ++l4) {
            const float4 s = s4[l4];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                S[l4 * 4 + 0][p] = fmaf(xv[p], s.x, S[l4 * 4 + 0][p]);
                S[l4 * 4 + 1][p] = fmaf(xv[p], s.y, S[l4 * 4 + 1][p]);
                S[l4 * 4 + 2][p] = fmaf(xv[p], s.z, S[l4 * 4 + 2][p]);
                S[l4 * 4 + 3][p] = fmaf(xv[p], s.w, S[l4 * 4 + 3][p]);
            }
        }
    }
}

// masked softmax over words, in place (GlobalAttention.py:104-109).  A fully masked pixel
// yields NaN exactly like the reference (softmax of all -inf).
template <int LP, int PX>
__device__ __forceinline__ void masked_softmax(float (&S)[LP][PX], const uint32_t (&mb)[PX]) {
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        float m = -INFINITY;
#pragma unroll
        for (int l = 0; l < LP; ++l) {
            const float s = ((mb[p] >> l) & 1u) ? -INFINITY : S[l][p];
            S[l][p] = s;
            m = fmaxf(m, s);
        }
        float sum = 0.f;
#pragma unroll
        for (int l = 0; l < LP; ++l) {
            const float e = __expf(S[l][p] - m);
            S[l][p] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int l = 0; l < LP; ++l) S[l][p] *= inv;
    }
}

template <int PX>
__device__ __forceinline__ void pixel_mask_bits(const uint32_t* __restrict__ mask_bits, int b, int q0, int B, int Q,
                                                int L, int mask_mode, uint32_t (&mb)[PX]) {
    const uint32_t pad = (L < 32) ? ~((1u << L) - 1u) : 0u;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        uint32_t m = pad;
        if (mask_bits != nullptr) m |= __ldg(mask_bits + mask_caption(b, q0 + p, B, Q, mask_mode));
        mb[p] = m;
    }
}

__device__ __forceinline__ void load_src(const float* __restrict__ sb, float* src_s, int idf, int L, int LP, int tid) {
    for (int o = tid; o < idf * LP; o += kThreads) {
        const int i = o / LP, l = o - i * LP;
        src_s[o] = (l < L) ? __ldg(sb + i * L + l) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// forward: c_code and attn in one pass over x
// ---------------------------------------------------------------------------------------
template <typename T, int LP, int PX>
__global__ void __launch_bounds__(kThreads)
k_attn_fwd(const T* __restrict__ x, const float* __restrict__ srcT, const uint32_t* __restrict__ mask_bits,
           T* __restrict__ c_code, T* __restrict__ attn, int B, int idf, int L, int Q, int mask_mode) {
    extern __shared__ __align__(16) float src_s[];  // [idf][LP], zero padded beyond L
    const int b = blockIdx.y, tid = threadIdx.x;
    load_src(srcT + (size_t)b * idf * L, src_s, idf, L, LP, tid);
    __syncthreads();
    const int q0 = (blockIdx.x * kThreads + tid) * PX;
    if (q0 >= Q) return;

    float S[LP][PX];
    scores<T, LP, PX>(x + (size_t)b * idf * Q + q0, src_s, idf, Q, S);
    uint32_t mb[PX];
    pixel_mask_bits<PX>(mask_bits, b, q0, B, Q, L, mask_mode, mb);
    masked_softmax<LP, PX>(S, mb);

    T* ab = attn + (size_t)b * L * Q + q0;
#pragma unroll
    for (int l = 0; l < LP; ++l)
        if (l < L) PixIO<T, PX>::store(ab + (size_t)l * Q, S[l]);

    T* cb = c_code + (size_t)b * idf * Q + q0;
#pragma unroll 2
    for (int i = 0; i < idf; ++i) {
        float acc[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[p] = 0.f;
        const float4* s4 = reinterpret_cast<const float4*>(src_s + i * LP);
#pragma unroll
        for (int l4 = 0; l4 < LP / 4; ++l4) {
            const float4 s = s4[l4];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                acc[p] = fmaf(s.x, S[l4 * 4 + 0][p], acc[p]);
                acc[p] = fmaf(s.y, S[l4 * 4 + 1][p], acc[p]);
                acc[p] = fmaf(s.z, S[l4 * 4 + 2][p], acc[p]);
                acc[p] = fmaf(s.w, S[l4 * 4 + 3][p], acc[p]);
            }
        }
        PixIO<T, PX>::store(cb + (size_t)i * Q, acc);
    }
}

// 32-lane butterfly reduce-scatter: on return lane j holds sum over lanes of v[j] in v[0].
// 31 shuffles for 32 values (a plain per-value butterfly would need 160).
__device__ __forceinline__ void reduce_scatter32(float (&v)[32], int lane) {
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

// ---------------------------------------------------------------------------------------
// backward: recompute P from x and srcT, then dX and the per-sample dSrc reduction
// ---------------------------------------------------------------------------------------
template <typename T, int LP, int PX, bool HAS_GA>
__global__ void __launch_bounds__(kThreads)
k_attn_bwd(const T* __restrict__ x, const float* __restrict__ srcT, const uint32_t* __restrict__ mask_bits,
           const T* __restrict__ g_c, const T* __restrict__ g_attn, T* __restrict__ dX, float* __restrict__ dSrc,
           int B, int idf, int L, int Q, int mask_mode) {
    extern __shared__ __align__(16) float smem[];
    float* src_s = smem;               // [idf][LP]
    float* dsrc_s = smem + idf * LP;   // [idf][32]
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    load_src(srcT + (size_t)b * idf * L, src_s, idf, L, LP, tid);
    for (int o = tid; o < idf * 32; o += kThreads) dsrc_s[o] = 0.f;
    __syncthreads();
    const int q0 = (blockIdx.x * kThreads + tid) * PX;
    const bool valid = q0 < Q;
    const size_t row0 = (size_t)b * idf * Q + (valid ? q0 : 0);
    const T* xb = x + row0;
    const T* gb = g_c + row0;

    float P[LP][PX], dS[LP][PX];
#pragma unroll
    for (int l = 0; l < LP; ++l)
#pragma unroll
        for (int p = 0; p < PX; ++p) { P[l][p] = 0.f; dS[l][p] = 0.f; }

    if (valid) {
        scores<T, LP, PX>(xb, src_s, idf, Q, P);
        uint32_t mb[PX];
        pixel_mask_bits<PX>(mask_bits, b, q0, B, Q, L, mask_mode, mb);
        masked_softmax<LP, PX>(P, mb);
        // dP[l][p] = sum_i g[i][p] * src[i][l]  (+ g_attn)
        scores<T, LP, PX>(gb, src_s, idf, Q, dS);
        if (HAS_GA) {
            const T* ga = g_attn + (size_t)b * L * Q + q0;
#pragma unroll
            for (int l = 0; l < LP; ++l)
                if (l < L) {
                    float t[PX];
                    PixIO<T, PX>::load(ga + (size_t)l * Q, t);
#pragma unroll
                    for (int p = 0; p < PX; ++p) dS[l][p] += t[p];
                }
        }
        // dS = P * (dP - sum_l P dP); masked / padded words have P = 0 (and dP finite)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            float dot = 0.f;
#pragma unroll
            for (int l = 0; l < LP; ++l) dot = fmaf(P[l][p], dS[l][p], dot);
#pragma unroll
            for (int l = 0; l < LP; ++l) dS[l][p] = P[l][p] * (dS[l][p] - dot);
        }
        // dX[i][p] = sum_l dS[l][p] * src[i][l]
        T* dxb = dX + row0;
#pragma unroll 2
        for (int i = 0; i < idf; ++i) {
            float acc[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[p] = 0.f;
            const float4* s4 = reinterpret_cast<const float4*>(src_s + i * LP);
#pragma unroll
            for (int l4 = 0; l4 < LP / 4; ++l4) {
                const float4 s = s4[l4];
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    acc[p] = fmaf(s.x, dS[l4 * 4 + 0][p], acc[p]);
                    acc[p] = fmaf(s.y, dS[l4 * 4 + 1][p], acc[p]);
                    acc[p] = fmaf(s.z, dS[l4 * 4 + 2][p], acc[p]);
                    acc[p] = fmaf(s.w, dS[l4 * 4 + 3][p], acc[p]);
                }
            }
            PixIO<T, PX>::store(dxb + (size_t)i * Q, acc);
        }
    }

    // dSrc[i][l] += sum_p g[i][p] P[l][p] + x[i][p] dS[l][p], reduced over the warp's pixels
    for (int i = 0; i < idf; ++i) {
        float xv[PX], gv[PX];
        if (valid) {
            PixIO<T, PX>::load(xb + (size_t)i * Q, xv);
            PixIO<T, PX>::load(gb + (size_t)i * Q, gv);
        } else {
#pragma unroll
            for (int p = 0; p < PX; ++p) { xv[p] = 0.f; gv[p] = 0.f; }
        }
        float v[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) {
            float a = 0.f;
            if (l < LP) {
#pragma unroll
                for (int p = 0; p < PX; ++p) a = fmaf(gv[p], P[l][p], fmaf(xv[p], dS[l][p], a));
            }
            v[l] = a;
        }
        reduce_scatter32(v, lane);
        if (lane < L) atomicAdd(dsrc_s + i * 32 + lane, v[0]);
    }
    __syncthreads();
    float* db = dSrc + (size_t)b * idf * L;
    for (int o = tid; o < idf * L; o += kThreads) {
        const int i = o / L, l = o - i * L;
        atomicAdd(db + o, dsrc_s[i * 32 + l]);
    }
}

// dW[i][c] = sum_b sum_l dSrc[b][i][l] * ctx[b][c][l]
__global__ void __launch_bounds__(256)
k_dw(const float* __restrict__ ctx, const float* __restrict__ dSrc, float* __restrict__ dW, int B, int idf, int cdf,
     int L) {
    const int i = blockIdx.x;
    for (int c = threadIdx.x; c < cdf; c += blockDim.x) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) {
            const float* d = dSrc + ((size_t)b * idf + i) * L;
            const float* cr = ctx + ((size_t)b * cdf + c) * L;
            float a = 0.f;
            for (int l = 0; l < L; ++l) a = fmaf(__ldg(d + l), __ldg(cr + l), a);
            acc += a;
        }
        dW[(size_t)i * cdf + c] = acc;
    }
}

// dCtx[b][c][l] = sum_i W[i][c] * dSrc[b][i][l]
__global__ void __launch_bounds__(256)
k_dctx(const float* __restrict__ W, const float* __restrict__ dSrc, float* __restrict__ dCtx, int idf, int cdf, int L) {
    extern __shared__ __align__(16) float ds[];  // [idf][L]
    const int b = blockIdx.x;
    for (int o = threadIdx.x; o < idf * L; o += blockDim.x) ds[o] = dSrc[(size_t)b * idf * L + o];
    __syncthreads();
    for (int o = threadIdx.x; o < cdf * L; o += blockDim.x) {
        const int c = o / L, l = o - c * L;
        float acc = 0.f;
        for (int i = 0; i < idf; ++i) acc = fmaf(__ldg(W + (size_t)i * cdf + c), ds[i * L + l], acc);
        dCtx[(size_t)b * cdf * L + o] = acc;
    }
}

inline int padded_words(int L) { return L <= 8 ? 8 : L <= 16 ? 16 : L <= 24 ? 24 : 32; }

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T, int LP, int PX>
int launch_fwd(const void* x, const float* srcT, const uint32_t* mb, void* c, void* a, const AttnShape& s,
               cudaStream_t st) {
    const size_t smem = (size_t)s.idf * LP * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_attn_fwd<T, LP, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(ceil_div(s.Q, kThreads * PX), s.B);
    k_attn_fwd<T, LP, PX><<<grid, kThreads, smem, st>>>(static_cast<const T*>(x), srcT, mb, static_cast<T*>(c),
                                                       static_cast<T*>(a), s.B, s.idf, s.L, s.Q, s.mask_mode);
    add_launches(1);
    return check_launch("attn_fwd(simt)");
}

template <typename T, int LP, int PX>
int launch_bwd(const void* x, const float* srcT, const uint32_t* mb, const void* g, const void* ga, void* dX,
               float* dSrc, const AttnShape& s, cudaStream_t st) {
    const size_t smem = (size_t)s.idf * (LP + 32) * sizeof(float);      // <= 64 KB for idf <= 256 (check_attn_shape)
    dim3 grid(ceil_div(s.Q, kThreads * PX), s.B);
    if (smem > 48 * 1024) {
        cudaError_t e = ga != nullptr
            ? cudaFuncSetAttribute(k_attn_bwd<T, LP, PX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(k_attn_bwd<T, LP, PX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("attn_bwd(simt): %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return SBA_ERR_UNSUPPORTED;
        }
    }
    if (ga != nullptr) {
        k_attn_bwd<T, LP, PX, true><<<grid, kThreads, smem, st>>>(
            static_cast<const T*>(x), srcT, mb, static_cast<const T*>(g), static_cast<const T*>(ga),
            static_cast<T*>(dX), dSrc, s.B, s.idf, s.L, s.Q, s.mask_mode);
    } else {
        k_attn_bwd<T, LP, PX, false><<<grid, kThreads, smem, st>>>(
            static_cast<const T*>(x), srcT, mb, static_cast<const T*>(g), nullptr, static_cast<T*>(dX), dSrc, s.B,
            s.idf, s.L, s.Q, s.mask_mode);
    }
    add_launches(1);
    return check_launch("attn_bwd(simt)");
}

template <typename T>
int dispatch_fwd(const void* x, const float* srcT, const uint32_t* mb, void* c, void* a, const AttnShape& s,
                 cudaStream_t st) {
    const int LP = padded_words(s.L);
    const bool vec = (s.Q % 4 == 0) && aligned(x, 16) && aligned(c, 16) && aligned(a, 16);
#define SBA_FWD(LPV, PXV) return launch_fwd<T, LPV, PXV>(x, srcT, mb, c, a, s, st)
    if (vec) {
        switch (LP) {
            case 8: SBA_FWD(8, 4);
            case 16: SBA_FWD(16, 4);
            case 24: SBA_FWD(24, 4);
            default: SBA_FWD(32, 2);
        }
    }
    switch (LP) {
        case 8: SBA_FWD(8, 1);
        case 16: SBA_FWD(16, 1);
        case 24: SBA_FWD(24, 1);
        default: SBA_FWD(32, 1);
    }
#undef SBA_FWD
}

template <typename T>
int dispatch_bwd(const void* x, const float* srcT, const uint32_t* mb, const void* g, const void* ga, void* dX,
                 float* dSrc, const AttnShape& s, cudaStream_t st) {
    const int LP = padded_words(s.L);
    const bool vec = (s.Q % 4 == 0) && aligned(x, 16) && aligned(g, 16) && aligned(dX, 16) &&
                     (ga == nullptr || aligned(ga, 16));
#define SBA_BWD(LPV, PXV) return launch_bwd<T, LPV, PXV>(x, srcT, mb, g, ga, dX, dSrc, s, st)
    if (vec) {
        switch (LP) {
            case 8: SBA_BWD(8, 4);
            case 16: SBA_BWD(16, 2);
            case 24: SBA_BWD(24, 2);
            default: SBA_BWD(32, 1);
        }
    }
    switch (LP) {
        case 8: SBA_BWD(8, 1);
        case 16: SBA_BWD(16, 1);
        case 24: SBA_BWD(24, 1);
        default: SBA_BWD(32, 1);
    }
#undef SBA_BWD
}

}  // namespace

int simt_project(const float* ctx, const float* W, const uint8_t* mask, float* srcT, uint32_t* mask_bits,
                 const AttnShape& s, cudaStream_t st) {
    const size_t smem = (size_t)s.cdf * s.L * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("project: cdf*L = %d words of context do not fit in shared memory", s.cdf * s.L);
        return SBA_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_project<<<s.B, 256, smem, st>>>(ctx, W, mask, srcT, mask_bits, s.idf, s.cdf, s.L);
    add_launches(1);
    return check_launch("project");
}

int simt_attn_fwd(const void* x, const float* srcT, const uint32_t* mask_bits, void* c_code, void* attn,
                  const AttnShape& s, cudaStream_t st) {
    if (s.dtype == SBA_BF16) return dispatch_fwd<__nv_bfloat16>(x, srcT, mask_bits, c_code, attn, s, st);
    return dispatch_fwd<float>(x, srcT, mask_bits, c_code, attn, s, st);
}

int simt_attn_bwd(const void* x, const float* srcT, const uint32_t* mask_bits, const void* g_c, const void* g_attn,
                  void* dX, float* dSrc, const AttnShape& s, cudaStream_t st) {
    if (s.dtype == SBA_BF16) return dispatch_bwd<__nv_bfloat16>(x, srcT, mask_bits, g_c, g_attn, dX, dSrc, s, st);
    return dispatch_bwd<float>(x, srcT, mask_bits, g_c, g_attn, dX, dSrc, s, st);
}

int simt_attn_bwd_epilogue(const float* ctx, const float* W, const float* dSrc, float* dW, float* dCtx,
                           const AttnShape& s, cudaStream_t st) {
    if (dW != nullptr) {
        k_dw<<<s.idf, 256, 0, st>>>(ctx, dSrc, dW, s.B, s.idf, s.cdf, s.L);
        add_launches(1);
        int rc = check_launch("dW");
        if (rc) return rc;
    }
    if (dCtx != nullptr) {
        const size_t smem = (size_t)s.idf * s.L * sizeof(float);
        k_dctx<<<s.B, 256, smem, st>>>(W, dSrc, dCtx, s.idf, s.cdf, s.L);
        add_launches(1);
        int rc = check_launch("dCtx");
        if (rc) return rc;
    }
    return SBA_OK;
}

}  // namespace sba

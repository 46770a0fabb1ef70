// Zero-fill and dW / dContext epilogue kernels around the mma.sync backward (attn_mma_bwd.cu), which
// accumulates dSrc[b] with fp32 atomics: a small grid in front clears dSrc / dW, a small grid behind forms
// dW = sum_b dSrc[b] . ctx[b]^T (atomics over 64 sample groups) and dCtx[b] = W^T . dSrc[b].  Both are
// chained by programmatic dependent launch.  (The tcgen05 backward does without either: attn_tc5_bwd.cu.)
#include "host_util.h"
#include "kernels.h"

namespace sba {
namespace {

// zero dSrc, the per-sample counters and dW; the streaming kernel waits for this grid only before its
// first atomic (griddepcontrol.wait), so the fill overlaps its prologue and first tiles
__global__ void __launch_bounds__(256) k_bwd_zero(float* __restrict__ a, size_t na, float* __restrict__ b, size_t nb) {
    // programmatic dependent of whatever precedes it (hides its launch latency); upstream work is complete
    // before its own dependent - the streaming kernel, which reads x / g_c at once - may start
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < na; i += step) a[i] = 0.f;
    if (b != nullptr)
        for (size_t i = i0; i < nb; i += step) b[i] = 0.f;
}

// dW += sum_{b in group} dSrc[b] . ctx[b]^T for a [idf x 32] slice, dCtx[b] = W^T . dSrc[b]; a programmatic
// dependent of the streaming kernel (griddepcontrol.wait = that grid is complete and flushed).
//   blocks [0, n_dw)          : (32 input channels c, one of 64 sample groups).  The K = (sample, word) axis
//                               of up to 4 samples is staged flat - ds [idf][K], cs [K][32 c] - so the inner
//                               loop is branch-free: one broadcast LDS + one LDS.128 per 4 FMAs per thread
//                               (thread = channel i x 4 channels c).  The 64 groups meet in fp32 atomics on
//                               dW (dW zeroed by k_bwd_zero); many small blocks hide each other's latency.
//   blocks [n_dw, n_dw + B)   : dCtx of one sample (only when words need a gradient)
constexpr int kPostCS = 36;        // row stride (floats) of cs: 16-byte aligned rows, 4-bank skew
constexpr int kPostDS = 132;       // row stride (floats) of ds: K <= 128, 4-bank skew between channels
template <int IDF>
__global__ void __launch_bounds__(256) k_bwd_post_atomic(const float* __restrict__ dSrc, const float* __restrict__ ctx,
                                                      const float* __restrict__ W, float* __restrict__ dW,
                                                      float* __restrict__ dCtx, int B, int cdf, int L, int n_dw) {
    extern __shared__ __align__(16) float sm[];
    constexpr bool HI = IDF > 32;          // a thread owns channel i0 and, for idf > 32, i0 + 32
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // (the head kernel of the next call, see above)
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
}

}  // namespace

int attn_bwd_zero(float* dSrc, size_t n_src, float* dW, size_t n_dw, cudaStream_t st) {
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
        cudaError_t ze = cudaLaunchKernelEx(&zc, k_bwd_zero, dSrc, n_src, dW, n_dw);
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
    int dev = 0, sms = 0;
    int rc0 = current_device(&dev, &sms, "attn_bwd(post)");
    if (rc0) return rc0;
    static std::atomic<unsigned long long> done32{0}, done48{0}, done64{0};
    rc0 = ensure_dynamic_smem(k_bwd_post_atomic<32>, cfg.dynamicSmemBytes, dev, done32, "attn_bwd(post)");
    if (!rc0) rc0 = ensure_dynamic_smem(k_bwd_post_atomic<48>, cfg.dynamicSmemBytes, dev, done48, "attn_bwd(post)");
    if (!rc0) rc0 = ensure_dynamic_smem(k_bwd_post_atomic<64>, cfg.dynamicSmemBytes, dev, done64, "attn_bwd(post)");
    if (rc0) return rc0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (idf == 32) e = cudaLaunchKernelEx(&cfg, k_bwd_post_atomic<32>, dSrc, ctx, W, dW, dCtx, B, cdf, L, n_dw);
    else if (idf == 48) e = cudaLaunchKernelEx(&cfg, k_bwd_post_atomic<48>, dSrc, ctx, W, dW, dCtx, B, cdf, L, n_dw);
    else if (idf == 64) e = cudaLaunchKernelEx(&cfg, k_bwd_post_atomic<64>, dSrc, ctx, W, dW, dCtx, B, cdf, L, n_dw);
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

}  // namespace sba

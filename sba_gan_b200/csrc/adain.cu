// the caller's AdaIN + concatenation right behind the attention
// (NEXT_STAGE_G.forward, AttnGAN2/code/model_bert.py:458-461; ADAIN_NORM.forward, :367-374):
//     style = Linear(w_code);  gamma, beta = style.chunk(2, 1)
//     h     = (gamma + 1) * InstanceNorm2d(h_code) + beta          (eps 1e-5, biased variance, no running stats)
//     h_c   = cat((h, c_code), 1)
// The reference runs this as instance-norm statistics + normalise + scale + shift + cat: ~10 passes over [B, idf, Q].
// Here one block owns one (sample, channel) row of Q pixels, keeps it in registers, and writes the normalised row
// straight into rows [0, idf) of the concatenated buffer: 1 read + 1 write; the attention forward writes c_code into
// rows [idf, 2 idf) of the same buffer (sba_attn_fwd_into), so the concatenation never exists as a copy.
// Backward likewise: one block per row reads x and the gradient slice in place, and ADDS its dX to the attention's.
// Bound: HBM (2 x idf x Q x s bytes per sample forward, 4 x backward).
#include "kernels.h"

namespace sba {
namespace {

constexpr int kAdThreads = 512;
constexpr int kAdCache = 32;          // row elements a thread keeps in registers: rows up to 512 * 32 = 16384 pixels

template <typename T> struct Vec;     // 16-byte vectors of row elements
template <> struct Vec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
            v[2 * k] = f.x; v[2 * k + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// sum over the block, result in every thread (fixed order: shuffles, then the per-warp partials in warp order)
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kAdThreads / 32; ++w) a += red[w * NV + k];
        v[k] = a;
    }
}

// Row access: VECTOR = the row base and length allow 16-byte accesses; rows longer than the register cache are
// re-read from memory (L2) in the later passes.
template <typename T, bool VECTOR>
struct Row {
    static constexpr int N = VECTOR ? Vec<T>::N : 1;
    static constexpr int NIT = kAdCache / N;
    const T* p;
    int Q;
    float c[kAdCache];
    __device__ __forceinline__ int idx(int it) const { return (it * kAdThreads + (int)threadIdx.x) * N; }
    __device__ __forceinline__ void load() {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = idx(it);
            if (e < Q) {
                if constexpr (VECTOR) {
                    float v[N];
                    Vec<T>::load(p + e, v);
#pragma unroll
                    for (int k = 0; k < N; ++k) c[it * N + k] = v[k];
                } else {
                    c[it] = ld1(p + e);
                }
            } else {
#pragma unroll
                for (int k = 0; k < N; ++k) c[it * N + k] = 0.f;
            }
        }
    }
    // elements beyond the cache (rows longer than kAdThreads * kAdCache): streamed with f(value)
    template <class F>
    __device__ __forceinline__ void tail(F f) const {
        for (int e = kAdThreads * kAdCache + (int)threadIdx.x; e < Q; e += kAdThreads) f(e, ld1(p + e));
    }
};

template <typename T, bool VECTOR>
__global__ void __launch_bounds__(kAdThreads) k_adain_fwd(const T* __restrict__ x, const float* __restrict__ style,
                                                          T* __restrict__ out, float* __restrict__ stats, int C, int Q,
                                                          int out_rows, int out_row0, float eps) {
    __shared__ float red[2 * kAdThreads / 32];
    const int b = blockIdx.x / C, ch = blockIdx.x - b * C;
    Row<T, VECTOR> r;
    r.p = x + (size_t)blockIdx.x * Q;
    r.Q = Q;
    r.load();
    // mean, then the centred second moment (two passes over the registers: no cancellation)
    float s[1] = {0.f};
#pragma unroll
    for (int k = 0; k < kAdCache; ++k) s[0] += r.c[k];
    r.tail([&](int, float v) { s[0] += v; });
    block_sum<1>(s, red);
    const float mean = s[0] / (float)Q;
    float q[1] = {0.f};
#pragma unroll
    for (int it = 0; it < Row<T, VECTOR>::NIT; ++it)
#pragma unroll
        for (int k = 0; k < Row<T, VECTOR>::N; ++k)
            if (r.idx(it) + k < Q) { const float d = r.c[it * Row<T, VECTOR>::N + k] - mean; q[0] = fmaf(d, d, q[0]); }
    r.tail([&](int, float v) { const float d = v - mean; q[0] = fmaf(d, d, q[0]); });
    block_sum<1>(q, red);
    const float rstd = rsqrtf(q[0] / (float)Q + eps);        // biased variance, like nn.InstanceNorm2d
    const float g = style[(size_t)b * 2 * C + ch] + 1.f, be = style[(size_t)b * 2 * C + C + ch];
    const float a = g * rstd, sh = be - mean * a;             // y = a x + sh
    if (threadIdx.x == 0) { stats[2 * blockIdx.x] = mean; stats[2 * blockIdx.x + 1] = rstd; }
    T* o = out + ((size_t)b * out_rows + out_row0 + ch) * Q;
#pragma unroll
    for (int it = 0; it < Row<T, VECTOR>::NIT; ++it) {
        const int e = r.idx(it);
        if (e < Q) {
            if constexpr (VECTOR) {
                float v[Row<T, VECTOR>::N];
#pragma unroll
                for (int k = 0; k < Row<T, VECTOR>::N; ++k) v[k] = fmaf(r.c[it * Row<T, VECTOR>::N + k], a, sh);
                Vec<T>::store(o + e, v);
            } else {
                st1(o + e, fmaf(r.c[it], a, sh));
            }
        }
    }
    r.tail([&](int e, float v) { st1(o + e, fmaf(v, a, sh)); });
}

// g = gradient of the normalised rows, read in place from rows [g_row0, g_row0 + C) of g_buf [B, g_rows, Q]
//   d_beta = sum g,  d_gamma = sum g xhat,  dx = (gamma + 1) rstd (g - mean(g) - xhat mean(g xhat));  dX (+)= dx
template <typename T, bool VECTOR>
__global__ void __launch_bounds__(kAdThreads) k_adain_bwd(const T* __restrict__ x, const float* __restrict__ style,
                                                          const float* __restrict__ stats, const T* __restrict__ g_buf,
                                                          T* __restrict__ dX, float* __restrict__ d_style, int C, int Q,
                                                          int g_rows, int g_row0, int accumulate) {
    __shared__ float red[2 * kAdThreads / 32];
    const int b = blockIdx.x / C, ch = blockIdx.x - b * C;
    const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
    Row<T, VECTOR> rx, rg;
    rx.p = x + (size_t)blockIdx.x * Q;
    rg.p = g_buf + ((size_t)b * g_rows + g_row0 + ch) * Q;
    rx.Q = rg.Q = Q;
    rx.load();
    rg.load();
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kAdCache; ++k) {
        rx.c[k] = (rx.c[k] - mean) * rstd;                    // xhat (elements beyond Q hold g = 0)
        s[0] += rg.c[k];
        s[1] = fmaf(rg.c[k], rx.c[k], s[1]);
    }
    rx.tail([&](int e, float v) { const float gg = ld1(rg.p + e); s[0] += gg; s[1] = fmaf(gg, (v - mean) * rstd, s[1]); });
    block_sum<2>(s, red);
    if (threadIdx.x == 0) {
        d_style[(size_t)b * 2 * C + ch] = s[1];               // d gamma
        d_style[(size_t)b * 2 * C + C + ch] = s[0];           // d beta
    }
    const float a = (style[(size_t)b * 2 * C + ch] + 1.f) * rstd;
    const float mg = s[0] / (float)Q, mgx = s[1] / (float)Q;
    T* o = dX + (size_t)blockIdx.x * Q;
    constexpr int N = Row<T, VECTOR>::N;
#pragma unroll
    for (int it = 0; it < Row<T, VECTOR>::NIT; ++it) {
        const int e = rx.idx(it);
        if (e < Q) {
            if constexpr (VECTOR) {
                float v[N], prev[N];
                if (accumulate) Vec<T>::load(o + e, prev);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    v[k] = a * (rg.c[it * N + k] - mg - rx.c[it * N + k] * mgx);
                    if (accumulate) v[k] += prev[k];
                }
                Vec<T>::store(o + e, v);
            } else {
                float v = a * (rg.c[it] - mg - rx.c[it] * mgx);
                if (accumulate) v += ld1(o + e);
                st1(o + e, v);
            }
        }
    }
    rx.tail([&](int e, float xv) {
        float v = a * (ld1(rg.p + e) - mg - (xv - mean) * rstd * mgx);
        if (accumulate) v += ld1(o + e);
        st1(o + e, v);
    });
}

inline bool vec_ok(const void* p, int Q, int es) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && ((size_t)Q * es) % 16 == 0;
}

}  // namespace

int adain_fwd(const void* x, const float* style, void* out, int out_rows, int out_row0, float* stats, int B, int C, int Q,
              int dtype, float eps, cudaStream_t st) {
    const int es = dtype == SBA_F32 ? 4 : 2;
    const bool v = vec_ok(x, Q, es) && vec_ok(out, Q, es);
    const dim3 grid(B * C), block(kAdThreads);
    if (dtype == SBA_F32) {
        auto xs = static_cast<const float*>(x);
        auto os = static_cast<float*>(out);
        if (v) k_adain_fwd<float, true><<<grid, block, 0, st>>>(xs, style, os, stats, C, Q, out_rows, out_row0, eps);
        else k_adain_fwd<float, false><<<grid, block, 0, st>>>(xs, style, os, stats, C, Q, out_rows, out_row0, eps);
    } else {
        auto xs = static_cast<const __nv_bfloat16*>(x);
        auto os = static_cast<__nv_bfloat16*>(out);
        if (v) k_adain_fwd<__nv_bfloat16, true><<<grid, block, 0, st>>>(xs, style, os, stats, C, Q, out_rows, out_row0, eps);
        else k_adain_fwd<__nv_bfloat16, false><<<grid, block, 0, st>>>(xs, style, os, stats, C, Q, out_rows, out_row0, eps);
    }
    add_launches(1);
    return check_launch("adain_fwd");
}

int adain_bwd(const void* x, const float* style, const float* stats, const void* g_buf, int g_rows, int g_row0, void* dX,
              int accumulate, float* d_style, int B, int C, int Q, int dtype, cudaStream_t st) {
    const int es = dtype == SBA_F32 ? 4 : 2;
    const bool v = vec_ok(x, Q, es) && vec_ok(g_buf, Q, es) && vec_ok(dX, Q, es);
    const dim3 grid(B * C), block(kAdThreads);
    if (dtype == SBA_F32) {
        auto xs = static_cast<const float*>(x);
        auto gs = static_cast<const float*>(g_buf);
        auto ds = static_cast<float*>(dX);
        if (v) k_adain_bwd<float, true><<<grid, block, 0, st>>>(xs, style, stats, gs, ds, d_style, C, Q, g_rows, g_row0, accumulate);
        else k_adain_bwd<float, false><<<grid, block, 0, st>>>(xs, style, stats, gs, ds, d_style, C, Q, g_rows, g_row0, accumulate);
    } else {
        auto xs = static_cast<const __nv_bfloat16*>(x);
        auto gs = static_cast<const __nv_bfloat16*>(g_buf);
        auto ds = static_cast<__nv_bfloat16*>(dX);
        if (v) k_adain_bwd<__nv_bfloat16, true><<<grid, block, 0, st>>>(xs, style, stats, gs, ds, d_style, C, Q, g_rows, g_row0, accumulate);
        else k_adain_bwd<__nv_bfloat16, false><<<grid, block, 0, st>>>(xs, style, stats, gs, ds, d_style, C, Q, g_rows, g_row0, accumulate);
    }
    add_launches(1);
    return check_launch("adain_bwd");
}

}  // namespace sba

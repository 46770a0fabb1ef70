// Building blocks of the tensor-core attention family (attn_mma_fwd.cu / attn_mma_bwd.cu):
// mbarrier + 1-D TMA bulk copies, warp-level mma.sync fragments, error-compensated
// fp16 splitting, movmatrix transposes.
//
// Why mma.sync and not tcgen05: the contractions here are 16 pixels x (18..32 words) x
// (32 channels) per warp - far below a tcgen05 tile - and tools/ubench shows that CUDA-core
// FFMA tops out at ~35 TFMA/s with the shared-memory operand path (0.41 broadcast LDS.128
// per clock per SM) saturating at the same time, which caps a SIMT kernel at ~45 % of the
// HBM roofline.  Warp-level HMMA runs at 547 TFLOP/s on B200 and keeps the word-feature
// operand in registers, so three fp16 MMAs (hi*hi + hi*lo + lo*hi, ~2^-22 relative) stay
// cheaper than the HBM time of the tile.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace sba {
namespace mma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / TMA bulk copy (1-D, no tensor map needed: one channel row is contiguous) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, both 16B aligned)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- warp-level tensor-core MMAs, fp32 accumulate ----------------------------------------
// HALF = true: fp16 operands; false: bf16 operands
template <bool HALF>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    if constexpr (HALF) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
}
template <bool HALF>
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    if constexpr (HALF) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(b0));
    } else {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(b0));
    }
}

// 8x8 b16 transpose across the warp: in: lane (g,c) holds row g, cols 2c,2c+1; out: the same of M^T
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

// ---- error-compensated fp16 split: x*sc = hi + lo (+ ~2^-22 |x*sc|) ----------------------
__device__ __forceinline__ void split2(float x0, float x1, float sc, uint32_t& hi, uint32_t& lo) {
    const float s0 = x0 * sc, s1 = x1 * sc;
    const __half2 h = __floats2half2_rn(s0, s1);   // .x (low 16 bits) = s0
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_bf16(float x0, float x1) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// power-of-two scale that brings `amax` into [2^12, 2^13) (so fp16 hi/lo keep ~22 bits for
// everything within 2^-14 of the tile maximum); returns 1 for zero / tiny / non-finite maxima
__device__ __forceinline__ void pow2_scale(float amax, float& sc, float& inv) {
    const uint32_t e = (__float_as_uint(amax) >> 23) & 0xffu;
    if (e >= 13u && e <= 250u) {
        sc = __uint_as_float((266u - e) << 23);
        inv = __uint_as_float((e - 12u) << 23);
    } else {
        sc = 1.f;
        inv = 1.f;
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}


// Channel held by MMA k-index of k-step ks on the fp32 path: lane c owns channels c, c+4, c+8,
// c+12 of the step (k = 2c, 2c+1, 2c+8, 2c+9) so that its four LDS.32 gathers hit four
// different 8-bank groups.  The same permutation is applied to the sourceT operand.
__device__ __forceinline__ int chan_of(int ks, int c, int j) { return 16 * ks + c + 4 * j; }

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// max over the warp of non-negative floats (their bit patterns order like unsigned integers)
__device__ __forceinline__ float warp_absmax_redux(float nonneg) {
    uint32_t r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(__float_as_uint(nonneg)));
    return __uint_as_float(r);
}

constexpr int TQ = 128;            // pixels per CTA tile: 8 consumer warps x one 16-pixel m-tile
constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kThreads = kConsumers + 32;   // + one TMA producer warp

// Shared-memory row stride (in elements) of a staged [channel][pixel] tile.
// fp32: stride = 8 (mod 32) words makes both fragment gathers conflict free
//       (LDS.32 at ch*RS + g with ch = c + 4j, and LDS.64 at ch*RS + 2c with ch = g).
// bf16: stride*2 bytes = 16 (mod 128) makes ldmatrix rows hit distinct 16-byte bank groups.
template <typename T> struct TileStride;
template <> struct TileStride<float> { static constexpr int value = TQ + 8; };
template <> struct TileStride<__nv_bfloat16> { static constexpr int value = TQ + 8; };

}  // namespace mma
}  // namespace sba

// Shared device/host helpers for the sm_100a word-region attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

#include "../../include/sba_attn.h"

namespace sba {

constexpr int kMaxWords = 32;   // caption mask is one 32-bit word per caption

// thread-local status shared by all translation units (defined in abi.cu)
void set_error(const char* fmt, ...);
void add_launches(int n);
int check_launch(const char* what);

// ---------------------------------------------------------------------------------------
// element I/O: PX consecutive pixels of one row, as fp32 in registers
// ---------------------------------------------------------------------------------------
template <typename T, int PX> struct PixIO;

template <> struct PixIO<float, 1> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <> struct PixIO<float, 2> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[2]) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
};
template <> struct PixIO<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct PixIO<__nv_bfloat16, 1> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) { v[0] = __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) { *p = __float2bfloat16_rn(v[0]); }
};
template <> struct PixIO<__nv_bfloat16, 2> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[2]) {
        __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(p);
        float2 f = __bfloat1622float2(t); v[0] = f.x; v[1] = f.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[2]) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
    }
};
template <> struct PixIO<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
        float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t*>(&a);
        raw.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = raw;
    }
};

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Which caption's padding mask applies to pixel (b, q): see SBA_MASK_* in sba_attn.h.
__device__ __forceinline__ int mask_caption(int b, int q, int B, int Q, int mask_mode) {
    if (mask_mode == SBA_MASK_PER_SAMPLE) return b;
    return (int)(((long long)b * Q + q) % B);
}

}  // namespace sba

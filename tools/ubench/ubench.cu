// Microbenchmarks that decide the kernel design for the word-region attention path on
// B200 (sm_100a): FP32 FMA issue (scalar vs packed f32x2), shared-memory broadcast load
// rates, legacy mma.sync rates, and a plain HBM stream.  Not part of the product.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc[j] = threadIdx.x * 0.001f + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc[j] = fmaf(acc[j], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    float2 acc[ILP];
    float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc[j] = make_float2(threadIdx.x * 0.001f + j, j * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc[j] = __ffma2_rn(acc[j], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 whose multiplicand comes from a broadcast shared-memory load of width W words;
// R = FFMA2 per loaded word-pair.  Models the "src operand from smem" inner loop.
template <int WORDS, int REUSE>
__global__ void k_lds_ffma2(float* out, int iters) {
    __shared__ __align__(16) float tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = 1.0f + i * 1e-6f;
    __syncthreads();
    float2 acc[REUSE * WORDS / 2 > 0 ? REUSE * (WORDS > 1 ? WORDS / 2 : 1) : 1];
    constexpr int NACC = REUSE * (WORDS > 1 ? WORDS / 2 : 1);
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = make_float2(j, threadIdx.x);
    float2 xv = make_float2(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int off = ((it * 8 + u) * WORDS) & 1023;
            if constexpr (WORDS == 1) {
                float s = tab[off];
#pragma unroll
                for (int r = 0; r < REUSE; ++r) acc[r] = __ffma2_rn(xv, make_float2(s, s), acc[r]);
            } else if constexpr (WORDS == 2) {
                float2 s = *reinterpret_cast<const float2*>(&tab[off]);
#pragma unroll
                for (int r = 0; r < REUSE; ++r) acc[r] = __ffma2_rn(xv, s, acc[r]);
            } else {
                float4 s = *reinterpret_cast<const float4*>(&tab[off]);
#pragma unroll
                for (int r = 0; r < REUSE; ++r) {
                    acc[2 * r] = __ffma2_rn(xv, make_float2(s.x, s.y), acc[2 * r]);
                    acc[2 * r + 1] = __ffma2_rn(xv, make_float2(s.z, s.w), acc[2 * r + 1]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// pure broadcast LDS rate
template <int WORDS>
__global__ void k_lds(float* out, int iters) {
    __shared__ __align__(16) float tab[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = i;
    __syncthreads();
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            int off = ((it * 16 + u) * 4) & 2047;
            if constexpr (WORDS == 1) { s0 += tab[off]; }
            else if constexpr (WORDS == 2) { float2 v = *reinterpret_cast<const float2*>(&tab[off]); s0 += v.x; s1 += v.y; }
            else { float4 v = *reinterpret_cast<const float4*>(&tab[off]); s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

// per-lane (conflict-free, non-broadcast) LDS.128
__global__ void k_lds128_lane(float* out, int iters) {
    __shared__ __align__(16) float tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i;
    __syncthreads();
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            int off = (((it * 16 + u) * 128) + lane * 4) & 4095;
            float4 v = *reinterpret_cast<const float4*>(&tab[off]); s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int ILP>
__global__ void k_mma(float* out, int iters) {
    float c[ILP][4];
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
    uint32_t b[2] = {0x3c003c00u, 0x3c003c00u + threadIdx.x};
#pragma unroll
    for (int j = 0; j < ILP; ++j) { c[j][0] = j; c[j][1] = 0; c[j][2] = 0; c[j][3] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if constexpr (KIND == 0) mma_f16(c[j], a, b);
            else if constexpr (KIND == 1) mma_bf16(c[j], a, b);
            else mma_tf32(c[j], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// warp-matrix data movement used by the tensor-core attention kernels
template <int ILP>
__global__ void k_movm(float* out, int iters) {
    uint32_t v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x * 7 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(v[j]) : "r"(v[j]));
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s ^= v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
// KIND 0: ldmatrix.x4.trans, 1: stmatrix.x4.trans, 2: LDG.64 hitting L1 (256 B per warp, 12 KB table)
template <int KIND>
__global__ void k_matmove(float* out, const uint2* __restrict__ tab, int iters) {
    __shared__ __align__(128) uint16_t sm[8 * 32 * 72];
    for (int i = threadIdx.x; i < 8 * 32 * 72; i += blockDim.x) sm[i] = (uint16_t)i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm + warp * 32 * 72 + (lane & 15) * 72 + (lane >> 4) * 8);
    uint32_t a0 = lane, a1 = 1, a2 = 2, a3 = 3, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if constexpr (KIND == 0) {
                uint32_t r0, r1, r2, r3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + (u & 1) * 16 * 72 * 2));
                acc ^= r0 ^ r1 ^ r2 ^ r3;
            } else if constexpr (KIND == 1) {
                asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" :: "r"(base + (u & 1) * 16 * 72 * 2), "r"(a0 + u), "r"(a1), "r"(a2), "r"(a3) : "memory");
            } else {
                const uint2 v = __ldg(tab + ((it * 8 + u) % 48) * 32 + lane);
                acc ^= v.x ^ v.y;
            }
        }
    }
    if (KIND == 1) { __syncthreads(); acc = sm[threadIdx.x]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}
// fp32 -> fp16 hi/lo split (the conversion load of the 3xFP16 scheme), 2 values per iteration
template <int ILP>
__global__ void k_split(float* out, int iters, float sc) {
    float x[ILP], y[ILP]; uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { x[j] = threadIdx.x * 0.37f + j; y[j] = j * 1.1f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const float s0 = x[j] * sc, s1 = y[j] * sc;
            const __half2 h = __floats2half2_rn(s0, s1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
            acc ^= *reinterpret_cast<const uint32_t*>(&h) + *reinterpret_cast<const uint32_t*>(&l);
            x[j] += 1.0f; y[j] += 0.5f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

__global__ void k_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}
__global__ void k_read(const float4* __restrict__ in, float* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    float s = 0;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        s += a.x + b.y + c.z + d.w;
    }
    if (s == 123.456f) out[0] = s;
}
__global__ void k_empty() {}

template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, clk_khz);
    float* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float) * 4));
    const int iters = 4096;
    for (int warps_per_sm : {8, 16, 32}) {
        int threads = 256, blocks = nsm * warps_per_sm * 32 / threads;
        double lanes = (double)blocks * threads;
        float ms = timeit([&] { k_ffma<16><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        printf("{\"bench\": \"ffma_scalar\", \"warps_per_sm\": %d, \"tfma_s\": %.2f, \"fma_per_clk_per_sm_at_1965\": %.1f}\n", warps_per_sm,
               lanes * iters * 16 / ms / 1e9, lanes * iters * 16 / (ms * 1e-3) / nsm / 1.965e9);
        ms = timeit([&] { k_ffma2<16><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        printf("{\"bench\": \"ffma2_packed\", \"warps_per_sm\": %d, \"tfma_s\": %.2f, \"fma_per_clk_per_sm_at_1965\": %.1f}\n", warps_per_sm,
               lanes * iters * 32 / ms / 1e9, lanes * iters * 32 / (ms * 1e-3) / nsm / 1.965e9);
    }
    {
        int threads = 256, blocks = nsm * 2; double lanes = (double)blocks * threads;
        float ms;
#define LF(W, R) ms = timeit([&] { k_lds_ffma2<W, R><<<blocks, threads>>>(out, 1024); }); \
        printf("{\"bench\": \"lds_ffma2\", \"lds_words\": %d, \"ffma2_per_ld\": %d, \"tfma_s\": %.2f, \"lds_per_clk_per_sm_at_1965\": %.3f}\n", W, (W > 1 ? W / 2 : 1) * R, \
               lanes * 1024 * 8 * (W > 1 ? W / 2 : 1) * R * 2 / ms / 1e9, (double)blocks * (threads / 32) * 1024 * 8 / (ms * 1e-3) / nsm / 1.965e9);
        LF(1, 1) LF(1, 2) LF(1, 4) LF(2, 1) LF(2, 2) LF(2, 4) LF(4, 1) LF(4, 2) LF(4, 4)
#define LD(W) ms = timeit([&] { k_lds<W><<<blocks, threads>>>(out, 2048); }); \
        printf("{\"bench\": \"lds_broadcast\", \"words\": %d, \"warp_lds_per_clk_per_sm_at_1965\": %.3f}\n", W, (double)blocks * (threads / 32) * 2048 * 16 / (ms * 1e-3) / nsm / 1.965e9);
        LD(1) LD(2) LD(4)
        ms = timeit([&] { k_lds128_lane<<<blocks, threads>>>(out, 2048); });
        printf("{\"bench\": \"lds128_per_lane\", \"warp_lds_per_clk_per_sm_at_1965\": %.3f}\n", (double)blocks * (threads / 32) * 2048 * 16 / (ms * 1e-3) / nsm / 1.965e9);
    }
    for (int warps_per_sm : {8, 16, 32}) {
        int threads = 256, blocks = nsm * warps_per_sm * 32 / threads; double warps = (double)blocks * threads / 32;
        float ms = timeit([&] { k_mma<0, 8><<<blocks, threads>>>(out, iters); });
        printf("{\"bench\": \"mma_sync_f16_m16n8k16\", \"warps_per_sm\": %d, \"tflops\": %.1f}\n", warps_per_sm, warps * iters * 8 * 2.0 * 16 * 8 * 16 / ms / 1e9);
        ms = timeit([&] { k_mma<1, 8><<<blocks, threads>>>(out, iters); });
        printf("{\"bench\": \"mma_sync_bf16_m16n8k16\", \"warps_per_sm\": %d, \"tflops\": %.1f}\n", warps_per_sm, warps * iters * 8 * 2.0 * 16 * 8 * 16 / ms / 1e9);
        ms = timeit([&] { k_mma<2, 8><<<blocks, threads>>>(out, iters); });
        printf("{\"bench\": \"mma_sync_tf32_m16n8k8\", \"warps_per_sm\": %d, \"tflops\": %.1f}\n", warps_per_sm, warps * iters * 8 * 2.0 * 16 * 8 * 8 / ms / 1e9);
    }

    for (int warps_per_sm : {8, 16}) {
        int threads = 256, blocks = nsm * warps_per_sm * 32 / threads; double warps = (double)blocks * threads / 32;
        float ms = timeit([&] { k_movm<8><<<blocks, threads>>>(out, iters); });
        printf("{\"bench\": \"movmatrix\", \"warps_per_sm\": %d, \"warp_inst_per_clk_per_sm_at_1965\": %.3f}\n", warps_per_sm, warps * iters * 8 / (ms * 1e-3) / nsm / 1.965e9);
        uint2* tab; CK(cudaMalloc(&tab, 48 * 32 * 8)); CK(cudaMemset(tab, 1, 48 * 32 * 8));
        ms = timeit([&] { k_matmove<0><<<blocks, threads>>>(out, tab, 1024); });
        printf("{\"bench\": \"ldmatrix_x4_trans\", \"warps_per_sm\": %d, \"warp_inst_per_clk_per_sm_at_1965\": %.3f}\n", warps_per_sm, warps * 1024 * 8 / (ms * 1e-3) / nsm / 1.965e9);
        ms = timeit([&] { k_matmove<1><<<blocks, threads>>>(out, tab, 1024); });
        printf("{\"bench\": \"stmatrix_x4_trans\", \"warps_per_sm\": %d, \"warp_inst_per_clk_per_sm_at_1965\": %.3f}\n", warps_per_sm, warps * 1024 * 8 / (ms * 1e-3) / nsm / 1.965e9);
        ms = timeit([&] { k_matmove<2><<<blocks, threads>>>(out, tab, 1024); });
        printf("{\"bench\": \"ldg64_l1_hit\", \"warps_per_sm\": %d, \"warp_inst_per_clk_per_sm_at_1965\": %.3f}\n", warps_per_sm, warps * 1024 * 8 / (ms * 1e-3) / nsm / 1.965e9);
        ms = timeit([&] { k_split<8><<<blocks, threads>>>(out, iters, 1.5f); });
        printf("{\"bench\": \"fp16_split_pair\", \"warps_per_sm\": %d, \"warp_pairs_per_clk_per_sm_at_1965\": %.3f}\n", warps_per_sm, warps * iters * 8 / (ms * 1e-3) / nsm / 1.965e9);
        CK(cudaFree(tab));
    }
    {
        size_t n = (size_t)1 << 28;  // 4 GiB of float4? no: 2^28 float4 = 4 GiB; use 2^26 = 1 GiB
        n = (size_t)1 << 26;
        float4 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 1, n * 16)); CK(cudaMemset(b, 0, n * 16));
        for (int bps : {4, 8, 16}) {
            float ms = timeit([&] { k_copy<<<nsm * bps, 256>>>(a, b, n); });
            printf("{\"bench\": \"hbm_copy_float4\", \"blocks_per_sm\": %d, \"gb_s\": %.1f}\n", bps, 2.0 * n * 16 / ms / 1e6);
            ms = timeit([&] { k_read<<<nsm * bps, 256>>>(a, out, n); });
            printf("{\"bench\": \"hbm_read_float4\", \"blocks_per_sm\": %d, \"gb_s\": %.1f}\n", bps, 1.0 * n * 16 / ms / 1e6);
        }
        float ms = timeit([&] { CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice)); });
        printf("{\"bench\": \"hbm_memcpy_d2d\", \"gb_s\": %.1f}\n", 2.0 * n * 16 / ms / 1e6);
        ms = timeit([&] { for (int i = 0; i < 1000; ++i) k_empty<<<1, 32>>>(); }, 3);
        printf("{\"bench\": \"empty_launch\", \"us_per_launch\": %.2f}\n", ms);
    }
    return 0;
}

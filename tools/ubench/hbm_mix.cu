// HBM stream ceilings for the read/write mixes of the attention kernels (development aid, not part
// of the product).  Shapes follow the bf16 forward at 128x128: per pixel 64 B of x in, 64 B of c_code
// and 36 B of attention map out ([B][rows][Q] tensors, Q = 16384 px, rows = 32 / 32 / 18).
//   flat   : grid-stride float4 streams (full 128-byte lines, any mix)
//   box64  : every warp moves [rows x 32 px] boxes = 64-byte row segments (the per-warp TMA store boxes)
//   box256 : every CTA (4 warps) moves [rows x 128 px] boxes = 256-byte row segments
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_flat(const float4* __restrict__ in, float4* __restrict__ o1, float4* __restrict__ o2, size_t n_in, size_t n1,
                       size_t n2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = i0; i < n_in; i += stride) { float4 v = in[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    for (size_t i = i0; i < n1; i += stride) o1[i] = acc;
    for (size_t i = i0; i < n2; i += stride) o2[i] = acc;
}
// interleaved per tile: read the x box, then write the c box and the attention box of the same pixels
template <int SEG16, bool RANGE = false>   // RANGE: every group walks a contiguous range of boxes (the kernels' tile schedule)
// 16-byte chunks per row segment: 4 (64 B, per warp) or 16 (256 B, per CTA of 128 threads)
__global__ void k_box(const uint4* __restrict__ x, uint4* __restrict__ c, uint4* __restrict__ a, int B, int Q, int rd, int w1, int w2) {
    constexpr int GROUP = SEG16 == 4 ? 32 : 128;                // threads that share one box
    constexpr int PX = SEG16 * 8;                               // pixels per box row (bf16)
    const int tid = threadIdx.x % GROUP, grp = (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
    const int ngrp = gridDim.x * blockDim.x / GROUP;
    const int boxes_per_row = Q / PX, nbox = B * boxes_per_row;
    const size_t pitch = (size_t)Q / 8;                         // uint4 per tensor row
    const int rpi = GROUP / SEG16;                              // rows covered per iteration
    const int lo = RANGE ? (int)((long long)grp * nbox / ngrp) : grp, hi = RANGE ? (int)((long long)(grp + 1) * nbox / ngrp) : nbox;
    for (int bx = lo; bx < hi; bx += RANGE ? 1 : ngrp) {
        const int b = bx / boxes_per_row, q16 = (bx % boxes_per_row) * SEG16 + tid % SEG16;
        uint4 acc = make_uint4(0, 0, 0, 0);
        if (rd) for (int r = tid / SEG16; r < 32; r += rpi) { uint4 v = x[((size_t)b * 32 + r) * pitch + q16]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
        if (w2) for (int r = tid / SEG16; r < 18; r += rpi) a[((size_t)b * 18 + r) * pitch + q16] = acc;
        if (w1) for (int r = tid / SEG16; r < 32; r += rpi) c[((size_t)b * 32 + r) * pitch + q16] = acc;
        if (!w1 && !w2 && acc.x == 0x12345u) c[0] = acc;
    }
}
template <typename F>
float timeit(F f, int reps = 8) {
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
    const int B = 1024, Q = 16384;                              // 1 GiB of x: far beyond L2
    const size_t nx = (size_t)B * 32 * Q / 8, nc = nx, na = (size_t)B * 18 * Q / 8;   // in uint4
    uint4 *x, *c, *a; CK(cudaMalloc(&x, nx * 16)); CK(cudaMalloc(&c, nc * 16)); CK(cudaMalloc(&a, na * 16));
    CK(cudaMemset(x, 1, nx * 16)); CK(cudaMemset(c, 0, nc * 16)); CK(cudaMemset(a, 0, na * 16));
    const int nsm = 148;
    struct { const char* name; int rd, w1, w2; } mixes[] = {{"x->c+attn (39R:61W, forward)", 1, 1, 1}, {"x->c (1R:1W)", 1, 1, 0},
                                                           {"x only (read)", 1, 0, 0}, {"c+attn only (write)", 0, 1, 1}};
    for (auto& m : mixes) {
        const double bytes = 16.0 * (m.rd * nx + m.w1 * nc + m.w2 * na);
        float ms = timeit([&] { k_flat<<<nsm * 16, 256>>>((const float4*)x, (float4*)c, (float4*)a, m.rd ? nx : 0, m.w1 ? nc : 0, m.w2 ? na : 0); });
        printf("{\"bench\": \"hbm_mix\", \"mix\": \"%s\", \"pattern\": \"flat, phases not interleaved\", \"gb_s\": %.0f}\n", m.name, bytes / ms / 1e6);
        for (int bps : {8, 16}) {
            ms = timeit([&] { k_box<4><<<nsm * bps, 128>>>(x, c, a, B, Q, m.rd, m.w1, m.w2); });
            printf("{\"bench\": \"hbm_mix\", \"mix\": \"%s\", \"pattern\": \"box 64-byte rows per warp\", \"ctas_per_sm\": %d, \"gb_s\": %.0f}\n", m.name, bps, bytes / ms / 1e6);
            ms = timeit([&] { k_box<16><<<nsm * bps, 128>>>(x, c, a, B, Q, m.rd, m.w1, m.w2); });
            printf("{\"bench\": \"hbm_mix\", \"mix\": \"%s\", \"pattern\": \"box 256-byte rows per CTA\", \"ctas_per_sm\": %d, \"gb_s\": %.0f}\n", m.name, bps, bytes / ms / 1e6);
            ms = timeit([&] { k_box<4, true><<<nsm * bps, 128>>>(x, c, a, B, Q, m.rd, m.w1, m.w2); });
            printf("{\"bench\": \"hbm_mix\", \"mix\": \"%s\", \"pattern\": \"box 64-byte rows per warp, contiguous range per warp\", \"ctas_per_sm\": %d, \"gb_s\": %.0f}\n", m.name, bps, bytes / ms / 1e6);
            ms = timeit([&] { k_box<16, true><<<nsm * bps, 128>>>(x, c, a, B, Q, m.rd, m.w1, m.w2); });
            printf("{\"bench\": \"hbm_mix\", \"mix\": \"%s\", \"pattern\": \"box 256-byte rows per CTA, contiguous range per CTA\", \"ctas_per_sm\": %d, \"gb_s\": %.0f}\n", m.name, bps, bytes / ms / 1e6);
        }
    }
    return 0;
}

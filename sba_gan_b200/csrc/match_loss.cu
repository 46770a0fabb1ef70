// The B x B matching tail shared by words_loss and sent_loss:
//   * same-class masking + the two cross-entropies of a B x B score matrix
//     (AttnGAN2/code/miscc/losses.py:24-34 + 53-59 for sent_loss, :73-76 + 116-129 for words_loss): the reference builds
//     the mask in numpy on the host, copies it over, fills -inf in place and runs two CrossEntropyLoss modules;
//   * sent_loss's cosine score matrix scores[i, j] = gamma3 <a_i, b_j> / max(|a_i| |b_j|, eps) (losses.py:42-49)
//     and its gradient.
// Everything is B x B with B <= a few hundred: latency, not bandwidth - so few launches, nothing on the host,
// fixed summation order (no atomics).  Masking is applied on the fly from the class ids; the -inf matrix is
// never materialised.  Formulas: oracle/attention.py::ce_tail / sent_scores.
#include "kernels.h"

namespace sba {
namespace {

constexpr int kCeThreads = 256;

__device__ __forceinline__ bool masked(const int* cls, int i, int j) {
    return cls != nullptr && i != j && cls[i] == cls[j];
}

// block-wide (max, sum exp) in fixed order: warp shuffles, then warp 0 over the per-warp partials
__device__ __forceinline__ float block_lse(float m, float s, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mm = fmaxf(m, m2);
        s = (mm == -INFINITY) ? 0.f : s * expf(m - mm) + s2 * expf(m2 - mm);
        m = mm;
    }
    if (lane == 0) { red[2 * warp] = m; red[2 * warp + 1] = s; }
    __syncthreads();
    if (warp == 0) {
        m = lane < kCeThreads / 32 ? red[2 * lane] : -INFINITY;
        s = lane < kCeThreads / 32 ? red[2 * lane + 1] : 0.f;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
            const float mm = fmaxf(m, m2);
            s = (mm == -INFINITY) ? 0.f : s * expf(m - mm) + s2 * expf(m2 - mm);
            m = mm;
        }
        if (lane == 0) red[0] = m + logf(s);          // all entries -inf: -inf + log 0 = -inf + -inf = -inf
    }
    __syncthreads();
    const float r = red[0];
    __syncthreads();
    return r;
}

// block j: log-sum-exp of row j (lse[j]) and of column j (lse[B + j]) of the masked matrix, and the two picked
// entries s[j, lab_j], s[lab_j, j] (pick[j], pick[B + j])
__global__ void __launch_bounds__(kCeThreads) k_ce_lse(const float* __restrict__ s, const int* __restrict__ cls,
                                                       const long long* __restrict__ labels, float* __restrict__ lse,
                                                       float* __restrict__ pick, int B) {
    __shared__ float red[2 * (kCeThreads / 32)];
    const int j = blockIdx.x, tid = threadIdx.x;
    float m = -INFINITY, sum = 0.f;
    for (int i = tid; i < B; i += kCeThreads) {           // row j
        const float v = masked(cls, j, i) ? -INFINITY : s[(size_t)j * B + i];
        const float mm = fmaxf(m, v);
        sum = (mm == -INFINITY) ? 0.f : sum * expf(m - mm) + expf(v - mm);
        m = mm;
    }
    const float lr = block_lse(m, sum, red);
    m = -INFINITY; sum = 0.f;
    for (int i = tid; i < B; i += kCeThreads) {           // column j
        const float v = masked(cls, i, j) ? -INFINITY : s[(size_t)i * B + j];
        const float mm = fmaxf(m, v);
        sum = (mm == -INFINITY) ? 0.f : sum * expf(m - mm) + expf(v - mm);
        m = mm;
    }
    const float lc = block_lse(m, sum, red);
    if (tid == 0) {
        const int lab = (int)labels[j];
        lse[j] = lr;
        lse[B + j] = lc;
        pick[j] = masked(cls, j, lab) ? -INFINITY : s[(size_t)j * B + lab];
        pick[B + j] = masked(cls, lab, j) ? -INFINITY : s[(size_t)lab * B + j];
    }
}

// losses[0] = mean_j (lse_row[j] - s[j, lab_j]), losses[1] = mean_j (lse_col[j] - s[lab_j, j]); one block, fixed order
__global__ void __launch_bounds__(kCeThreads) k_ce_mean(const float* __restrict__ lse, const float* __restrict__ pick,
                                                        float* __restrict__ losses, int B) {
    __shared__ float red[2][kCeThreads];
    const int tid = threadIdx.x;
    float a0 = 0.f, a1 = 0.f;
    for (int j = tid; j < B; j += kCeThreads) {
        a0 += lse[j] - pick[j];
        a1 += lse[B + j] - pick[B + j];
    }
    red[0][tid] = a0;
    red[1][tid] = a1;
    __syncthreads();
    for (int o = kCeThreads / 2; o >= 1; o >>= 1) {
        if (tid < o) { red[0][tid] += red[0][tid + o]; red[1][tid] += red[1][tid + o]; }
        __syncthreads();
    }
    if (tid == 0) { losses[0] = red[0][0] / B; losses[1] = red[1][0] / B; }
}

// d_s[i, j] = g0/B (softmax_row_i[j] - [j == lab_i]) + g1/B (softmax_col_j[i] - [i == lab_j]); 0 where masked
__global__ void __launch_bounds__(kCeThreads) k_ce_bwd(const float* __restrict__ s, const int* __restrict__ cls,
                                                       const long long* __restrict__ labels, const float* __restrict__ lse,
                                                       const float* __restrict__ g, float* __restrict__ d_s, int B) {
    const int i = blockIdx.x;
    const float g0 = g[0] / B, g1 = g[1] / B;
    const int lab_i = (int)labels[i];
    const float lr = lse[i];
    for (int j = threadIdx.x; j < B; j += kCeThreads) {
        float d = 0.f;
        if (!masked(cls, i, j)) {
            const float v = s[(size_t)i * B + j];
            d = g0 * (expf(v - lr) - (j == lab_i ? 1.f : 0.f)) + g1 * (expf(v - lse[B + j]) - (i == (int)labels[j] ? 1.f : 0.f));
        }
        d_s[(size_t)i * B + j] = d;
    }
}

// ---- sent_loss scores -----------------------------------------------------------------------------
// norms[i] = |a_i|, norms[B + j] = |b_j| (one warp per row)
__global__ void __launch_bounds__(256) k_row_norms(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ norms, int B, int nef) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= 2 * B) return;
    const float* p = row < B ? a + (size_t)row * nef : b + (size_t)(row - B) * nef;
    float acc = 0.f;
    for (int c = lane; c < nef; c += 32) acc = fmaf(p[c], p[c], acc);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norms[row] = sqrtf(acc);
}

// scores[i, j] = g3 <a_i, b_j> / max(|a_i| |b_j|, eps): 32 x 32 output tile per block, K in chunks of 32
__global__ void __launch_bounds__(256) k_sent_scores(const float* __restrict__ a, const float* __restrict__ b,
                                                     const float* __restrict__ norms, float* __restrict__ scores, int B,
                                                     int nef, float g3, float eps) {
    __shared__ float as[32][33], bs[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 8 warps: 4 rows each
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < nef; c0 += 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = ty * 4 + r;
            as[row][tx] = (i0 + row < B && c0 + tx < nef) ? a[(size_t)(i0 + row) * nef + c0 + tx] : 0.f;
            bs[row][tx] = (j0 + row < B && c0 + tx < nef) ? b[(size_t)(j0 + row) * nef + c0 + tx] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
            const float bv = bs[tx][c];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(as[ty * 4 + r][c], bv, acc[r]);
        }
        __syncthreads();
    }
    const int j = j0 + tx;
    if (j < B) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + ty * 4 + r;
            if (i < B) scores[(size_t)i * B + j] = acc[r] / fmaxf(norms[i] * norms[B + j], eps) * g3;
        }
    }
}

// d_a[i, c] = sum_j coef_ij b[j, c] - a[i, c] t_i,  coef_ij = dS_ij g3 / max(|a_i||b_j|, eps),
// t_i = sum_{j: |a_i||b_j| > eps} dS_ij S_ij / |a_i|^2   (S = the unmasked scores);  side = 1: the same for b with
// the roles swapped (dS, S read transposed).  One block per row, thread = channel.
__global__ void __launch_bounds__(256) k_sent_bwd(const float* __restrict__ a, const float* __restrict__ b,
                                                  const float* __restrict__ norms, const float* __restrict__ S,
                                                  const float* __restrict__ dS, float* __restrict__ d_a,
                                                  float* __restrict__ d_b, int B, int nef, float g3, float eps) {
    extern __shared__ float coef[];                  // [B] + reduction scratch [8]
    float* red = coef + B;
    const int side = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
    const float* self = side == 0 ? a : b;
    const float* other = side == 0 ? b : a;
    float* out = side == 0 ? d_a : d_b;
    const float ni = norms[side * B + i];
    float t = 0.f;
    for (int j = tid; j < B; j += 256) {
        const size_t o = side == 0 ? (size_t)i * B + j : (size_t)j * B + i;
        const float nj = norms[(1 - side) * B + j];
        const float prod = ni * nj, ds = dS[o];
        coef[j] = ds * g3 / fmaxf(prod, eps);
        if (prod > eps && ds != 0.f) t += ds * S[o];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((tid & 31) == 0) red[tid >> 5] = t;
    __syncthreads();
    t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    t = ni > 0.f ? t / (ni * ni) : 0.f;
    for (int c = tid; c < nef; c += 256) {
        float acc = 0.f;
        for (int j = 0; j < B; ++j) acc = fmaf(coef[j], other[(size_t)j * nef + c], acc);
        out[(size_t)i * nef + c] = acc - self[(size_t)i * nef + c] * t;
    }
}

}  // namespace

int match_ce_fwd(const float* scores, const int* cls, const long long* labels, float* losses, float* lse, int B,
                 cudaStream_t st) {
    k_ce_lse<<<B, kCeThreads, 0, st>>>(scores, cls, labels, lse, lse + 2 * (size_t)B, B);
    k_ce_mean<<<1, kCeThreads, 0, st>>>(lse, lse + 2 * (size_t)B, losses, B);
    add_launches(2);
    return check_launch("match_ce_fwd");
}

int match_ce_bwd(const float* scores, const int* cls, const long long* labels, const float* lse, const float* g,
                 float* d_scores, int B, cudaStream_t st) {
    k_ce_bwd<<<B, kCeThreads, 0, st>>>(scores, cls, labels, lse, g, d_scores, B);
    add_launches(1);
    return check_launch("match_ce_bwd");
}

int sent_scores_fwd(const float* cnn, const float* rnn, float* scores, float* norms, int B, int nef, float g3, float eps,
                    cudaStream_t st) {
    k_row_norms<<<ceil_div(2 * B, 8), 256, 0, st>>>(cnn, rnn, norms, B, nef);
    k_sent_scores<<<dim3(ceil_div(B, 32), ceil_div(B, 32)), 256, 0, st>>>(cnn, rnn, norms, scores, B, nef, g3, eps);
    add_launches(2);
    return check_launch("sent_scores_fwd");
}

int sent_scores_bwd(const float* cnn, const float* rnn, const float* norms, const float* scores, const float* d_scores,
                    float* d_cnn, float* d_rnn, int B, int nef, float g3, float eps, cudaStream_t st) {
    const size_t smem = (size_t)(B + 8) * sizeof(float);
    if (smem > 48 * 1024) {
        set_error("sent_scores_bwd: B=%d exceeds the supported maximum of 12280", B);
        return SBA_ERR_UNSUPPORTED;
    }
    k_sent_bwd<<<dim3(B, 2), 256, smem, st>>>(cnn, rnn, norms, scores, d_scores, d_cnn, d_rnn, B, nef, g3, eps);
    add_launches(1);
    return check_launch("sent_scores_bwd");
}

}  // namespace sba

// Internal launch interface between abi.cu and the kernel translation units.
#pragma once
#include "common.cuh"

namespace sba {

struct AttnShape {
    int B, idf, cdf, L, Q;
    int dtype, mask_mode;
};

// attn_simt.cu - CUDA-core (FFMA + warp-shuffle) kernels, any shape with L <= 32
int simt_project(const float* ctx, const float* W, const uint8_t* mask, float* srcT, uint32_t* mask_bits,
                 const AttnShape& s, cudaStream_t st);
int simt_attn_fwd(const void* x, const float* srcT, const uint32_t* mask_bits, void* c_code, void* attn,
                  const AttnShape& s, cudaStream_t st);
int simt_attn_bwd(const void* x, const float* srcT, const uint32_t* mask_bits, const void* g_c,
                  const void* g_attn, void* dX, float* dSrc, const AttnShape& s, cudaStream_t st);
int simt_attn_bwd_epilogue(const float* ctx, const float* W, const float* dSrc, float* dW, float* dCtx,
                           const AttnShape& s, cudaStream_t st);

// attn_mma_fwd.cu / attn_mma_bwd.cu - tensor-core (mma.sync) family with TMA-staged tiles
bool mma_supports(const AttnShape& s);
int mma_attn_fwd(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                 float* srcT, uint32_t* mask_bits, const AttnShape& s, cudaStream_t st);
// dSrc holds B*idf*L floats followed by B+1 scratch words; dW / dCtx (nullable) are formed inside the kernel
int mma_attn_bwd(const void* x, const float* ctx, const float* W, const float* srcT, const uint8_t* mask, const void* g_c,
                 const void* g_attn, void* dX, float* dSrc, float* dW, float* dCtx, const AttnShape& s, cudaStream_t st);

// attn_tc5_fwd.cu / attn_tc5_bwd.cu - tcgen05 family: TMA tensor loads, UMMA with TMEM accumulators,
// one pixel per thread in the softmax / epilogue
bool tc5_supports(const AttnShape& s);
int tc5_attn_fwd(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                 float* srcT, uint32_t* mask_bits, const AttnShape& s, cudaStream_t st);

// dW (+)= sum_b dSrc[b] . ctx[b]^T (dW zeroed by the caller's launch sequence) and dCtx[b] = W^T . dSrc[b];
// launched as a programmatic dependent of the kernel that produced dSrc
// zero-fill grid in front of a backward kernel (which waits for it with griddepcontrol.wait before its first atomic)
int attn_bwd_zero(float* dSrc, size_t n_src, float* dW, size_t n_dw, cudaStream_t st);
int attn_bwd_post(const float* dSrc, const float* ctx, const float* W, float* dW, float* dCtx, int B, int idf, int cdf,
                  int L, cudaStream_t st);
bool tc5_bwd_supports(const AttnShape& s);     // bf16 tensors for now; fp32 backward stays on the mma.sync family
// dSrc holds B*idf*L floats followed by B+1 scratch words, like mma_attn_bwd
int tc5_attn_bwd(const void* x, const float* ctx, const float* W, const float* srcT, const uint8_t* mask,
                 const uint32_t* mask_bits, const void* g_c,
                 const void* g_attn, void* dX, float* dSrc, float* dW, float* dCtx, const AttnShape& s, cudaStream_t st);

// words_loss.cu - fused DAMSM region-word similarity (kernel c) and its backward
size_t words_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw);
int words_sim_fwd(const float* img, const float* words, const int* cap_lens, float* sim, float* att_diag, float* wc_out,
                  int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float g1, float g2, float g3, float eps,
                  int paired, cudaStream_t st);
int words_sim_bwd(const float* img, const float* words, const int* cap_lens, const float* d_sim, float* d_img,
                  float* d_words, void* workspace, int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float g1,
                  float g2, float g3, float eps, cudaStream_t st);

}  // namespace sba

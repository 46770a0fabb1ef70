// Internal launch interface between abi.cu and the kernel translation units.
#pragma once
#include "common.cuh"

namespace sba {

struct AttnShape {
    int B, idf, cdf, L, Q;
    int dtype, mask_mode;
    // forward output placement (tcgen05 family): c_code is written into rows [c_row0, c_row0 + idf) of a
    // [B, c_rows, Q] buffer; c_rows = 0 means a plain [B, idf, Q] tensor.  Backward: the same for g_c.
    int c_rows = 0, c_row0 = 0;
    // tcgen05 family: which of the call's two kernels to launch (SBA_PHASE_*): the first one (forward: projection;
    // backward: streaming kernel), the second one (forward: streaming kernel; backward: finish kernel) or both
    int phase = 0;
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

// attn_bwd_post.cu - helpers of the mma.sync backward: a zero-fill grid in front of it (it accumulates dSrc with
// atomics) and dW (+)= sum_b dSrc[b] . ctx[b]^T, dCtx[b] = W^T . dSrc[b] behind it, both programmatic dependents
int attn_bwd_zero(float* dSrc, size_t n_src, float* dW, size_t n_dw, cudaStream_t st);
int attn_bwd_post(const float* dSrc, const float* ctx, const float* W, float* dW, float* dCtx, int B, int idf, int cdf,
                  int L, cudaStream_t st);

// attn_tc5_bwd.cu - tcgen05 backward (bf16 tensors; fp32 backward stays on the mma.sync family): streaming kernel
// + finish kernel, no zero fill, no atomics.  `ws` is the caller's workspace of attn_bwd_workspace_floats() floats:
// [0, B*idf*L) receives dSrc, the rest holds per-(CTA, sample) partial sums and the finish kernel's scratch.
bool tc5_bwd_supports(const AttnShape& s);
size_t attn_bwd_workspace_floats(int B, int idf, int cdf, int L);
int tc5_attn_bwd(const void* x, const float* ctx, const float* W, const float* srcT, const uint8_t* mask,
                 uint32_t* mask_bits, const void* g_c, const void* g_attn, void* dX, float* ws, size_t ws_floats,
                 float* dW, float* dCtx, const AttnShape& s, cudaStream_t st);

// words_loss.cu - fused DAMSM region-word similarity (kernel c) and its backward
size_t words_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw);
int words_sim_fwd(const float* img, const float* words, const int* cap_lens, float* sim, float* att_diag, float* wc_out,
                  int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float g1, float g2, float g3, float eps,
                  int paired, cudaStream_t st);
int words_sim_bwd(const float* img, const float* words, const int* cap_lens, const float* d_sim, float* d_img,
                  float* d_words, void* workspace, int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float g1,
                  float g2, float g3, float eps, cudaStream_t st);

int words_att_diag(const float* img, const float* words, const int* cap_lens, float* att_diag, int B_img, int B_cap,
                   int row_offset, int nef, int R, int Lw, float g1, cudaStream_t st);

// words_tc5.cu - the forward of kernel (c) on tcgen05 (3xTF32 UMMA): sim only; the diagonal attention maps stay with
// words_att_diag.  `workspace`: words_tc5_workspace_bytes() bytes (0 = shape not covered), 256-byte aligned.
bool words_tc5_supports(int B_img, int B_cap, int nef, int R, int Lw);
size_t words_tc5_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw);
int words_sim_fwd_tc5(const float* img, const float* words, const int* cap_lens, float* sim, void* workspace, size_t ws_bytes,
                      int B_img, int B_cap, int nef, int R, int Lw, float g1, float g2, float g3, float eps, cudaStream_t st);

// backward of the same on the tensor cores: d_img, and d_words when asked for (NULL: words detached, GAN training); needs
// words_tc5_bwd_workspace_bytes() bytes (0 = not covered): phase A (forward again + per-column scalars, wc), phase B
// (u^T, a2^T) and the d_img GEMM, all tcgen05 3xTF32.
size_t words_tc5_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw, bool need_words);
int words_sim_bwd_tc5(const float* img, const float* words, const int* cap_lens, const float* d_sim, float* d_img,
                      float* d_words, void* workspace, size_t ws_bytes, int B_img, int B_cap, int nef, int R, int Lw, float g1,
                      float g2, float g3, float eps, cudaStream_t st);

// match_loss.cu - the B x B matching tail of words_loss / sent_loss: class masking + two-way cross-entropy, and
// sent_loss's cosine score matrix
// lse: [4*B] floats (row / column log-sum-exp, then the picked label entries), kept for the backward
int match_ce_fwd(const float* scores, const int* cls, const long long* labels, float* losses, float* lse, int B,
                 cudaStream_t st);
int match_ce_bwd(const float* scores, const int* cls, const long long* labels, const float* lse, const float* g,
                 float* d_scores, int B, cudaStream_t st);
int sent_scores_fwd(const float* cnn, const float* rnn, float* scores, float* norms, int B, int nef, float g3, float eps,
                    cudaStream_t st);
int sent_scores_bwd(const float* cnn, const float* rnn, const float* norms, const float* scores, const float* d_scores,
                    float* d_cnn, float* d_rnn, int B, int nef, float g3, float eps, cudaStream_t st);

// adain.cu - the caller's AdaIN written into the concatenated buffer next to c_code
// style [B, 2C] fp32 = (gamma | beta); stats [2 * B * C] fp32 = (mean, rstd) per row, kept for the backward
int adain_fwd(const void* x, const float* style, void* out, int out_rows, int out_row0, float* stats, int B, int C, int Q,
              int dtype, float eps, cudaStream_t st);
int adain_bwd(const void* x, const float* style, const float* stats, const void* g_buf, int g_rows, int g_row0, void* dX,
              int accumulate, float* d_style, int B, int C, int Q, int dtype, cudaStream_t st);

}  // namespace sba

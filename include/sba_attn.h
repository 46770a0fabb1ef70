/*
 * sba_attn.h - C ABI of libsba_attn.so: the B200 (sm_100a) word-region attention hot path
 * of zhengfei0908/SBA-GAN.
 *
 * The reference has no native boundary: its hot path is Python calling torch ops.  Each
 * entry point below names the reference interface it replaces (paths relative to
 * AttnGAN2/code in the reference tree).  A reference-side binding (ctypes) is shown in
 * INTEGRATION.md; the in-repo one is sba_gan_b200/_abi.py.
 *
 * Contract:
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller;
 *  - nothing here allocates, frees or synchronises; all work is enqueued on `stream`
 *    (a cudaStream_t passed as void*, e.g. torch.cuda.current_stream().cuda_stream);
 *  - returns SBA_OK (0) or a non-zero sba_status; the message is in sba_last_error()
 *    (thread-local); no C++ exception crosses this boundary;
 *  - stateless and re-entrant (one process per GPU under torchrun is fine);
 *  - big pixel tensors must be contiguous; 16-byte aligned bases enable the 128-bit path;
 *  - there is no CPU implementation behind these symbols.
 */
#ifndef SBA_ATTN_H_
#define SBA_ATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBA_ABI_VERSION 4

#if defined(__GNUC__)
#define SBA_API __attribute__((visibility("default")))
#else
#define SBA_API
#endif

typedef enum {
    SBA_OK = 0,
    SBA_ERR_ARG = 1,         /* null pointer / non-positive size */
    SBA_ERR_UNSUPPORTED = 2, /* shape outside what the kernels cover (e.g. L > 32) */
    SBA_ERR_ALIGN = 3,       /* pointer alignment */
    SBA_ERR_CUDA = 4         /* launch failure; message carries cudaGetErrorString */
} sba_status;

/* element type of the big per-pixel tensors (x, c_code, attn, g_c, g_attn, dX);
 * arithmetic is always fp32; small per-caption tensors are always fp32. */
enum { SBA_F32 = 0, SBA_BF16 = 1 };

/* how the B x L caption padding mask is applied to pixel (b, q):
 *  REFERENCE : caption (b*Q + q) mod B  - bug-compatible with mask.repeat(queryL, 1)
 *              at GlobalAttention.py:104-108; the default everywhere;
 *  PER_SAMPLE: caption b. */
enum { SBA_MASK_REFERENCE = 0, SBA_MASK_PER_SAMPLE = 1 };

/* kernel family: AUTO picks the fastest one that supports the shape
 *  SIMT    : CUDA-core FFMA + warp shuffles, any shape with L <= 32
 *  MMA     : warp-level mma.sync with TMA-staged tiles
 *  TCGEN05 : tcgen05.mma with TMEM accumulators, TMA tensor loads / stores, one pixel per thread */
enum { SBA_ALGO_AUTO = 0, SBA_ALGO_SIMT = 1, SBA_ALGO_MMA = 2, SBA_ALGO_TCGEN05 = 3 };
/* which of a call's two kernels to launch (sba_attn_fwd_phase / sba_attn_bwd_phase) */
enum { SBA_PHASE_ALL = 0, SBA_PHASE_FIRST = 1, SBA_PHASE_SECOND = 2 };

SBA_API int sba_abi_version(void);
SBA_API const char* sba_last_error(void);

/* Number of kernels the last successful call on this thread launched (bench.py's
 * gpu_launches). */
SBA_API int sba_last_launch_count(void);

/* ---- GlobalAttentionGeneral.forward  (GlobalAttention.py:82-121) -------------------
 * x        [B, idf, Q]   dtype      input regions (NCHW, Q = ih*iw)
 * ctx      [B, cdf, L]   fp32       word features
 * W        [idf, cdf]    fp32       conv_context.weight ([idf,cdf,1,1], GlobalAttention.py:25-28,75)
 * mask     [B, L] uint8  nullable   1 = padding word (applyMask, GlobalAttention.py:79-80)
 * c_code   [B, idf, Q]   dtype  out weightedContext
 * attn     [B, L, Q]     dtype  out attention map
 * srcT     [B, idf, L]   fp32   out sourceT = W.ctx, kept for the backward
 * scratch  [3*B] uint32  scratch    [0,B): caption mask bit words (kept for the backward),
 *                                   [B,3B): free for the kernels of this launch (word B is the
 *                                   tile-schedule counter of the tcgen05 forward)
 */
SBA_API int sba_attn_fwd(const void* x, const float* ctx, const float* W, const uint8_t* mask,
                 void* c_code, void* attn, float* srcT, uint32_t* scratch,
                 int B, int idf, int cdf, int L, int Q,
                 int dtype, int mask_mode, int algo, void* stream);

/* Kernel family that served the last successful sba_attn_fwd* / sba_attn_bwd* call on this thread
 * (an SBA_ALGO_* value), and whether a concrete family covers a shape (which: 0 = forward,
 * 1 = backward).  An explicit `algo` is honoured strictly: a family that does not cover the shape
 * is refused with SBA_ERR_UNSUPPORTED, never silently replaced. */
SBA_API int sba_last_algo(void);
SBA_API int sba_attn_supported(int which, int algo, int B, int idf, int cdf, int L, int Q, int dtype);

/* ---- autograd backward of the above ------------------------------
 * srcT     [B, idf, L] fp32         from the forward
 * scratch  [3*B] uint32             from the forward (mask words)
 * g_c      [B, idf, Q] dtype        grad of c_code
 * g_attn   [B, L, Q]   dtype nullable grad of attn (NULL in GAN training: attn is discarded,
 *                                   trainer_bert.py:267)
 * dX       [B, idf, Q] dtype  out
 * ws       [ws_floats] fp32         workspace of at least sba_attn_bwd_workspace_floats(B, idf, cdf, L)
 *                                   floats, 16-byte aligned, contents undefined on entry (nothing needs
 *                                   zeroing).  When dW or dCtx is requested its first B*idf*L floats
 *                                   hold the gradient of sourceT [B, idf, L] on return.
 * dW       [idf, cdf]  fp32   out   nullable
 * dCtx     [B, cdf, L] fp32   out   nullable (words are detached in GAN training)
 * The tcgen05 family accumulates in a fixed order: results are bit-identical run to run.
 */
SBA_API size_t sba_attn_bwd_workspace_floats(int B, int idf, int cdf, int L);
SBA_API int sba_attn_bwd(const void* x, const float* ctx, const float* W, const uint8_t* mask,
                 const float* srcT, uint32_t* scratch,
                 const void* g_c, const void* g_attn,
                 void* dX, float* ws, size_t ws_floats, float* dW, float* dCtx,
                 int B, int idf, int cdf, int L, int Q,
                 int dtype, int mask_mode, int algo, void* stream);

/* ---- the two kernels of a call, separately (tcgen05 family only) ----------------------
 * Both calls consist of a streaming kernel over the pixels and a small kernel that does not touch them:
 *   forward : FIRST  = the projection sourceT = conv_context(context) (+ mask words): needs only ctx, W, mask and
 *                      writes srcT / scratch - it can run as soon as the word features exist (in G_NET.forward,
 *                      model_bert.py:580-588, that is before the first stage has produced h_code), on any stream;
 *             SECOND = the streaming kernel: needs x, mask, srcT / scratch of a completed FIRST, writes c_code, attn.
 *   backward: FIRST  = the streaming kernel: writes dX and the partial sums in ws;
 *             SECOND = the finish kernel: ws -> dW, dCtx (parameter / word-feature gradients: nothing in the backward
 *                      chain waits for them, so it can run on a side stream under the next layer's backward).
 * Ordering between the phases is the caller's (same stream, or an event).  A projection serves exactly ONE streaming
 * call: scratch also carries that call's dynamic tile counter.  SBA_PHASE_ALL = the plain calls above.
 * Pointers a phase does not use may be NULL.  Results are bit-identical to the one-call form.
 */
SBA_API int sba_attn_fwd_phase(const void* x, const float* ctx, const float* W, const uint8_t* mask,
                 void* c_code, void* attn, float* srcT, uint32_t* scratch,
                 int B, int idf, int cdf, int L, int Q,
                 int dtype, int mask_mode, int phase, void* stream);
SBA_API int sba_attn_bwd_phase(const void* x, const float* ctx, const float* W, const uint8_t* mask,
                 const float* srcT, uint32_t* scratch,
                 const void* g_c, const void* g_attn,
                 void* dX, float* ws, size_t ws_floats, float* dW, float* dCtx,
                 int B, int idf, int cdf, int L, int Q,
                 int dtype, int mask_mode, int phase, void* stream);

/* ---- the caller's torch.cat((h_code, c_code), 1) folded into the attention ----------
 * (NEXT_STAGE_G.forward, model_bert.py:459-461)
 * sba_attn_fwd_into writes weightedContext into rows [c_row0, c_row0 + idf) of every sample of
 * c_buf [B, c_rows, Q] (e.g. c_rows = 2*idf, c_row0 = idf: the second half of h_c_code) instead
 * of a tensor of its own; sba_attn_bwd_from reads g_c from the same rows of g_buf [B, g_rows, Q]
 * (the gradient of the concatenated buffer), so neither the concatenation nor its backward
 * slice copy exists.  tcgen05 family only (shapes: sba_attn_supported(.., SBA_ALGO_TCGEN05, ..)).
 */
SBA_API int sba_attn_fwd_into(const void* x, const float* ctx, const float* W, const uint8_t* mask,
                 void* c_buf, int c_rows, int c_row0, void* attn, float* srcT, uint32_t* scratch,
                 int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode, void* stream);
SBA_API int sba_attn_bwd_from(const void* x, const float* ctx, const float* W, const uint8_t* mask,
                 const float* srcT, uint32_t* scratch,
                 const void* g_buf, int g_rows, int g_row0, const void* g_attn,
                 void* dX, float* ws, size_t ws_floats, float* dW, float* dCtx,
                 int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode, void* stream);

/* ---- DAMSM region-word similarity: func_attention + cosine + LSE --------------------
 * (GlobalAttention.py:31-69, miscc/losses.py:11-17, 72-123)
 * sim[j, i] = g3 * log sum_{t < len_i} exp(g2 * cos(words[i,:,t], wc_{j,i}[:,t]))
 * for local image rows j in [0, B_img) against all B_cap captions (rows = images,
 * cols = captions, like `similarities` at losses.py:115).
 * img      [B_img, nef, R] fp32     region features (R = ih*iw = 17*17 <= 320)
 * words    [B_cap, nef, Lw] fp32    word embeddings, Lw = padded width (<= 32), nef <= 256
 * cap_lens [B_cap] int32            true lengths, 1 <= len <= Lw
 * sim      [B_img, B_cap] fp32 out
 * att_diag [B_cap, Lw, R] fp32 out nullable: row i receives the region attention of pair
 *          (local image j = i - row_offset, caption i) when that image is local - the
 *          att_maps of losses.py:92; other rows are left untouched
 */
SBA_API int sba_words_sim_fwd(const float* img, const float* words, const int32_t* cap_lens,
                      float* sim, float* att_diag,
                      int B_img, int B_cap, int row_offset, int nef, int R, int Lw,
                      float gamma1, float gamma2, float gamma3, float eps, void* stream);

/* The same on the tensor cores (tcgen05, 3xTF32: every product as hi.hi + hi.lo + lo.hi, fp32 parity kept): both
 * contractions of func_attention as chained UMMA GEMMs per (image, 128 packed word columns), softmaxes on chip.
 * sba_words_sim_fwd_workspace_bytes() is 0 when the shape is not covered (nef % 32, nef <= 256, R <= 384); the
 * workspace (256-byte aligned, contents undefined on entry) holds the tf32-split, TMA-addressable operand copies.
 * Same results as sba_words_sim_fwd (att_diag comes from the CUDA-core kernel, diagonal pairs only). */
SBA_API size_t sba_words_sim_fwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw);
SBA_API int sba_words_sim_fwd_ws(const float* img, const float* words, const int32_t* cap_lens,
                      float* sim, float* att_diag, void* workspace, size_t workspace_bytes,
                      int B_img, int B_cap, int row_offset, int nef, int R, int Lw,
                      float gamma1, float gamma2, float gamma3, float eps, void* stream);

/* Gradient of sim w.r.t. img (always) and words (d_words nullable: only DAMSM pre-training
 * needs it, pretrain_DAMSM_bert.py:86-92).  d_img [B_img, nef, R] and d_words
 * [B_cap, nef, Lw] are overwritten.  `workspace` holds per-pair intermediates
 * (sba_words_sim_bwd_workspace_bytes, 256-byte aligned). */
SBA_API size_t sba_words_sim_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw);
SBA_API int sba_words_sim_bwd(const float* img, const float* words, const int32_t* cap_lens,
                      const float* d_sim, float* d_img, float* d_words, void* workspace,
                      int B_img, int B_cap, int row_offset, int nef, int R, int Lw,
                      float gamma1, float gamma2, float gamma3, float eps, void* stream);

/* ---- func_attention (GlobalAttention.py:31-69) as a standalone operator -------------
 * query [B, nef, T], context [B, nef, R]  ->  wc [B, nef, T], attn [B, T, R]
 * (batch element b of the query attends over batch element b of the context). */
SBA_API int sba_func_attention(const float* query, const float* context, float* wc, float* attn,
                       int B, int nef, int T, int R, float gamma1, void* stream);

/* ---- the caller's AdaIN, written next to c_code ----------------------
 * ADAIN_NORM.forward (model_bert.py:367-374) of NEXT_STAGE_G.forward (:458-461):
 *   out[b, out_row0 + c, :] = (gamma[b,c] + 1) * instance_norm(x[b, c, :]) + beta[b,c]
 * with InstanceNorm2d semantics (biased variance over the Q pixels of the row, eps, no affine,
 * no running statistics).  out is a [B, out_rows, Q] buffer - with out_rows = 2*C, out_row0 = 0
 * and sba_attn_fwd_into writing c_code into rows [C, 2C) the reference's
 * torch.cat((h_code, c_code), 1) is complete without a copy.  One read + one write of x.
 * x       [B, C, Q] dtype            style [B, 2C] fp32 = (gamma | beta) (the output of ADAIN_NORM.style)
 * stats   [2*B*C] fp32 out: (mean, rstd) per row, kept for the backward
 * Backward: g_buf [B, g_rows, Q] is the gradient of the buffer, rows [g_row0, g_row0 + C) are read in
 * place; dX [B, C, Q] receives (accumulate = 0) or is incremented by (accumulate = 1: on top of the
 * attention's dX) the gradient w.r.t. x; d_style [B, 2C] = (d gamma | d beta).
 */
SBA_API int sba_adain_fwd(const void* x, const float* style, void* out, int out_rows, int out_row0,
                     float* stats, int B, int C, int Q, int dtype, float eps, void* stream);
SBA_API int sba_adain_bwd(const void* x, const float* style, const float* stats,
                     const void* g_buf, int g_rows, int g_row0, void* dX, int accumulate, float* d_style,
                     int B, int C, int Q, int dtype, void* stream);

/* ---- the B x B matching tail of words_loss and sent_loss -----------
 * sba_match_ce_*: same-class masking + the two cross-entropies (miscc/losses.py:24-34, 53-59 and
 * :73-76, 116-129).  Entry (i, j), i != j, counts as -inf when class_ids[i] == class_ids[j]
 * (class_ids NULL: no masking); losses[0] = CE(scores, labels), losses[1] = CE(scores^T, labels),
 * mean over B.  The mask is applied on the fly - nothing is built on the host, the -inf matrix
 * is never written.
 * scores    [B, B] fp32 (unmasked)      class_ids [B] int32 nullable      labels [B] int64
 * losses    [2] fp32 out                lse [4*B] fp32 out: kept for the backward
 * g         [2] fp32: upstream gradients of the two losses          d_scores [B, B] fp32 out
 */
SBA_API int sba_match_ce_fwd(const float* scores, const int32_t* class_ids, const int64_t* labels,
                     float* losses, float* lse, int B, void* stream);
SBA_API int sba_match_ce_bwd(const float* scores, const int32_t* class_ids, const int64_t* labels,
                     const float* lse, const float* g, float* d_scores, int B, void* stream);

/* sba_sent_scores_*: the score matrix of sent_loss (miscc/losses.py:42-49):
 * scores[i, j] = gamma3 * <cnn_i, rnn_j> / max(|cnn_i| |rnn_j|, eps), and its gradient.
 * cnn, rnn  [B, nef] fp32      scores [B, B] fp32 out      norms [2*B] fp32 out (kept for the backward)
 */
SBA_API int sba_sent_scores_fwd(const float* cnn, const float* rnn, float* scores, float* norms,
                     int B, int nef, float gamma3, float eps, void* stream);
SBA_API int sba_sent_scores_bwd(const float* cnn, const float* rnn, const float* norms, const float* scores,
                     const float* d_scores, float* d_cnn, float* d_rnn,
                     int B, int nef, float gamma3, float eps, void* stream);

/* Backward of sba_words_sim_fwd_ws on the tensor cores (tcgen05, 3xTF32): the forward again with the upstream gradient
 * (per-column scalars, wc), then S = X^T W and V = X^T wc as chained UMMA GEMMs with both softmax backwards on chip, then
 * d_img as one GEMM over the word columns and - when d_words is not NULL - d_words as one split-K GEMM over (image,
 * region) whose partials are added in a fixed order.  d_words == NULL is the GAN-training call (trainer_bert.py:257
 * detaches the words); non-NULL is DAMSM pre-training (pretrain_DAMSM.py:80-91).  The workspace query takes the same
 * choice (need_words) and returns 0 when the shape is not covered.  Same results as sba_words_sim_bwd. */
SBA_API size_t sba_words_sim_bwd_tc_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw, int need_words);
SBA_API int sba_words_sim_bwd_tc(const float* img, const float* words, const int32_t* cap_lens,
                      const float* d_sim, float* d_img, float* d_words, void* workspace, size_t workspace_bytes,
                      int B_img, int B_cap, int row_offset, int nef, int R, int Lw,
                      float gamma1, float gamma2, float gamma3, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SBA_ATTN_H_ */

// extern "C" boundary of libsba_attn.so (declared in include/sba_attn.h).
// Validates arguments, picks a kernel family, enqueues on the caller's stream.
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace sba {

namespace {
thread_local char g_err[512] = "";
thread_local int g_launches = 0;
thread_local int g_algo = SBA_ALGO_AUTO;      // kernel family of the last attention call
}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void add_launches(int n) { g_launches += n; }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    return SBA_OK;
}

namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_attn_shape(const char* fn, int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode, int algo) {
    if (B <= 0 || idf <= 0 || cdf <= 0 || L <= 0 || Q <= 0) {
        set_error("%s: sizes must be positive (B=%d idf=%d cdf=%d L=%d Q=%d)", fn, B, idf, cdf, L, Q);
        return SBA_ERR_ARG;
    }
    if (L > kMaxWords) {
        set_error("%s: L=%d words exceeds the supported maximum of %d", fn, L, kMaxWords);
        return SBA_ERR_UNSUPPORTED;
    }
    // the CUDA-core family keeps sourceT (and, backward, a 32-pixel staging tile per channel) in shared memory:
    // idf * 64 floats <= 64 KB; the word features of one sample (cdf * L floats) must fit as well
    if (idf > 256) {
        set_error("%s: idf=%d exceeds the supported maximum of 256", fn, idf);
        return SBA_ERR_UNSUPPORTED;
    }
    if ((size_t)cdf * L * sizeof(float) > 200 * 1024) {
        set_error("%s: cdf*L = %d context words per sample exceed the supported maximum of 51200", fn, cdf * L);
        return SBA_ERR_UNSUPPORTED;
    }
    if (dtype != SBA_F32 && dtype != SBA_BF16) {
        set_error("%s: unknown dtype %d", fn, dtype);
        return SBA_ERR_ARG;
    }
    if (mask_mode != SBA_MASK_REFERENCE && mask_mode != SBA_MASK_PER_SAMPLE) {
        set_error("%s: unknown mask_mode %d", fn, mask_mode);
        return SBA_ERR_ARG;
    }
    if (algo < SBA_ALGO_AUTO || algo > SBA_ALGO_TCGEN05) {
        set_error("%s: unknown algo %d", fn, algo);
        return SBA_ERR_ARG;
    }
    if ((size_t)B * idf * Q >= ((size_t)1 << 40)) {
        set_error("%s: tensor too large", fn);
        return SBA_ERR_UNSUPPORTED;
    }
    return SBA_OK;
}

}  // namespace
}  // namespace sba

using namespace sba;

extern "C" {

int sba_abi_version(void) { return SBA_ABI_VERSION; }
const char* sba_last_error(void) { return g_err; }
int sba_last_launch_count(void) { return g_launches; }

int sba_last_algo(void) { return g_algo; }

// which == 0: forward, 1: backward.  1 when `algo` (a concrete family) covers the shape, else 0.
int sba_attn_supported(int which, int algo, int B, int idf, int cdf, int L, int Q, int dtype) {
    if (B <= 0 || idf <= 0 || idf > 256 || cdf <= 0 || L <= 0 || L > kMaxWords || Q <= 0) return 0;
    if (dtype != SBA_F32 && dtype != SBA_BF16) return 0;
    AttnShape s{B, idf, cdf, L, Q, dtype, SBA_MASK_REFERENCE};
    switch (algo) {
        case SBA_ALGO_SIMT: return 1;
        case SBA_ALGO_MMA: return mma_supports(s) ? 1 : 0;
        case SBA_ALGO_TCGEN05: return (which == 0 ? tc5_supports(s) : tc5_bwd_supports(s)) ? 1 : 0;
        case SBA_ALGO_AUTO: return 1;
        default: return 0;
    }
}

size_t sba_attn_bwd_workspace_floats(int B, int idf, int cdf, int L) {
    if (B <= 0 || idf <= 0 || cdf <= 0 || L <= 0) return 0;
    return attn_bwd_workspace_floats(B, idf, cdf, L);
}

static int attn_fwd_impl(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                         float* srcT, uint32_t* mask_bits, AttnShape s, int algo, void* stream, const char* fn) {
    g_launches = 0;
    g_err[0] = 0;
    const bool first = s.phase != SBA_PHASE_SECOND, second = s.phase != SBA_PHASE_FIRST;
    if (s.phase < SBA_PHASE_ALL || s.phase > SBA_PHASE_SECOND) {
        set_error("%s: phase %d is not one of SBA_PHASE_*", fn, s.phase);
        return SBA_ERR_ARG;
    }
    if (!srcT || !mask_bits || (first && (!ctx || !W)) || (second && (!x || !c_code || !attn))) {
        set_error("%s: null pointer argument", fn);
        return SBA_ERR_ARG;
    }
    int rc = check_attn_shape(fn, s.B, s.idf, s.cdf, s.L, s.Q, s.dtype, s.mask_mode, algo);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool strided = s.c_rows > 0;          // c_code goes into a wider buffer: tensor-map stores only
    const bool can_mma = !strided && mma_supports(s) && aligned16(x);
    const bool can_tc5 = tc5_supports(s) && (!first || aligned16(W)) &&
                         (!second || (aligned16(x) && aligned16(c_code) && aligned16(attn)));
    if (algo == SBA_ALGO_TCGEN05 || (algo == SBA_ALGO_AUTO && can_tc5) || strided) {
        if (!can_tc5) {
            set_error("%s: SBA_ALGO_TCGEN05 does not cover idf=%d L=%d Q=%d B=%d (or x is not 16-byte aligned)", fn, s.idf,
                      s.L, s.Q, s.B);
            return SBA_ERR_UNSUPPORTED;
        }
        g_algo = SBA_ALGO_TCGEN05;
        return tc5_attn_fwd(x, ctx, W, mask, c_code, attn, srcT, mask_bits, s, st);
    }
    if (algo == SBA_ALGO_MMA && !can_mma) {
        set_error("%s: SBA_ALGO_MMA does not cover idf=%d L=%d Q=%d cdf=%d B=%d (or x is not 16-byte aligned)", fn, s.idf,
                  s.L, s.Q, s.cdf, s.B);
        return SBA_ERR_UNSUPPORTED;
    }
    if (algo != SBA_ALGO_SIMT && can_mma) {
        g_algo = SBA_ALGO_MMA;
        return mma_attn_fwd(x, ctx, W, mask, c_code, attn, srcT, mask_bits, s, st);
    }
    g_algo = SBA_ALGO_SIMT;
    rc = simt_project(ctx, W, mask, srcT, mask_bits, s, st);
    if (rc) return rc;
    return simt_attn_fwd(x, srcT, mask ? mask_bits : nullptr, c_code, attn, s, st);
}

static int attn_bwd_impl(const void* x, const float* ctx, const float* W, const uint8_t* mask, const float* srcT,
                         uint32_t* mask_bits, const void* g_c, const void* g_attn, void* dX, float* ws,
                         size_t ws_floats, float* dW, float* dCtx, AttnShape s, int algo, void* stream, const char* fn) {
    g_launches = 0;
    g_err[0] = 0;
    const bool first = s.phase != SBA_PHASE_SECOND, second = s.phase != SBA_PHASE_FIRST;
    if (s.phase < SBA_PHASE_ALL || s.phase > SBA_PHASE_SECOND) {
        set_error("%s: phase %d is not one of SBA_PHASE_*", fn, s.phase);
        return SBA_ERR_ARG;
    }
    if (!ws || (first && (!x || !srcT || !g_c || !dX || !mask_bits)) || (second && (!ctx || !W))) {
        set_error("%s: null pointer argument", fn);
        return SBA_ERR_ARG;
    }
    int rc = check_attn_shape(fn, s.B, s.idf, s.cdf, s.L, s.Q, s.dtype, s.mask_mode, algo);
    if (rc) return rc;
    if (ws_floats < (size_t)s.B * s.idf * s.L + s.B + 1) {
        set_error("%s: workspace of %zu floats is too small (sba_attn_bwd_workspace_floats)", fn, ws_floats);
        return SBA_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool strided = s.c_rows > 0;          // g_c is a slice of a wider buffer: tensor-map loads only
    const bool can_mma = !strided && mma_supports(s) && aligned16(x) && aligned16(g_c);
    const bool can_tc5 = tc5_bwd_supports(s) && aligned16(ws) && (!first || (aligned16(x) && aligned16(g_c) && aligned16(dX)));
    if (algo == SBA_ALGO_TCGEN05 || (algo == SBA_ALGO_AUTO && can_tc5) || strided) {
        if (!can_tc5) {
            set_error("%s: SBA_ALGO_TCGEN05 does not cover idf=%d L=%d Q=%d B=%d dtype=%d (bf16 tensors, 16-byte aligned)", fn,
                      s.idf, s.L, s.Q, s.B, s.dtype);
            return SBA_ERR_UNSUPPORTED;
        }
        g_algo = SBA_ALGO_TCGEN05;
        return tc5_attn_bwd(x, ctx, W, srcT, mask, mask_bits, g_c, g_attn, dX, ws, ws_floats, dW, dCtx, s, st);
    }
    if (algo == SBA_ALGO_MMA && !can_mma) {
        set_error("%s: SBA_ALGO_MMA does not cover idf=%d L=%d Q=%d cdf=%d B=%d (or x / g_c is not 16-byte aligned)", fn,
                  s.idf, s.L, s.Q, s.cdf, s.B);
        return SBA_ERR_UNSUPPORTED;
    }
    if (algo != SBA_ALGO_SIMT && can_mma) {
        g_algo = SBA_ALGO_MMA;
        return mma_attn_bwd(x, ctx, W, srcT, mask, g_c, g_attn, dX, ws, dW, dCtx, s, st);
    }
    g_algo = SBA_ALGO_SIMT;
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)s.B * s.idf * s.L * sizeof(float), st);
    if (e != cudaSuccess) {
        set_error("%s: memset: %s", fn, cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    rc = simt_attn_bwd(x, srcT, mask ? mask_bits : nullptr, g_c, g_attn, dX, ws, s, st);
    if (rc) return rc;
    return simt_attn_bwd_epilogue(ctx, W, ws, dW, dCtx, s, st);
}

int sba_attn_fwd(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                 float* srcT, uint32_t* mask_bits, int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode,
                 int algo, void* stream) {
    return attn_fwd_impl(x, ctx, W, mask, c_code, attn, srcT, mask_bits, AttnShape{B, idf, cdf, L, Q, dtype, mask_mode}, algo,
                         stream, "sba_attn_fwd");
}

int sba_attn_bwd(const void* x, const float* ctx, const float* W, const uint8_t* mask, const float* srcT,
                 uint32_t* mask_bits, const void* g_c, const void* g_attn, void* dX, float* ws, size_t ws_floats,
                 float* dW, float* dCtx, int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode, int algo,
                 void* stream) {
    return attn_bwd_impl(x, ctx, W, mask, srcT, mask_bits, g_c, g_attn, dX, ws, ws_floats, dW, dCtx,
                         AttnShape{B, idf, cdf, L, Q, dtype, mask_mode}, algo, stream, "sba_attn_bwd");
}

// One kernel of a call at a time (tcgen05 family): the projection / finish kernels do not touch the pixel tensors and can
// be scheduled off the critical path by the caller (include/sba_attn.h).
int sba_attn_fwd_phase(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                       float* srcT, uint32_t* mask_bits, int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode,
                       int phase, void* stream) {
    AttnShape s{B, idf, cdf, L, Q, dtype, mask_mode};
    s.phase = phase;
    return attn_fwd_impl(x, ctx, W, mask, c_code, attn, srcT, mask_bits, s, SBA_ALGO_TCGEN05, stream, "sba_attn_fwd_phase");
}

int sba_attn_bwd_phase(const void* x, const float* ctx, const float* W, const uint8_t* mask, const float* srcT,
                       uint32_t* mask_bits, const void* g_c, const void* g_attn, void* dX, float* ws, size_t ws_floats,
                       float* dW, float* dCtx, int B, int idf, int cdf, int L, int Q, int dtype, int mask_mode, int phase,
                       void* stream) {
    AttnShape s{B, idf, cdf, L, Q, dtype, mask_mode};
    s.phase = phase;
    return attn_bwd_impl(x, ctx, W, mask, srcT, mask_bits, g_c, g_attn, dX, ws, ws_floats, dW, dCtx, s, SBA_ALGO_TCGEN05, stream,
                         "sba_attn_bwd_phase");
}

// NEXT_STAGE_G's torch.cat((h_code, c_code), 1) folded into the attention (model_bert.py:460-461): the forward
// writes c_code into rows [c_row0, c_row0 + idf) of every sample of a [B, c_rows, Q] buffer, the backward reads
// g_c from the same rows of the buffer's gradient.  tcgen05 family only (tensor-map addressing).
int sba_attn_fwd_into(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_buf, int c_rows,
                      int c_row0, void* attn, float* srcT, uint32_t* mask_bits, int B, int idf, int cdf, int L, int Q,
                      int dtype, int mask_mode, void* stream) {
    if (c_rows < idf || c_row0 < 0 || c_row0 + idf > c_rows) {
        set_error("sba_attn_fwd_into: rows [%d, %d) do not fit a %d-row buffer", c_row0, c_row0 + idf, c_rows);
        return SBA_ERR_ARG;
    }
    AttnShape s{B, idf, cdf, L, Q, dtype, mask_mode};
    s.c_rows = c_rows;
    s.c_row0 = c_row0;
    return attn_fwd_impl(x, ctx, W, mask, c_buf, attn, srcT, mask_bits, s, SBA_ALGO_TCGEN05, stream, "sba_attn_fwd_into");
}

int sba_attn_bwd_from(const void* x, const float* ctx, const float* W, const uint8_t* mask, const float* srcT,
                      uint32_t* mask_bits, const void* g_buf, int g_rows, int g_row0, const void* g_attn, void* dX,
                      float* ws, size_t ws_floats, float* dW, float* dCtx, int B, int idf, int cdf, int L, int Q, int dtype,
                      int mask_mode, void* stream) {
    if (g_rows < idf || g_row0 < 0 || g_row0 + idf > g_rows) {
        set_error("sba_attn_bwd_from: rows [%d, %d) do not fit a %d-row buffer", g_row0, g_row0 + idf, g_rows);
        return SBA_ERR_ARG;
    }
    AttnShape s{B, idf, cdf, L, Q, dtype, mask_mode};
    s.c_rows = g_rows;
    s.c_row0 = g_row0;
    return attn_bwd_impl(x, ctx, W, mask, srcT, mask_bits, g_buf, g_attn, dX, ws, ws_floats, dW, dCtx, s, SBA_ALGO_TCGEN05,
                         stream, "sba_attn_bwd_from");
}

static int check_words_shape(const char* fn, int B_img, int B_cap, int nef, int R, int Lw) {
    if (B_img <= 0 || B_cap <= 0 || nef <= 0 || R <= 0 || Lw <= 0) {
        set_error("%s: sizes must be positive (B_img=%d B_cap=%d nef=%d R=%d Lw=%d)", fn, B_img, B_cap, nef, R, Lw);
        return SBA_ERR_ARG;
    }
    if (Lw > kMaxWords) {
        set_error("%s: Lw=%d words exceeds the supported maximum of %d", fn, Lw, kMaxWords);
        return SBA_ERR_UNSUPPORTED;
    }
    return SBA_OK;
}

int sba_words_sim_fwd(const float* img, const float* words, const int32_t* cap_lens, float* sim, float* att_diag,
                      int B_img, int B_cap, int row_offset, int nef, int R, int Lw, float gamma1, float gamma2,
                      float gamma3, float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!img || !words || !cap_lens || !sim) {
        set_error("sba_words_sim_fwd: null pointer argument");
        return SBA_ERR_ARG;
    }
    int rc = check_words_shape("sba_words_sim_fwd", B_img, B_cap, nef, R, Lw);
    if (rc) return rc;
    return words_sim_fwd(img, words, cap_lens, sim, att_diag, nullptr, B_img, B_cap, row_offset, nef, R, Lw, gamma1,
                         gamma2, gamma3, eps, 0, static_cast<cudaStream_t>(stream));
}

size_t sba_words_sim_fwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw) {
    if (B_img <= 0 || B_cap <= 0 || nef <= 0 || R <= 0 || Lw <= 0) return 0;
    return words_tc5_workspace_bytes(B_img, B_cap, nef, R, Lw);
}

int sba_words_sim_fwd_ws(const float* img, const float* words, const int32_t* cap_lens, float* sim, float* att_diag,
                         void* workspace, size_t workspace_bytes, int B_img, int B_cap, int row_offset, int nef, int R, int Lw,
                         float gamma1, float gamma2, float gamma3, float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!img || !words || !cap_lens || !sim || !workspace) {
        set_error("sba_words_sim_fwd_ws: null pointer argument");
        return SBA_ERR_ARG;
    }
    int rc = check_words_shape("sba_words_sim_fwd_ws", B_img, B_cap, nef, R, Lw);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = words_sim_fwd_tc5(img, words, cap_lens, sim, workspace, workspace_bytes, B_img, B_cap, nef, R, Lw, gamma1, gamma2, gamma3,
                           eps, st);
    if (rc || att_diag == nullptr) return rc;
    return words_att_diag(img, words, cap_lens, att_diag, B_img, B_cap, row_offset, nef, R, Lw, gamma1, st);
}

size_t sba_words_sim_bwd_tc_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw, int need_words) {
    if (B_img <= 0 || B_cap <= 0 || nef <= 0 || R <= 0 || Lw <= 0) return 0;
    return words_tc5_bwd_workspace_bytes(B_img, B_cap, nef, R, Lw, need_words != 0);
}

int sba_words_sim_bwd_tc(const float* img, const float* words, const int32_t* cap_lens, const float* d_sim, float* d_img,
                         float* d_words, void* workspace, size_t workspace_bytes, int B_img, int B_cap, int row_offset, int nef,
                         int R, int Lw, float gamma1, float gamma2, float gamma3, float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    (void)row_offset;        // the similarity of a pair does not depend on where the image rows sit in the global batch
    if (!img || !words || !cap_lens || !d_sim || !d_img || !workspace) {
        set_error("sba_words_sim_bwd_tc: null pointer argument");
        return SBA_ERR_ARG;
    }
    int rc = check_words_shape("sba_words_sim_bwd_tc", B_img, B_cap, nef, R, Lw);
    if (rc) return rc;
    return words_sim_bwd_tc5(img, words, cap_lens, d_sim, d_img, d_words, workspace, workspace_bytes, B_img, B_cap, nef, R, Lw,
                             gamma1, gamma2, gamma3, eps, static_cast<cudaStream_t>(stream));
}

size_t sba_words_sim_bwd_workspace_bytes(int B_img, int B_cap, int nef, int R, int Lw) {
    if (B_img <= 0 || B_cap <= 0 || nef <= 0 || R <= 0 || Lw <= 0) return 0;
    return words_bwd_workspace_bytes(B_img, B_cap, nef, R, Lw);
}

int sba_words_sim_bwd(const float* img, const float* words, const int32_t* cap_lens, const float* d_sim, float* d_img,
                      float* d_words, void* workspace, int B_img, int B_cap, int row_offset, int nef, int R, int Lw,
                      float gamma1, float gamma2, float gamma3, float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!img || !words || !cap_lens || !d_sim || !d_img || !workspace) {
        set_error("sba_words_sim_bwd: null pointer argument");
        return SBA_ERR_ARG;
    }
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) {
        set_error("sba_words_sim_bwd: workspace must be 256-byte aligned");
        return SBA_ERR_ALIGN;
    }
    int rc = check_words_shape("sba_words_sim_bwd", B_img, B_cap, nef, R, Lw);
    if (rc) return rc;
    return words_sim_bwd(img, words, cap_lens, d_sim, d_img, d_words, workspace, B_img, B_cap, row_offset, nef, R, Lw,
                         gamma1, gamma2, gamma3, eps, static_cast<cudaStream_t>(stream));
}

int sba_func_attention(const float* query, const float* context, float* wc, float* attn, int B, int nef, int T, int R,
                       float gamma1, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!query || !context || !wc || !attn) {
        set_error("sba_func_attention: null pointer argument");
        return SBA_ERR_ARG;
    }
    int rc = check_words_shape("sba_func_attention", B, B, nef, R, T);
    if (rc) return rc;
    return words_sim_fwd(context, query, nullptr, nullptr, attn, wc, B, B, 0, nef, R, T, gamma1, 1.f, 1.f, 1e-8f, 1,
                         static_cast<cudaStream_t>(stream));
}

int sba_match_ce_fwd(const float* scores, const int32_t* class_ids, const int64_t* labels, float* losses, float* lse, int B,
                     void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!scores || !labels || !losses || !lse || B <= 0) {
        set_error("sba_match_ce_fwd: null pointer argument or B <= 0");
        return SBA_ERR_ARG;
    }
    return match_ce_fwd(scores, class_ids, reinterpret_cast<const long long*>(labels), losses, lse, B,
                        static_cast<cudaStream_t>(stream));
}

int sba_match_ce_bwd(const float* scores, const int32_t* class_ids, const int64_t* labels, const float* lse, const float* g,
                     float* d_scores, int B, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!scores || !labels || !lse || !g || !d_scores || B <= 0) {
        set_error("sba_match_ce_bwd: null pointer argument or B <= 0");
        return SBA_ERR_ARG;
    }
    return match_ce_bwd(scores, class_ids, reinterpret_cast<const long long*>(labels), lse, g, d_scores, B,
                        static_cast<cudaStream_t>(stream));
}

int sba_sent_scores_fwd(const float* cnn, const float* rnn, float* scores, float* norms, int B, int nef, float gamma3,
                        float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!cnn || !rnn || !scores || !norms || B <= 0 || nef <= 0) {
        set_error("sba_sent_scores_fwd: null pointer argument or non-positive size");
        return SBA_ERR_ARG;
    }
    return sent_scores_fwd(cnn, rnn, scores, norms, B, nef, gamma3, eps, static_cast<cudaStream_t>(stream));
}

int sba_sent_scores_bwd(const float* cnn, const float* rnn, const float* norms, const float* scores, const float* d_scores,
                        float* d_cnn, float* d_rnn, int B, int nef, float gamma3, float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!cnn || !rnn || !norms || !scores || !d_scores || !d_cnn || !d_rnn || B <= 0 || nef <= 0) {
        set_error("sba_sent_scores_bwd: null pointer argument or non-positive size");
        return SBA_ERR_ARG;
    }
    return sent_scores_bwd(cnn, rnn, norms, scores, d_scores, d_cnn, d_rnn, B, nef, gamma3, eps,
                           static_cast<cudaStream_t>(stream));
}

int sba_adain_fwd(const void* x, const float* style, void* out, int out_rows, int out_row0, float* stats, int B, int C,
                  int Q, int dtype, float eps, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!x || !style || !out || !stats || B <= 0 || C <= 0 || Q <= 0 || (dtype != SBA_F32 && dtype != SBA_BF16)) {
        set_error("sba_adain_fwd: null pointer, non-positive size or unknown dtype");
        return SBA_ERR_ARG;
    }
    if (out_rows < C || out_row0 < 0 || out_row0 + C > out_rows) {
        set_error("sba_adain_fwd: rows [%d, %d) do not fit a %d-row buffer", out_row0, out_row0 + C, out_rows);
        return SBA_ERR_ARG;
    }
    return adain_fwd(x, style, out, out_rows, out_row0, stats, B, C, Q, dtype, eps, static_cast<cudaStream_t>(stream));
}

int sba_adain_bwd(const void* x, const float* style, const float* stats, const void* g_buf, int g_rows, int g_row0, void* dX,
                  int accumulate, float* d_style, int B, int C, int Q, int dtype, void* stream) {
    g_launches = 0;
    g_err[0] = 0;
    if (!x || !style || !stats || !g_buf || !dX || !d_style || B <= 0 || C <= 0 || Q <= 0 ||
        (dtype != SBA_F32 && dtype != SBA_BF16)) {
        set_error("sba_adain_bwd: null pointer, non-positive size or unknown dtype");
        return SBA_ERR_ARG;
    }
    if (g_rows < C || g_row0 < 0 || g_row0 + C > g_rows) {
        set_error("sba_adain_bwd: rows [%d, %d) do not fit a %d-row buffer", g_row0, g_row0 + C, g_rows);
        return SBA_ERR_ARG;
    }
    return adain_bwd(x, style, stats, g_buf, g_rows, g_row0, dX, accumulate, d_style, B, C, Q, dtype,
                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"

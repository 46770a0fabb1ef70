// Kernel (a): fused GlobalAttentionGeneral forward on the 5th-generation tensor cores
// (SBA_ALGO_TCGEN05): TMA tensor loads -> tcgen05.mma -> TMEM -> one pixel per thread.
//
// One persistent CTA = 1 TMA producer warp + 1 MMA-issuing warp + 4 consumer warps, working
// through 128-pixel tiles: a contiguous static share first, then small chunks fetched from a
// global counter (tc5_common.cuh: ChunkReader; SMs stream at different rates under load):
//   producer : 2-D TMA box loads of the [idf x 128 px] tile of x (128-byte swizzled rows) into a
//              shared-memory ring; the tile as it lands is the MN-major A operand of MMA1.
//   MMA warp : MMA1  S[128 x 32]  = x_tile^T . (log2e * sourceT)      (K = idf)
//              MMA2  c[128 x idf] = P . sourceT^T                     (A = P in TMEM, K = words)
//              completion is signalled with tcgen05.commit on mbarriers.
//   consumers: thread = pixel.  tcgen05.ld S row -> masked softmax over words in registers (no
//              shuffles) -> P written back to TMEM as MMA2's A operand -> attention map staged
//              [word][32 px] per warp, one TMA box store -> tcgen05.ld c row -> staged, TMA box store.
// S is double buffered in TMEM (MMA1 of the next tile runs ahead of the softmax of this one).
// The B operands (sourceT of the current sample, both orientations, zero padded) are rebuilt in
// shared memory by the consumers whenever the tile sequence enters a new sample (prepared one
// tile ahead: srcT is fetched under MMA2, the swap follows c_full).
// bf16 tensors: single bf16 MMAs.  fp32 tensors: 3xTF32 - every operand is split into a tf32
// "hi" part and a residual "lo" part and multiplied as hi.hi + lo.hi + hi.lo (~2^-21 relative);
// the hi part of x is the tile itself (the tensor core reads the top 19 bits), its lo part
// x - trunc(x) is written by the consumers into a second shared-memory tile.
//
// Reference semantics: AttnGAN2/code/GlobalAttention.py:82-121 (oracle/attention.py).
#include "host_util.h"
#include "kernels.h"
#include "tc5_common.cuh"

namespace sba {
namespace tc5 {

namespace {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda); resolved once (thread-safe static)
EncodeFn encode_fn() {
    static const EncodeFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
        return (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) ? reinterpret_cast<EncodeFn>(f) : (EncodeFn) nullptr;
    }();
    return fn;
}
// A tensor map is a pure function of (base, dtype, shape, box, swizzle): the training loop calls with the same
// few tensors' shapes over and over (and the caching allocator hands the same blocks back), so the last maps
// are kept per host thread instead of being re-encoded on every launch.
struct MapKey {
    const void* base;
    int dtype, rows, cols, box_rows, box_cols, swizzle;
    bool operator==(const MapKey& o) const {
        return base == o.base && dtype == o.dtype && rows == o.rows && cols == o.cols && box_rows == o.box_rows &&
               box_cols == o.box_cols && swizzle == o.swizzle;
    }
};
constexpr int kMapCache = 32;
struct MapCache {
    MapKey key[kMapCache];
    CUtensorMap map[kMapCache];
    int used = 0, next = 0;
};
thread_local MapCache g_maps;
}  // namespace

int make_tile_map(CUtensorMap* out, const void* base, int dtype, int rows, int cols, int box_rows, int box_cols,
                  bool swizzle) {
    const MapKey key{base, dtype, rows, cols, box_rows, box_cols, swizzle ? 1 : 0};
    MapCache& mc = g_maps;
    for (int i = 0; i < mc.used; ++i)
        if (mc.key[i] == key) {
            *out = mc.map[i];
            return SBA_OK;
        }
    bind_context_to_thread();
    const EncodeFn encode = encode_fn();
    if (encode == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SBA_ERR_CUDA;
    }
    const int es = dtype == SBA_F32 ? 4 : 2;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * es};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    // 32-bit MN-major UMMA operands exist only in the 32-byte-atom flavour of the 128-byte swizzle
    const CUtensorMapSwizzle sw = !swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                                  : dtype == SBA_F32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = encode(out, dtype == SBA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a %d x %d tensor, box %d x %d", (int)r, rows, cols,
                  box_rows, box_cols);
        return SBA_ERR_CUDA;
    }
    const int slot = mc.used < kMapCache ? mc.used++ : (mc.next = (mc.next + 1) % kMapCache);
    mc.key[slot] = key;
    mc.map[slot] = *out;
    return SBA_OK;
}

}  // namespace tc5

namespace {
using namespace tc5;

struct Tc5FwdParams {
    float* srcT;
    const float* ctx;     // [B, cdf, L]
    const float* W;       // [idf, cdf]
    int cdf;
    const uint8_t* mask;
    void* c_code;         // [B, c_rows, Q]: weightedContext goes to rows [c_row0, c_row0 + idf) of every sample
    int c_rows, c_row0;   // (c_rows = idf, c_row0 = 0: a plain [B, idf, Q] tensor; 2*idf / idf: second half of the
                          //  concatenated h_c_code buffer of NEXT_STAGE_G, model_bert.py:460-461)
    void* attn;
    uint32_t* mask_bits;
    int B, L, Q, mask_mode;
    int tiles_per_sample;
    int n_tiles;
    uint32_t* sched;      // dynamic tile scheduler: chunk counter (scratch word, zeroed by k_project_tc5)
    int static_tiles;     // tiles of the contiguous static share every CTA starts with (0: all tiles are dynamic)
    int chunk;            // tiles per dynamic chunk
    int dyn_first;        // first tile of the dynamic region = gridDim.x * static_tiles
    unsigned long long* tl;   // development timeline stamps (tc5_common.cuh), NULL in the product library
    int phase;            // SBA_PHASE_*: host side only
};

constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <typename T, int IDF, int NQ>
struct Tc5FwdCfg {
    static constexpr bool F32 = sizeof(T) == 4;
    static constexpr int ES = (int)sizeof(T);
    static constexpr int LP = 4 * NQ;                          // words held per thread (L rounded up to 4)
    static constexpr int BOX_PX = 128 / ES;                    // pixels per 128-byte box row
    static constexpr int NBOX = TQ / BOX_PX;                   // boxes per tile: 2 (bf16) / 4 (fp32)
    static constexpr int BOX_BYTES = IDF * 128;
    static constexpr int STAGE_BYTES = NBOX * BOX_BYTES;       // = IDF * TQ * ES
    static constexpr int KMMA = 32 / ES;                       // K per instruction: 16 (bf16) / 8 (tf32)
    static constexpr int KS1 = IDF / KMMA;                     // MMA1 k-steps (channels)
    static constexpr int KSTEP_A = KMMA * 128;                 // bytes between k-steps of the swizzled x tile
    static constexpr int K2 = ((LP + KMMA - 1) / KMMA) * KMMA; // MMA2 K extent (words, zero padded)
    static constexpr int KS2 = K2 / KMMA;
    static constexpr int NS = 32;                              // MMA1 N (words, zero padded)
    static constexpr int KCH1 = IDF * ES / 16, KCH2 = K2 * ES / 16;
    static constexpr int B1_BYTES = NS * IDF * ES;
    static constexpr int B2_BYTES = IDF * K2 * ES;
    static constexpr int NSPLIT = F32 ? 2 : 1;
#ifndef SBA_NST_BF16
#define SBA_NST_BF16 3      // measured: 2 and 3 beat 4 by 2-7 % (smaller footprint; the ring never runs dry), 6 loses a CTA per SM
#endif
#ifndef SBA_NST_F32
#define SBA_NST_F32 3
#endif
    static constexpr int NST = F32 ? SBA_NST_F32 : SBA_NST_BF16;   // x ring depth (compile-time switches for A/B measurements)
    static constexpr int NLO = F32 ? 1 : 0;                    // lo tile (fp32)
    static constexpr int PC = F32 ? K2 : K2 / 2;               // TMEM columns of one P operand
    static constexpr int COL_S = 0, COL_P = 64, COL_PLO = 96, COL_C = F32 ? 128 : 96;
    static constexpr int TMEM_COLS = pow2_cols(COL_C + IDF);
    // per-warp output staging: [LP words][32 px] and [IDF channels][32 px], dense rows (TMA store boxes)
    static constexpr int OUT_ROW = 32 * ES;
    static constexpr int OUT_A_BYTES = LP * OUT_ROW, OUT_C_BYTES = IDF * OUT_ROW;
    static constexpr int OUT_WARP_BYTES = OUT_A_BYTES + OUT_C_BYTES;
    static constexpr int SMEM_BYTES = (NST + NLO) * STAGE_BYTES + NSPLIT * (B1_BYTES + B2_BYTES) + 4 * OUT_WARP_BYTES;
    // MN-major x tile: bf16 = SWIZZLE_128B atoms of 8 channel rows (1024 B); tf32 = SWIZZLE_128B_BASE32B
    // atoms of 4 channel rows (512 B), the only MN-major layout 32-bit operands have
    static constexpr uint32_t A_SWIZZLE = F32 ? kSwizzle128B_Base32B : kSwizzle128B;
    static constexpr uint32_t A_SBO = F32 ? 512 : 1024;
    static constexpr uint32_t IDESC1 = make_idesc(F32 ? 2 : 1, 1, 0, TQ, NS);
    static constexpr uint32_t IDESC2 = make_idesc(F32 ? 2 : 1, 0, 0, TQ, IDF);
    static constexpr int CTAS_PER_SM = 512 / TMEM_COLS;        // resident CTAs must all fit their TMEM allocation
    static_assert(IDF % 16 == 0 && IDF <= 128, "idf must be a multiple of 16");
    static_assert(LP <= 32 && PC <= 32 && COL_PLO + PC <= 128, "at most 32 words");
};

// srcT = W . ctx (the bias-free 1x1 conv_context, GlobalAttention.py:95-97) as a small grid in
// front of the streaming kernel, chained with programmatic dependent launch: block = (sample,
// 8 output channels); warp = one eighth of the cdf reduction, lane = word.  Everything a warp
// needs is requested up front (its W slice through warp-private shared memory, up to 32 ctx rows
// in registers), so the kernel lasts about one memory round trip plus 256 FMAs per lane.
__global__ void __launch_bounds__(256) k_project_tc5(const float* __restrict__ ctx, const float* __restrict__ W,
                                                     float* __restrict__ srcT, const uint8_t* __restrict__ mask,
                                                     uint32_t* __restrict__ mask_bits, uint32_t* __restrict__ sched, int idf,
                                                     int cdf, int L, int early_trigger, unsigned long long* tl) {
    // Launched as a programmatic dependent of whatever precedes it in the stream, which only hides its launch
    // latency: it waits for that work to complete BEFORE it lets its own dependent (the streaming kernel,
    // whose producer starts reading x at once) go, so nothing downstream can run ahead of upstream results.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) SBA_TL(tl, 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) *sched = 0u;        // chunk counter of the streaming kernel's tile scheduler
    // caption padding mask -> one 32-bit word per caption (bit l = word l is padding), by the first block of each sample;
    // the byte is requested here, with everything else, and consumed after the FMAs (one memory round trip in all)
    const bool mask_warp = mask != nullptr && blockIdx.x % (idf / 8) == 0 && threadIdx.x < 32;
    uint8_t mask_byte = 0;
    if (mask_warp && (int)threadIdx.x < L) mask_byte = mask[(size_t)(blockIdx.x / (idf / 8)) * L + threadIdx.x];
    __shared__ __align__(16) float w_s[8][8][32];
    __shared__ float red_s[8][8][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rg = idf / 8;
    const int b = blockIdx.x / rg, i0 = (blockIdx.x - b * rg) * 8;
    const int kq = cdf / 8 / 4 * 4 + ((cdf / 8) % 4 ? 4 : 0);          // slice length, rounded up to a multiple of 4
    const int cbeg = warp * kq;                                        // (cdf % 16 == 0: slices tile cdf, the last may be short)
    const int cend = cbeg + kq < cdf ? cbeg + kq : cdf;
    const int lw = lane < L ? lane : L - 1;
    const float* cb = ctx + (size_t)b * cdf * L + lw;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
        const int kc = cend - c0 < 32 ? cend - c0 : 32;               // multiple of 4
        // one memory round trip: this lane's share of the W slice (8 rows x kc) and its ctx values are all
        // requested before anything is consumed
        float4 wq[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int o = lane + 32 * r;
            const int row = o / (kc / 4), c4 = o - row * (kc / 4);
            wq[r] = (o < 8 * (kc / 4)) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(i0 + row) * cdf + c0) + c4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = (c < kc) ? __ldg(cb + (size_t)(c0 + c) * L) : 0.f;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int o = lane + 32 * r;
            const int row = o / (kc / 4), c4 = o - row * (kc / 4);
            if (o < 8 * (kc / 4)) reinterpret_cast<float4*>(&w_s[warp][row][0])[c4] = wq[r];
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            if (c < kc) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&w_s[warp][k][c]);
                    acc[k] = fmaf(w4.x, v[c], acc[k]);
                    acc[k] = fmaf(w4.y, v[c + 1], acc[k]);
                    acc[k] = fmaf(w4.z, v[c + 2], acc[k]);
                    acc[k] = fmaf(w4.w, v[c + 3], acc[k]);
                }
            }
        }
    }
    if (!early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (mask_warp) {
        const uint32_t bits = __ballot_sync(0xffffffffu, mask_byte != 0);
        if (threadIdx.x == 0) mask_bits[blockIdx.x / (idf / 8)] = bits;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red_s[warp][k][lane] = acc[k];
    __syncthreads();
    {
        const int k = tid >> 5, l = tid & 31;
        if (l < L) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += red_s[w][k][l];
            srcT[((size_t)b * idf + i0 + k) * L + l] = a;
        }
    }
    if (threadIdx.x == 0) SBA_TL(tl, 1);
}


// one arrival per consumer warp: every lane orders its tcgen05 / shared-memory traffic first
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

template <typename T, int IDF, int NQ>
__global__ void __launch_bounds__(kThreads, Tc5FwdCfg<T, IDF, NQ>::CTAS_PER_SM)
    k_attn_fwd_tc5(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tma_attn,
                   const __grid_constant__ CUtensorMap tma_c, const Tc5FwdParams p) {
    using C = Tc5FwdCfg<T, IDF, NQ>;
    constexpr bool F32 = C::F32;
    constexpr int LP = C::LP, NST = C::NST, ES = C::ES;
    constexpr float kLog2e = 1.4426950408889634f;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzle atoms want 1024-byte alignment
    unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t s_x = sbase;                                           // [NST] x tiles
    const uint32_t s_lo = s_x + NST * C::STAGE_BYTES;                     // [NLO] x - trunc(x) tile (fp32)
    const uint32_t s_b1 = s_lo + C::NLO * C::STAGE_BYTES;                 // [NSPLIT] log2e * sourceT, rows = words
    const uint32_t s_b2 = s_b1 + C::NSPLIT * C::B1_BYTES;                 // [NSPLIT] sourceT, rows = channels
    const uint32_t s_out = s_b2 + C::NSPLIT * C::B2_BYTES;                // [4 warps] output staging
    unsigned char* g_lo = sgen + NST * C::STAGE_BYTES;
    unsigned char* g_b1 = g_lo + C::NLO * C::STAGE_BYTES;
    unsigned char* g_b2 = g_b1 + C::NSPLIT * C::B1_BYTES;
    unsigned char* g_out = g_b2 + C::NSPLIT * C::B2_BYTES;

    __shared__ __align__(8) unsigned long long bar_x_full[NST], bar_x_empty[NST], bar_s_full[2], bar_s_free[2], bar_lo_ready,
        bar_p_ready, bar_c_full, bar_b_ready, bar_q_full[kQueueDepth], bar_q_empty[kQueueDepth];
    __shared__ int2 q_ent[kQueueDepth];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L, TPS = p.tiles_per_sample;
    // a programmatic dependent behind this grid (the head kernel of the next call) may become resident now; it
    // parks in griddepcontrol.wait until this grid has completed, which takes its launch latency off the stream
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) SBA_TL(p.tl, 2);
    const int CH = p.chunk, ST = p.static_tiles;
    // first chunk: this CTA's static share, or (no static share) dynamic chunk blockIdx.x
    const int w_begin = (int)blockIdx.x * (ST > 0 ? ST : CH);
    const int cnt0 = ST > 0 ? ST : (p.n_tiles - w_begin < CH ? p.n_tiles - w_begin : CH);
    const int b0 = w_begin / TPS, t0 = w_begin - b0 * TPS;
    const int n_pre = cnt0 < NST ? cnt0 : NST;              // tiles whose loads thread 0 issues before the prologue

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            mbar_init(smem_u32(&bar_x_full[s]), 1);
            mbar_init(smem_u32(&bar_x_empty[s]), 1);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_s_full[s]), 1);
            mbar_init(smem_u32(&bar_s_free[s]), 4);
        }
        mbar_init(smem_u32(&bar_lo_ready), 4);
        mbar_init(smem_u32(&bar_p_ready), 4);
        mbar_init(smem_u32(&bar_c_full), 1);
        mbar_init(smem_u32(&bar_b_ready), 4);
#pragma unroll
        for (int s = 0; s < kQueueDepth; ++s) {
            mbar_init(smem_u32(&bar_q_full[s]), 1);
            mbar_init(smem_u32(&bar_q_empty[s]), 5);       // MMA warp + 4 consumer warps
        }
        fence_barrier_init();
        // the first ring of x tiles is requested before the rest of the prologue (TMEM allocation, operand
        // buffers) so that its DRAM latency runs under it; x does not depend on the projection grid
        {
            int b = b0, t = t0;
            for (int j = 0; j < n_pre; ++j) {
                const uint32_t full = smem_u32(&bar_x_full[j]);
                mbar_expect_tx(full, (uint32_t)C::STAGE_BYTES);
#pragma unroll
                for (int bx = 0; bx < C::NBOX; ++bx)
                    tma_load_2d(s_x + j * C::STAGE_BYTES + bx * C::BOX_BYTES, &tmx, t * TQ + bx * C::BOX_PX, b * IDF, full);
                if (++t == TPS) { t = 0; ++b; }
            }
        }
        prefetch_tensormap(&tma_attn);
        prefetch_tensormap(&tma_c);
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
    // zero the B operand buffers once: the padding (words >= L) is never written again
    for (int o = tid; o < C::NSPLIT * (C::B1_BYTES + C::B2_BYTES) / 16; o += kThreads)
        reinterpret_cast<uint4*>(g_b1)[o] = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);     // provably warp-uniform


    if (warp == kProducerWarp) {
        // --------------------------------- TMA producer -----------------------------------------
        // (the whole warp runs the loop; one elected lane issues - see elect_one())
        int j = 0;                                             // tiles issued so far (ring position)
        int first = w_begin, cnt = cnt0;
        for (int n = 0;; ++n) {
            const int slot = n & (kQueueDepth - 1);
            if (n >= kQueueDepth) mbar_wait(smem_u32(&bar_q_empty[slot]), (uint32_t)((n / kQueueDepth) - 1) & 1u);
            if (lane == 0) {
                q_ent[slot] = make_int2(first, cnt);
                mbar_arrive(smem_u32(&bar_q_full[slot]));          // release: the entry is visible to the readers
            }
            __syncwarp();
            if (cnt == 0) break;
            int bq = first / TPS, t = first - bq * TPS;
            for (int i = 0; i < cnt; ++i, ++j, ++t) {
                if (t == TPS) { t = 0; ++bq; }
                if (n == 0 && i < n_pre) continue;                 // issued by thread 0 before the prologue
                const int stage = j % NST;
                if (j >= NST) mbar_wait(smem_u32(&bar_x_empty[stage]), (uint32_t)((j / NST) - 1) & 1u);
                const uint32_t full = smem_u32(&bar_x_full[stage]);
                const uint32_t dst = s_x + stage * C::STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(full, (uint32_t)C::STAGE_BYTES);
#pragma unroll
                    for (int bx = 0; bx < C::NBOX; ++bx)
                        tma_load_2d(dst + bx * C::BOX_BYTES, &tmx, t * TQ + bx * C::BOX_PX, bq * IDF, full);
                }
                __syncwarp();
            }
            // next chunk: the counter was zeroed by k_project_tc5, complete once griddepcontrol.wait returns
            if (n == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
            int next = 0;
            if (lane == 0) next = (ST > 0 ? 0 : (int)gridDim.x) + (int)atomicAdd(p.sched, 1u);
            next = __shfl_sync(0xffffffffu, next, 0);
            // (64-bit: the counter keeps growing while the last CTAs drain)
            const long long nf = (long long)p.dyn_first + (long long)next * CH;
            first = nf < p.n_tiles ? (int)nf : p.n_tiles;
            cnt = p.n_tiles - first < CH ? p.n_tiles - first : CH;
        }
    } else if (warp == kMmaWarp) {
        // --------------------------------- MMA issuer -------------------------------------------
        // (the whole warp runs the loop; one elected lane issues - see elect_one())
        // descriptor halves (tc5_common.cuh): x tile / lo tile MN-major swizzled, B operands K-major plain
        constexpr uint32_t kAHi = desc_hi(C::A_SBO, C::A_SWIZZLE);
        constexpr uint32_t kBHi1 = desc_hi(C::KCH1 * 128, kSwizzleNone), kBHi2 = desc_hi(C::KCH2 * 128, kSwizzleNone);
        const uint32_t a_lo0 = desc_lo(s_x, C::BOX_BYTES), alo_lo = desc_lo(s_lo, C::BOX_BYTES);
        const uint32_t b1_lo = desc_lo(s_b1, 128), b2_lo = desc_lo(s_b2, 128);
        // MMA1(j): S[j & 1] = x_tile^T . B1   (3xTF32: hi.hi + hi.lo + lo.hi)
        auto mma1 = [&](int j) {
            const int stage = j % NST, buf = j & 1;
            mbar_wait(smem_u32(&bar_x_full[stage]), (uint32_t)(j / NST) & 1u);
            if constexpr (F32) mbar_wait(smem_u32(&bar_lo_ready), (uint32_t)j & 1u);
            if (j >= 2) mbar_wait(smem_u32(&bar_s_free[buf]), (uint32_t)((j >> 1) - 1) & 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + C::COL_S + 32 * buf;
            const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (C::STAGE_BYTES >> 4));
            if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < C::KS1; ++ks) {
                const uint32_t ka = (uint32_t)(ks * (C::KSTEP_A >> 4)), kb = (uint32_t)(ks * 16);
                umma_ss<F32>(d, a_lo + ka, kAHi, b1_lo + kb, kBHi1, C::IDESC1, ks > 0 ? 1u : 0u);
                if constexpr (F32) {
                    umma_ss<F32>(d, a_lo + ka, kAHi, b1_lo + (uint32_t)(C::B1_BYTES >> 4) + kb, kBHi1, C::IDESC1, 1u);
                    umma_ss<F32>(d, alo_lo + ka, kAHi, b1_lo + kb, kBHi1, C::IDESC1, 1u);
                }
            }
            umma_commit(smem_u32(&bar_x_empty[stage]));
            umma_commit(smem_u32(&bar_s_full[buf]));
            }
            __syncwarp();
        };
        // MMA2(j): c = P . B2
        auto mma2 = [&](int j) {
            mbar_wait(smem_u32(&bar_p_ready), (uint32_t)j & 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + C::COL_C;
            if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < C::KS2; ++ks) {
                const uint32_t a = tmem_base + C::COL_P + ks * 8;        // 8 columns per k-step either way
                const uint32_t kb = (uint32_t)(ks * 16);
                umma_ts<F32>(d, a, b2_lo + kb, kBHi2, C::IDESC2, ks > 0 ? 1u : 0u);
                if constexpr (F32) {
                    umma_ts<F32>(d, a, b2_lo + (uint32_t)(C::B2_BYTES >> 4) + kb, kBHi2, C::IDESC2, 1u);
                    umma_ts<F32>(d, tmem_base + C::COL_PLO + ks * 8, b2_lo + kb, kBHi2, C::IDESC2, 1u);
                }
            }
            umma_commit(smem_u32(&bar_c_full));
            }
            __syncwarp();
        };
        ChunkReader rd;
        rd.init(smem_u32(&bar_q_full[0]), smem_u32(&bar_q_empty[0]), q_ent, TPS, lane);
        uint32_t nb = 0;
        if (!rd.done()) {
            mbar_wait(smem_u32(&bar_b_ready), nb & 1u);
            ++nb;
            mma1(0);
        }
        for (int j = 0; !rd.done(); ++j) {
            const int bn = rd.peek_sample(lane);
            const bool has_next = bn >= 0;
            const bool next_same = bn == rd.b;
            // bf16: S of the next tile is produced ahead of the softmax of this one; fp32: the lo tile of
            // the next tile is written only after P of this one, so MMA2 goes first
            if (!F32 && next_same) mma1(j + 1);
            mma2(j);
            if (has_next && !next_same) {
                mbar_wait(smem_u32(&bar_b_ready), nb & 1u);     // operands of the next sample are in place
                ++nb;
                mma1(j + 1);
            } else if (F32 && next_same) {
                mma1(j + 1);
            }
            rd.advance(lane);
        }
    } else {
        // --------------------------------- consumers: thread = pixel ----------------------------
        const int ct = tid - 64;                       // 0..127
        const int cw = warp & 3;                       // TMEM lane quarter this warp may access
        const int px = cw * 32 + lane;                 // pixel (= accumulator row) within the tile
        const uint32_t tl = tmem_base + ((uint32_t)(cw * 32) << 16);
        const uint32_t pad_bits = (L < 32) ? ~((1u << L) - 1u) : 0u;
        const uint32_t Bu = (uint32_t)p.B;
        const uint32_t step_mod = (uint32_t)TQ % Bu;
        // reference mask order: pixel n = b*Q + q uses caption n mod B (GlobalAttention.py:104-108)
        uint32_t cap = 0;
        int b = 0, t = 0, cur_b = -1;
        ChunkReader rd;
        rd.init(smem_u32(&bar_q_full[0]), smem_u32(&bar_q_empty[0]), q_ent, TPS, lane);
        // this warp's output staging: lane = pixel column
        const uint32_t so_a = s_out + cw * C::OUT_WARP_BYTES, so_c = so_a + C::OUT_A_BYTES;
        T* go_a = reinterpret_cast<T*>(g_out + cw * C::OUT_WARP_BYTES) + lane;
        T* go_c = reinterpret_cast<T*>(g_out + cw * C::OUT_WARP_BYTES + C::OUT_A_BYTES) + lane;

        // fp32: lo tile of local tile j = x - trunc(x), elementwise on the swizzled image
        auto make_lo = [&](int j) {
            if constexpr (F32) {
                const int stage = j % NST;
                mbar_wait(smem_u32(&bar_x_full[stage]), (uint32_t)(j / NST) & 1u);
                const float4* xs = reinterpret_cast<const float4*>(sgen + stage * C::STAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(g_lo);
#pragma unroll
                for (int r = 0; r < C::STAGE_BYTES / 16 / kConsumers; ++r) {
                    float4 v = xs[r * kConsumers + ct];
                    v.x -= tf32_trunc(v.x); v.y -= tf32_trunc(v.y); v.z -= tf32_trunc(v.z); v.w -= tf32_trunc(v.w);
                    lo[r * kConsumers + ct] = v;
                }
                fence_proxy_async();
                warp_arrive(smem_u32(&bar_lo_ready), lane);
            }
        };

        // operands of a sample: B1[word][channel] = log2e * srcT, B2[channel][word] = srcT.
        // thread -> (channel ct / 4 [+ 32 g], words (ct % 4) + 4 k): all loads of a thread are issued before the first
        // is consumed (one L2 round trip), and no division by the runtime L.  load_src may run while the MMAs of the
        // previous sample are still in flight; store_ops only after they have completed (c_full of its last tile).
        constexpr int NG = IDF / 32 + (IDF % 32 != 0), NK = LP / 4;
        float sv[NG][NK];
        auto load_src = [&](int bs) {
            const float* sb = p.srcT + (size_t)bs * IDF * L;
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const int ch = (ct >> 2) + 32 * g, l = (ct & 3) + 4 * k;
                    sv[g][k] = (ch < IDF && l < L) ? __ldcg(sb + ch * L + l) : 0.f;
                }
        };
        auto store_ops = [&]() {
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const int ch = (ct >> 2) + 32 * g, l = (ct & 3) + 4 * k;
                    if (ch < IDF && l < L) {
                        const float v = sv[g][k], v1 = v * kLog2e;
                        const uint32_t o1 = kmajor_off<ES>(l, ch, C::KCH1), o2 = kmajor_off<ES>(ch, l, C::KCH2);
                        if constexpr (F32) {
                            const float h1 = tf32_rna(v1), h2 = tf32_rna(v);
                            *reinterpret_cast<float*>(g_b1 + o1) = h1;
                            *reinterpret_cast<float*>(g_b1 + C::B1_BYTES + o1) = tf32_rna(v1 - h1);
                            *reinterpret_cast<float*>(g_b2 + o2) = h2;
                            *reinterpret_cast<float*>(g_b2 + C::B2_BYTES + o2) = tf32_rna(v - h2);
                        } else {
                            *reinterpret_cast<__nv_bfloat16*>(g_b1 + o1) = __float2bfloat16_rn(v1);
                            *reinterpret_cast<__nv_bfloat16*>(g_b2 + o2) = __float2bfloat16_rn(v);
                        }
                    }
                }
            fence_proxy_async();
            warp_arrive(smem_u32(&bar_b_ready), lane);
        };

        asm volatile("griddepcontrol.wait;" ::: "memory");      // srcT of k_project_tc5 is complete and visible
        if (ct == 0) SBA_TL(p.tl, 6);
        if (!rd.done()) make_lo(0);

        for (int j = 0; !rd.done(); ++j) {
            b = rd.b;
            t = rd.t;
            if (rd.chunk_start()) cap = (uint32_t)((((unsigned long long)b * TPS + t) * TQ + px) % Bu);
            if (b != cur_b) {            // first tile of the CTA (later boundaries are prepared one tile ahead, below)
                cur_b = b;
                load_src(b);
                store_ops();
            }

            // ---- S row of this pixel (already in the log2 domain) ---------------------------------
            const int buf = j & 1;
            mbar_wait(smem_u32(&bar_s_full[buf]), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            uint32_t sr[LP];
            tmem_ld<LP>(tl + C::COL_S + 32 * buf, sr);
            tmem_wait_ld();
            tc_fence_before();
            warp_arrive(smem_u32(&bar_s_free[buf]), lane);

            // ---- mask (GlobalAttention.py:104-108) + softmax over words (:109) -----------------------
            uint32_t mb = pad_bits;
            if (p.mask != nullptr) mb |= __ldcg(p.mask_bits + (p.mask_mode == SBA_MASK_PER_SAMPLE ? (uint32_t)b : cap));   // written by the PDL predecessor: no ld.global.nc
            float s[LP];
            float m = -INFINITY;
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                s[l] = ((mb >> l) & 1u) ? -INFINITY : __uint_as_float(sr[l]);
                m = fmaxf(m, s[l]);
            }
            float sum = 0.f;
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                s[l] = mma::ex2_approx(s[l] - m);          // all-masked row: -inf - -inf = NaN, as the reference
                sum += s[l];
            }
            const float inv = mma::rcp_approx(sum);
#pragma unroll
            for (int l = 0; l < LP; ++l) s[l] *= inv;

            // ---- P -> TMEM (A operand of MMA2) ---------------------------------------------------
            if constexpr (F32) {
                uint32_t ph[C::K2], pl[C::K2];
#pragma unroll
                for (int l = 0; l < C::K2; ++l) {
                    if (l < LP) {
                        const float h = tf32_trunc(s[l]);
                        ph[l] = __float_as_uint(h);
                        pl[l] = __float_as_uint(s[l] - h);
                    } else {
                        ph[l] = 0u;
                        pl[l] = 0u;
                    }
                }
                tmem_st<C::K2>(tl + C::COL_P, ph);
                tmem_st<C::K2>(tl + C::COL_PLO, pl);
            } else {
                uint32_t pk[C::PC];
#pragma unroll
                for (int i = 0; i < C::PC; ++i) pk[i] = (2 * i < LP) ? mma::pack_bf16(s[2 * i], s[2 * i + 1]) : 0u;
                tmem_st<C::PC>(tl + C::COL_P, pk);
            }
            tmem_wait_st();
            tc_fence_before();
            warp_arrive(smem_u32(&bar_p_ready), lane);

            const int bn = rd.peek_sample(lane);
            if constexpr (F32) {
                if (bn >= 0) make_lo(j + 1);               // MMA1(j) has completed, the lo tile is free
            }
            // sample boundary ahead: fetch the next sample's srcT now, under MMA2(j) and the attention-map store
            const bool boundary = bn >= 0 && bn != b;
            if (boundary) load_src(bn);

            // ---- attention map: staged [word][32 px] per warp, one TMA box store -------------------
            const int q0 = t * TQ + cw * 32;
            if (lane == 0) bulk_wait_read<1>();        // the previous attn store has finished reading its staging
            __syncwarp();
#pragma unroll
            for (int l = 0; l < LP; ++l) {
                if constexpr (F32) go_a[l * 32] = s[l];
                else go_a[l * 32] = __float2bfloat16_rn(s[l]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tma_attn, q0, b * L, so_a);
                bulk_commit();
            }

            // ---- c row of this pixel ---------------------------------------------------------------
            mbar_wait(smem_u32(&bar_c_full), (uint32_t)j & 1u);
            tc_fence_after();
            if (boundary) {                            // every MMA of this sample has completed: swap the operands now,
                cur_b = bn;                            // so that MMA1 of the next tile runs under this tile's epilogue
                store_ops();
            }
            if (lane == 0) bulk_wait_read<1>();        // the previous c store has finished reading its staging
            __syncwarp();
#pragma unroll
            for (int h = 0; h < IDF / 16; ++h) {
                uint32_t cr[16];
                tmem_ld<16>(tl + C::COL_C + 16 * h, cr);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if constexpr (F32) go_c[(16 * h + i) * 32] = __uint_as_float(cr[i]);
                    else go_c[(16 * h + i) * 32] = __float2bfloat16_rn(__uint_as_float(cr[i]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tma_c, q0, b * p.c_rows + p.c_row0, so_c);
                bulk_commit();
            }

            rd.advance(lane);
            cap += step_mod;
            if (cap >= Bu) cap -= Bu;
        }
        if (ct == 0) SBA_TL(p.tl, 7);
        if (lane == 0) bulk_wait<0>();                 // all output stores have landed before the CTA retires
        tc_fence_before();
    }

    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (tid == 0) SBA_TL(p.tl, 3);
}

template <typename T, int IDF, int NQ>
int launch_fwd_tc5(const void* x, const Tc5FwdParams& p_in, int dtype, cudaStream_t st) {
    Tc5FwdParams p = p_in;
    using C = Tc5FwdCfg<T, IDF, NQ>;
    auto kern = k_attn_fwd_tc5<T, IDF, NQ>;
    constexpr size_t smem = (size_t)C::SMEM_BYTES + 1024 + 16;
    static_assert(smem <= 220 * 1024, "shared memory budget of the tcgen05 forward exceeded");
    int dev = 0, sms = 0;
    int rc = current_device(&dev, &sms, "attn_fwd(tcgen05)");
    if (rc) return rc;
    static std::atomic<unsigned long long> smem_done{0};
    rc = ensure_dynamic_smem(kern, smem, dev, smem_done, "attn_fwd(tcgen05)");
    if (rc) return rc;
    // Resident CTAs per SM: the occupancy API answers 1 for kernels that allocate tensor memory, the
    // hardware co-schedules as many as registers, shared memory and the 512 TMEM columns allow.
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > C::CTAS_PER_SM) per_sm = C::CTAS_PER_SM;
    if (per_sm < 1) per_sm = 1;
    // Tile schedule: every CTA starts with a contiguous static share (few sample switches), the rest is handed
    // out dynamically in small chunks so that SMs that stream faster take more (per-SM rates differ by +-40 %).
    // Short streams (a few tiles per CTA): whole static shares, single leftover tiles go to whoever is first.
    int pct = -1, ch = -1;
    int early_trigger = 1;          // the streaming kernel may become resident (prologue, x prefetch) under the projection
#ifdef SBA_DEV_AIDS
    {
        const DevTuning& tu = dev_tuning();
        if (tu.late_trigger > 0) early_trigger = 0;
        if (tu.ctas_per_sm > 0) per_sm = tu.ctas_per_sm;
        pct = tu.static_pct;
        ch = tu.chunk;
    }
#endif
    const int max_ctas = sms * per_sm;
    int grid = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;
    if (pct < 0 || pct > 100) pct = p.n_tiles < 8 * grid ? 100 : 80;
    if (ch < 1) ch = p.n_tiles < 8 * grid ? 1 : 2;
    p.static_tiles = (int)((long long)p.n_tiles * pct / 100 / grid);
    p.chunk = ch;
    if (p.static_tiles == 0 && grid > (p.n_tiles + ch - 1) / ch) grid = (p.n_tiles + ch - 1) / ch;   // one first chunk each
    p.dyn_first = grid * p.static_tiles;
    p.tl = SBA_TL_SLOT();
    if (p.phase != SBA_PHASE_SECOND) {
        PdlLaunch pl(dim3(p.B * (IDF / 8)), dim3(256), 0, st);
        cudaError_t pe = cudaLaunchKernelEx(&pl.cfg, k_project_tc5, p.ctx, p.W, p.srcT, p.mask, p.mask_bits, p.sched,
                                            (int)IDF, p.cdf, p.L, early_trigger, p.tl);
        if (pe != cudaSuccess) {
            set_error("project(tcgen05): launch: %s", cudaGetErrorString(pe));
            return SBA_ERR_CUDA;
        }
        add_launches(1);
        if (p.phase == SBA_PHASE_FIRST) return check_launch("project(tcgen05)");
    }
    CUtensorMap tmx, tma_attn, tma_c;
    const int es = dtype == SBA_F32 ? 4 : 2;
    rc = make_tile_map(&tmx, x, dtype, p.B * IDF, p.Q, IDF, 128 / es, true);
    if (!rc) rc = make_tile_map(&tma_attn, p.attn, dtype, p.B * p.L, p.Q, p.L, 32, false);
    if (!rc) rc = make_tile_map(&tma_c, p.c_code, dtype, p.B * p.c_rows, p.Q, IDF, 32, false);
    if (rc) return rc;
    PdlLaunch ml(dim3(grid), dim3(kThreads), smem, st);
    cudaError_t e = cudaLaunchKernelEx(&ml.cfg, kern, tmx, tma_attn, tma_c, p);
    if (e != cudaSuccess) {
        set_error("attn_fwd(tcgen05): launch: %s", cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    add_launches(1);
    return check_launch("attn_fwd(tcgen05)");
}

template <typename T, int IDF>
int dispatch_nq(const void* x, const Tc5FwdParams& p, int dtype, cudaStream_t st) {
    switch ((p.L + 3) / 4) {
        case 1: return launch_fwd_tc5<T, IDF, 1>(x, p, dtype, st);
        case 2: return launch_fwd_tc5<T, IDF, 2>(x, p, dtype, st);
        case 3: return launch_fwd_tc5<T, IDF, 3>(x, p, dtype, st);
        case 4: return launch_fwd_tc5<T, IDF, 4>(x, p, dtype, st);
        case 5: return launch_fwd_tc5<T, IDF, 5>(x, p, dtype, st);
        case 6: return launch_fwd_tc5<T, IDF, 6>(x, p, dtype, st);
        case 7: return launch_fwd_tc5<T, IDF, 7>(x, p, dtype, st);
        case 8: return launch_fwd_tc5<T, IDF, 8>(x, p, dtype, st);
        default: return -1;
    }
}

}  // namespace

bool tc5_supports(const AttnShape& s) {
    if (s.idf != 32 && s.idf != 48 && s.idf != 64) return false;
    if (s.L < 1 || s.L > 32) return false;
    if (s.Q % tc5::TQ != 0) return false;
    if (s.cdf % 16 != 0) return false;                  // k_project_tc5 splits the reduction over 4 warps, float4 loads
    if (s.B > 4096 || (unsigned long long)s.B * s.Q >= (1ull << 31)) return false;
    return true;
}

int tc5_attn_fwd(const void* x, const float* ctx, const float* W, const uint8_t* mask, void* c_code, void* attn,
                 float* srcT, uint32_t* mask_bits, const AttnShape& s, cudaStream_t st) {
    Tc5FwdParams p{};
    p.srcT = srcT; p.ctx = ctx; p.W = W; p.cdf = s.cdf; p.mask = mask; p.c_code = c_code; p.attn = attn; p.mask_bits = mask_bits;
    p.B = s.B; p.L = s.L; p.Q = s.Q; p.mask_mode = s.mask_mode;
    p.c_rows = s.c_rows > 0 ? s.c_rows : s.idf;
    p.c_row0 = s.c_rows > 0 ? s.c_row0 : 0;
    p.tiles_per_sample = s.Q / tc5::TQ;
    p.n_tiles = s.B * p.tiles_per_sample;
    p.sched = mask_bits + s.B;            // scratch holds 3B words: [0, B) mask words, [B] chunk counter
    p.phase = s.phase;
    int rc = -1;
    if (s.dtype == SBA_F32) {
        if (s.idf == 32) rc = dispatch_nq<float, 32>(x, p, s.dtype, st);
        else if (s.idf == 48) rc = dispatch_nq<float, 48>(x, p, s.dtype, st);
        else if (s.idf == 64) rc = dispatch_nq<float, 64>(x, p, s.dtype, st);
    } else {
        if (s.idf == 32) rc = dispatch_nq<__nv_bfloat16, 32>(x, p, s.dtype, st);
        else if (s.idf == 48) rc = dispatch_nq<__nv_bfloat16, 48>(x, p, s.dtype, st);
        else if (s.idf == 64) rc = dispatch_nq<__nv_bfloat16, 64>(x, p, s.dtype, st);
    }
    if (rc == -1) {
        set_error("attn_fwd(tcgen05): unsupported shape idf=%d L=%d", s.idf, s.L);
        return SBA_ERR_UNSUPPORTED;
    }
    return rc;
}

}  // namespace sba

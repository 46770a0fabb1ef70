// Building blocks of the tcgen05 attention family (attn_tc5_fwd.cu / attn_tc5_bwd.cu):
// 2-D TMA tensor loads, UMMA shared-memory / instruction descriptors, tcgen05.mma issue,
// TMEM allocation and TMEM <-> register transfers.
//
// Why tcgen05 fits although idf = 32 and L = 18 are tiny: the M dimension of every contraction
// on this path is the PIXEL index.  A 128-pixel tile of x (as TMA delivers it: [channel][pixel],
// 128-byte swizzled rows) is a legal MN-major A operand, so
//     S[128 px x 32 words] = x_tile^T . sourceT      (M = 128, N = 32, K = idf)
//     c[128 px x idf]      = P[128 x words] . sourceT^T   (A = P from TMEM, K = words)
// run as UMMA instructions issued by one thread, with the accumulators in TMEM.  tcgen05.ld
// (32x32b) hands every thread one whole pixel row, so the softmax over words needs no
// shuffles and the outputs are written one pixel per lane, 128 contiguous bytes per warp.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "mma_common.cuh"

namespace sba {
namespace tc5 {

using mma::smem_u32;
using mma::mbar_init;
using mma::mbar_expect_tx;
using mma::mbar_arrive;
using mma::fence_barrier_init;

// Bounded mbarrier wait: a protocol bug traps (-> a CUDA error the ABI reports) instead of
// hanging the GPU box.  Every try_wait suspends for up to the hinted 20 us, so the bound is
// between tens of milliseconds and ~1 s - far beyond any legitimate wait on this path.
#ifndef SBA_MBAR_SPIN_LIMIT
#define SBA_MBAR_SPIN_LIMIT (1u << 16)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)      // suspend-time hint (ns): sleep in hardware, not in a spin loop
            : "memory");
        if (!done && ++spins > SBA_MBAR_SPIN_LIMIT) __trap();
    } while (!done);
}

// One lane of a fully converged warp.  The producer / MMA warps run their loops warp-uniformly and
// predicate only the issuing instructions on this, so that descriptors, TMEM addresses and
// barrier addresses stay in uniform registers (a divergent `if (lane == 0)` region makes ptxas
// wrap every UTCHMMA / UTMALDG in an elect-and-broadcast loop).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- TMA: 2-D tensor-map box load, completion on an mbarrier --------------------------------
// L2 eviction-priority hint for streaming tensors (read once / written once): evict_first keeps them from
// displacing each other's dirty lines.  Compile-time switch for A/B measurements.
#ifndef SBA_TC5_L2_HINTS
#define SBA_TC5_L2_HINTS 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
#if SBA_TC5_L2_HINTS & 1
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar), "l"(l2_policy_evict_first())
        : "memory");
#else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
#endif
}
// shared -> global box store (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
#if SBA_TC5_L2_HINTS & 2
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tm),
                 "r"(c0), "r"(c1), "r"(src), "l"(l2_policy_evict_first())
                 : "memory");
#else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(src)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {      // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all previously issued tcgen05.mma of this thread complete -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32x32b: thread t of warp w reads / writes N consecutive 32-bit columns of TMEM lane 32*(w%4)+t
#define SBA_R4(a, o) "=r"(a[o]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3])
#define SBA_W4(a, o) "r"(a[o]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3])
template <int OFF, class A>
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, A& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : SBA_R4(r, OFF) : "r"(taddr) : "memory");
}
template <int OFF, class A>
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, A& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : SBA_R4(r, OFF), SBA_R4(r, OFF + 4)
                 : "r"(taddr)
                 : "memory");
}
template <int OFF, class A>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, A& r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : SBA_R4(r, OFF), SBA_R4(r, OFF + 4), SBA_R4(r, OFF + 8), SBA_R4(r, OFF + 12)
        : "r"(taddr)
        : "memory");
}
template <int OFF, class A>
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const A& r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), SBA_W4(r, OFF) : "memory");
}
template <int OFF, class A>
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const A& r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), SBA_W4(r, OFF),
                 SBA_W4(r, OFF + 4)
                 : "memory");
}
template <int OFF, class A>
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const A& r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        SBA_W4(r, OFF), SBA_W4(r, OFF + 4), SBA_W4(r, OFF + 8), SBA_W4(r, OFF + 12)
        : "memory");
}
#undef SBA_R4
#undef SBA_W4

// N columns (N % 4 == 0) starting at register index OFF, as the fewest x16 / x8 / x4 transfers
template <int N, int OFF = 0, class A>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, A& r) {
    if constexpr (N >= 16) { tmem_ld16<OFF>(taddr, r); tmem_ld<N - 16, OFF + 16>(taddr + 16, r); }
    else if constexpr (N >= 8) { tmem_ld8<OFF>(taddr, r); tmem_ld<N - 8, OFF + 8>(taddr + 8, r); }
    else if constexpr (N >= 4) { tmem_ld4<OFF>(taddr, r); tmem_ld<N - 4, OFF + 4>(taddr + 4, r); }
}
template <int N, int OFF = 0, class A>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const A& r) {
    if constexpr (N >= 16) { tmem_st16<OFF>(taddr, r); tmem_st<N - 16, OFF + 16>(taddr + 16, r); }
    else if constexpr (N >= 8) { tmem_st8<OFF>(taddr, r); tmem_st<N - 8, OFF + 8>(taddr + 8, r); }
    else if constexpr (N >= 4) { tmem_st4<OFF>(taddr, r); tmem_st<N - 4, OFF + 4>(taddr + 4, r); }
}

// ---- UMMA descriptors ---------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100): start address [0,14) and leading / stride byte
// offsets [16,30) / [32,46) in 16-byte units, version 1 at [46,48), swizzle mode at [61,64).
constexpr uint32_t kSwizzleNone = 0, kSwizzle128B_Base32B = 1, kSwizzle128B = 2;
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swizzle) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)swizzle << 61);
}
// Instruction descriptor: fp32 accumulate; fmt 1 = bf16 (kind::f16), 2 = tf32 (kind::tf32);
// a_mn / b_mn = 1 when that operand is MN-major (M or N contiguous in shared memory).
constexpr uint32_t make_idesc(int fmt, int a_mn, int b_mn, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The issuing thread is a single dependent instruction stream, so descriptors are kept as two 32-bit
// halves: the high half (stride offset, version, swizzle) is a compile-time constant per operand,
// the low half (start address, leading offset) advances by a constant number of 16-byte units per
// k-step - one integer add per MMA instead of rebuilding 64-bit descriptors.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t swizzle) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (swizzle << 29);
}
template <bool TF32>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
    if constexpr (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
template <bool TF32>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
    if constexpr (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// D[tmem] (+)= A[smem] . B[smem]^T ; one thread issues for the CTA (64-bit descriptor form)
template <bool TF32>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// D[tmem] (+)= A[tmem] . B[smem]^T
template <bool TF32>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// ---- tf32 splitting -----------------------------------------------------------------------------
// The tensor core reads the top 19 bits of an fp32 word (sign, 8 exponent, 10 mantissa bits).
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// K-major, non-swizzled canonical operand: 8-row x 16-byte core matrices of 128 contiguous bytes;
// core (row block rb, k chunk kc) at (rb * KCH + kc) * 128, KCH = chunks along K.
// Descriptor: LBO = 128 (next k chunk), SBO = KCH * 128 (next 8 rows); one MMA consumes 2 chunks.
template <int ES>   // element size in bytes
__device__ __forceinline__ uint32_t kmajor_off(int row, int k, int kch) {
    constexpr int EPC = 16 / ES;
    return (uint32_t)((((row >> 3) * kch + k / EPC) << 7) + ((row & 7) << 4) + (k % EPC) * ES);
}

constexpr int TQ = 128;               // pixels per tile = UMMA M
constexpr int kProducerWarp = 0, kMmaWarp = 1, kFirstConsumerWarp = 2;
constexpr int kConsumers = 128;
constexpr int kThreads = 64 + kConsumers;

// Host side: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda).
// 2-D view [rows][cols = Q] of a contiguous [B, rows_per_sample, Q] tensor with a box of
// box_rows x box_cols elements.  swizzle = true (loads of x / g_c; box_cols * es must be 128):
// 128-byte swizzle in the flavour the MN-major UMMA operand needs - bf16: SWIZZLE_128B atoms of
// 8 rows x 128 B; fp32: SWIZZLE_128B_ATOM_32B (UMMA SWIZZLE_128B_BASE32B), 4 rows x 128 B.
// swizzle = false (stores): dense box rows.
int make_tile_map(CUtensorMap* out, const void* base, int dtype, int rows, int cols, int box_rows, int box_cols,
                  bool swizzle);

// ---- development aids (python -m sba_gan_b200.build --dev; csrc/dev_aids.cu) ------------------------
// Kernel parameter blocks always carry a `tl` pointer (NULL in the product library, where the stamps
// below compile to nothing): 16 globaltimer stamps per ABI call, even = earliest, odd = latest over CTAs:
//   0/1 head kernel (projection) first instruction / last exit     2/3 streaming kernel entry / exit
//   4/5 tail kernel (backward finish) past its grid dependency / exit
//   6   streaming kernel past its grid dependency                  7   streaming kernel: last tile done
//   8/9 tail kernel block entry   10/11 past its dependency   12/13 first round staged   14/15 partial stored + counted
#ifdef SBA_DEV_AIDS
struct DevTuning { int ctas_per_sm, static_pct, chunk, late_trigger, variant; };   // SBA_TC5_CTAS_PER_SM / _STATIC / _CHUNK / _LATE_TRIGGER / _VARIANT, read once
const DevTuning& dev_tuning();
unsigned long long* timeline_slot();                           // device pointer to this call's 8 stamps, or NULL
__device__ __forceinline__ void tl_stamp(unsigned long long* tl, int stamp) {
    if (tl == nullptr) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (stamp & 1) atomicMax(tl + stamp, t);
    else atomicMin(tl + stamp, t);
}
#define SBA_TL(tl, stamp) ::sba::tc5::tl_stamp(tl, stamp)
#define SBA_TL_SLOT() ::sba::tc5::timeline_slot()
#else
#define SBA_TL(tl, stamp) ((void)0)
#define SBA_TL_SLOT() nullptr
#endif

// ----------------------------------------------------------------------------------------------
// Dynamic tile schedule.  SMs do not stream at the same rate once the memory system is saturated
// (per-SM duration of an equal static share spreads by +-40 % on B200, grouped by TPC position),
// so tiles are handed out in chunks: chunk blockIdx.x first, then gridDim.x + atomicAdd(counter).
// The producer warp fetches chunk ids and publishes {first tile, count} entries through a small
// shared-memory queue; the MMA warp and the consumer warps read every entry (count 0 = end).
// ----------------------------------------------------------------------------------------------
constexpr int kQueueDepth = 8;
struct ChunkReader {
    uint32_t full0, empty0;      // shared addresses of the queue barriers
    const volatile int2* q;
    int n;                       // next queue entry to read
    int tps;                     // tiles per sample
    int cnt, i;                  // tiles of the current chunk, position in it
    int b, t;                    // sample and tile-in-sample of the current tile
    int nfirst, ncnt, nb;        // next chunk (valid when have_next)
    bool have_next;
    __device__ __forceinline__ void fetch(int lane) {
        const int slot = n & (kQueueDepth - 1);
        mbar_wait(full0 + 8 * slot, (uint32_t)(n / kQueueDepth) & 1u);
        nfirst = q[slot].x;
        ncnt = q[slot].y;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * slot);
        ++n;
        nb = nfirst / tps;       // the only division: once per chunk
        have_next = true;
    }
    __device__ __forceinline__ void take() {
        cnt = ncnt; i = 0; b = nb; t = nfirst - nb * tps; have_next = false;
    }
    __device__ __forceinline__ void init(uint32_t full, uint32_t empty, const int2* queue, int tiles_per_sample, int lane) {
        full0 = full; empty0 = empty; q = queue; n = 0; tps = tiles_per_sample;
        fetch(lane);
        take();
    }
    __device__ __forceinline__ bool done() const { return cnt == 0; }
    __device__ __forceinline__ bool chunk_start() const { return i == 0; }
    // sample of the tile after this one, -1 if this is the CTA's last tile (may wait for the producer)
    __device__ __forceinline__ int peek_sample(int lane) {
        if (i + 1 < cnt) return t + 1 == tps ? b + 1 : b;
        if (!have_next) fetch(lane);
        return ncnt > 0 ? nb : -1;
    }
    __device__ __forceinline__ void advance(int lane) {
        if (i + 1 < cnt) {
            ++i;
            if (++t == tps) { t = 0; ++b; }
            return;
        }
        if (!have_next) fetch(lane);
        take();
    }
};

}  // namespace tc5
}  // namespace sba

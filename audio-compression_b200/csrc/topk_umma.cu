// topk_umma.cu — A4 on the 5th-generation tensor cores: exact cosine-similarity
// candidate search as a fused split-fp16 tcgen05 contraction (replaces
// range_candidates_from_embedding_emb, /root/reference/fractal.py:535-552, for all
// ranges at once).
//
// Shape of the work: scores = Q (n_q x 16) . E^T (16 x n_d), n_d ~ 2e6..9e7, top
// 32 (or 64) per row.  K = 16 is tiny, so the contraction is "all epilogue": every
// score has to leave TMEM and be looked at once, and a tcgen05.mma with so little K
// is bound by its per-instruction cost, not by the tensor array
// (scripts/umma_microbench.cu on B200: 118 cycles for M=128 N=128 K=16 whatever the
// operand layout, 173 for the CTA-pair M=256 N=256 one that does four times the work).
//
// One kernel skeleton (scan_kernel<MODE, HI, CG>), shared by three epilogues:
//   * CG = 2 (the exact list kernel): a CLUSTER OF TWO CTAs (one TPC) owns 256 queries, 128 per CTA, and
//     issues tcgen05.mma.cta_group::2 with M=256, N=256: each CTA keeps its own 128 query rows (A) and HALF
//     of every 256-domain stage (B) in shared memory; commits are multicast to the mbarriers of BOTH CTAs and
//     the peer relays "my half of the stage has landed" with a remote mbarrier arrive.
//     CG = 1 (the streaming passes): a CTA on its own, M=128, N=256, both tiles of a stage in its shared
//     memory: the same tensor rate per SM and no cross-CTA signalling in the hand-over chain;
//   * operands are pre-split into fp16 hi/lo parts (x = hi + lo, 22 significant
//     bits) and pre-tiled by a pack kernel into the K-major SWIZZLE_32B UMMA layout,
//     so one plain bulk copy (cp.async.bulk, the TMA engine, no tensor map) lands a
//     128-domain tile in shared memory ready for the tensor core;
//   * per stage three K=16 instructions (hi*lo, lo*hi, hi*hi) -- or the hi*hi one alone (HI) where a filter
//     with a 2e-3 error bound is enough -- go into one of two 256-column TMEM accumulator buffers; two
//     issuing threads, one per buffer;
//   * epilogue warps read the accumulators with tcgen05.ld 32x32b.x32 — one query
//     row per thread — hand the buffer back as soon as their share is in registers,
//     reduce each 32-column chunk to its maximum with 3-input max instructions and
//     compare it with the row's threshold; only chunks that reach it do more.
//
// MODE_LISTS keeps a running top-48 by tensor-core score in shared memory and PROVES per row that the exact top-K is
// among them (rows that cannot -- a boundary crowded within the filter's error -- go to the FFMA kernel); small tables, fallback;
// MODE_THETA + MODE_COLLECT + finalize_kernel are the fast path: a per-query
// threshold from a strided sample of the table, one scan that only appends the
// indices of the domains that reach it, then an exact re-score with a proof that
// nothing was missed; queries that fail the proof get a second, full-split pass with
// their own buffers and only then an exact kernel.  Either way every returned candidate carries the canonical
// float32 score (ascending-k FMA chain) and the rows are ordered by it, ties by
// index: the result equals the FFMA kernel's bit for bit.
//
// Bound: tensor pipe, 2*16 algorithmic flop per pair (the split issues 3x that);
// what actually limits it today is in DESIGN.md 4.3.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"
#include "fwav_math.cuh"

namespace {

constexpr int ED = 16;                 // embedding dim this kernel is built for
constexpr int kQTile = 128;            // queries per CTA (UMMA M = 256 across the CTA pair)
constexpr int kQPair = 2 * kQTile;     // queries per cluster
constexpr int kDTile = 128;            // domain rows per packed tile = one CTA's half of a stage
constexpr int kDStage = 2 * kDTile;    // domains per stage (UMMA N)
constexpr int kStages = 8;             // 64 KB of domain tiles in flight per CTA (power of two)
// Epilogue warps per CTA: 8 for MODE_LISTS (two per TMEM lane quadrant, 128 columns each: its lists need the shared
// memory), 16 for the streaming modes (four per quadrant, 64 columns each).  An epilogue warp's stage is a serial chain
// of long-latency operations (mbarrier wait ~90 cycles, tcgen05.ld + wait ~130, 38 half-rate max instructions, arrive);
// measured with two warps per scheduler the chain takes ~1250 cycles a stage while the tensor pipe needs 490, so the
// streaming modes put four warps on every scheduler to hide it.  Then: producer + TMEM alloc warp, two MMA issuer warps.
__host__ __device__ constexpr int epi_warps(int mode) { return mode == 0 ? 8 : 16; }
__host__ __device__ constexpr int n_threads(int mode) { return (epi_warps(mode) + 3) * 32; }
constexpr int kChunks = 4;             // 32-column TMEM chunks per warp and stage (128 columns)
constexpr int kKeep = 48;              // candidates kept per query (top_k <= 32 plus a 16-entry margin)
constexpr int kCap = kKeep;            // eight-byte keys per query row in shared memory
constexpr uint32_t kPartBytes = kDTile * ED * 2;       // 4 KB: one 128-row hi or lo tile (fp16)
constexpr uint32_t kTileBytes = 2 * kPartBytes;        // 8 KB: hi | lo
constexpr uint32_t kSBO = 256;                         // bytes between 8-row groups (8 rows x 32 B)
constexpr unsigned kFull = 0xffffffffu;

enum { MODE_LISTS = 0, MODE_THETA = 1, MODE_COLLECT = 2 };

// shared memory map (dynamic, 1024-aligned; identical in both CTAs of a pair):
//   query tile | barriers | mode-specific area | ring of domain stages
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffBars = kOffA + kTileBytes;
constexpr uint32_t kBarBytes = 16 * kStages + 72;      // full[], empty[], tfull[2], tempty[2], a, done[2], tmem slot
constexpr uint32_t kOffMode = kOffBars + 256;
static_assert(kBarBytes <= 256, "barrier block");
constexpr int kTheta = 16;             // ranks of the merged sample list the probe may ask for; default threshold rank is kTheta / 2
constexpr int kThetaPart = 6;          // kept per column group (four groups per row)
constexpr int kThetaWarm = 24;         // first stages of pass 1 that only look at chunk maxima
// MODE_LISTS: [row][column half][kCap] keys, then one owner's 128 scores per warp
constexpr uint32_t kOffScratch = kOffMode + kQTile * 2 * kCap * 8;
// MODE_THETA: [row][3][kThetaPart] scores of column groups 1..3 for the final merge
__host__ __device__ constexpr uint32_t mode_bytes(int mode) {
    return mode == MODE_LISTS ? (kOffScratch - kOffMode) + 8 * 128 * 4 : mode == MODE_THETA ? kQTile * 3 * kThetaPart * 4 : 0;
}
__host__ __device__ constexpr uint32_t off_ring(int mode) { return (kOffMode + mode_bytes(mode) + 1023u) & ~1023u; }
// a stage is 256 domains: one 8 KB tile per CTA of a pair, or both tiles (16 KB) in a CTA on its own
__host__ __device__ constexpr uint32_t stage_bytes(int cg) { return cg == 1 ? 2 * kTileBytes : kTileBytes; }
__host__ __device__ constexpr uint32_t smem_bytes(int mode, int cg) { return off_ring(mode) + kStages * stage_bytes(cg); }


struct ScanArgs {
    const uint4 *q_tiles;      // packed query tiles (one per CTA)
    const uint4 *e_tiles;      // packed domain tiles (two per stage)
    const float *Q, *E;        // the float32 tables (exact re-score)
    long long n_q, n_d;
    int n_stages;              // stages in the table
    int n_split;               // CTA pairs per 256 queries: each scans 1/n_split of the stages (MODE_LISTS: > 1 writes `parts`)
    unsigned long long *parts; // MODE_LISTS, n_split > 1: [query][split][2 * kCap] tensor-core-score keys for merge_parts_kernel
    int top_k;
    const uint8_t *active;
    int32_t *cand;             // MODE_LISTS out
    float *scores;             // MODE_LISTS out (optional)
    float *theta;              // MODE_THETA out, MODE_COLLECT in
    int hi_rank;               // MODE_THETA: which sampled score estimates the top_k-th best of the table (top_k / stride)
    int theta_rank;            // MODE_THETA: which of the merged best sampled scores becomes theta (1 .. 24)
    float *theta_hi;           // MODE_THETA out: the hi_rank-th best sampled score (about the top_k-th of the table)
    int32_t *cbuf;             // MODE_COLLECT out: [query][split][column group][cap] domain indices
    int *ccount;               // MODE_COLLECT out: [query][split][column group] how many passed (may exceed cap)
    int cap;
    int dbg;
    long long *trace;          // profiling: clock64 stamps of pair 0's leader CTA, [stage][8] (dbg bit 64)
    // MODE_LISTS: rows whose kept lists cannot PROVE the top_k (see lists_unsafe) are appended here for the FFMA kernel
    int *lfail_list, *lfail_count;
    const unsigned *norms;     // largest squared row norms (row_norm2_max_kernel): the error bound of the filter scores
    // collect_hi_kernel: where in the table the CTAs of this launch currently are (stage index relative to the split's
    // first stage).  A CTA starts its lap where the others are and publishes its own position now and then, so that
    // the 148 resident CTAs stream the SAME part of the 63 MB table at any time and L2 serves all but the first
    // touch (NULL: every CTA starts at a position derived from its queries).
    int *front;
};


// UMMA instruction descriptor: D=F32, A=B=F16, both K-major, N=256, M=256 (cta_group::2)
// (acc16: half-precision accumulators, D=F16 -- one value per 32-bit TMEM column, read two per register with
// tcgen05.ld ...pack::16b; see collect_hi_kernel)
constexpr uint32_t idesc_for(int cg, bool acc16 = false) { return ((acc16 ? 0u : 1u) << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kDStage >> 3) << 17) | ((uint32_t)((kQTile * cg) >> 4) << 24); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the pair
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank)
        : "memory");
}
// CTA-scope acquire (a cluster-scope one makes ptxas flush L1 with CCTL.IVALL after every wait): what the
// waiters go on to touch is TMEM, or shared memory through the async proxy, never the other CTA's generic writes
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();   // a lost arrival must fail loudly, never hang the GPU
    }
}
// hot-loop wait: no watchdog (the MMA threads and the producer keep theirs, so a lost arrival still traps)
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
// one lane of a converged warp (the compiler keeps warp-uniform operands of the guarded instructions in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// completion of every MMA issued so far -> the barrier at this offset (CG == 2: in BOTH CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3)
                     : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_32B matrix descriptor (cute::UMMA::SmemDescriptor, version 1): 8-row groups
// kSBO bytes apart, the two 16-byte K chunks of a row adjacent (leading offset 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(kSBO >> 4) << 32) |
           (1ull << 46) | (6ull << 61);
}
// D[tmem] (+)= A[smem] . B[smem]^T, fp16 operands, K = 16, f32 accumulate; CG == 2: M = 256 over the CTA pair
template <int CG, bool A16 = false>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    if (A16)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc_for(1, true)), "r"(accumulate)
            : "memory");
    else if (CG == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc_for(2)), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc_for(1)), "r"(accumulate)
            : "memory");
}

#define FWAV_R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                    "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
                    "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
                    "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
                    "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define FWAV_RW32(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),  \
                     "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),          \
                     "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),        \
                     "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),        \
                     "+r"(v[29]), "+r"(v[30]), "+r"(v[31])

// 32 lanes x 32 columns of one TMEM lane quadrant -> 32 registers per thread (asynchronous)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : FWAV_R32(v)
        : "r"(taddr)
        : "memory");
}
// wait for the loads; the registers are threaded through so no use can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld1(uint32_t (&a)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : FWAV_RW32(a)::"memory");
}
// the compiler forgets what it knows about the 32 values (nothing is emitted): a rare path that starts with this
// recomputes what it needs instead of keeping the common path's intermediate results alive in registers
// (`after`: a value the common path ends with, so that this cannot move above it)
__device__ __forceinline__ void launder32(uint32_t (&a)[32], unsigned after) { asm volatile("" : FWAV_RW32(a) : "r"(after)); }
__device__ __forceinline__ void tmem_wait_ld2(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : FWAV_RW32(a), FWAV_RW32(b)::"memory");
}

// 32 lanes x 64 columns of half-precision accumulators -> 32 registers per thread, two adjacent columns per register
// (column 2j in the low half of register j, column 2j + 1 in the high half)
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : FWAV_R32(v)
        : "r"(taddr)
        : "memory");
}
// Per-half signed 16-bit maximum of three packed registers (DPX, SASS VIMNMX3.S16x2).  fp16 bit patterns of
// non-negative values order like signed 16-bit integers, and every negative value sorts below every non-negative
// one -- all a threshold filter with a non-negative threshold needs.
__device__ __forceinline__ unsigned pmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
// [even-column maximum | odd-column maximum] of a packed 64-column chunk: the same 3-input tree as chunk_max
__device__ __forceinline__ unsigned chunk_max_p(const uint32_t (&v)[32]) {
    unsigned m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = pmax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    m[10] = pmax3(v[30], v[31], v[31]);
    const unsigned a = pmax3(m[0], m[1], m[2]), b = pmax3(m[3], m[4], m[5]), c = pmax3(m[6], m[7], m[8]);
    return pmax3(pmax3(a, b, c), m[9], m[10]);
}
// does either half of v exceed the threshold t1 (packed twice in t1x2)?
__device__ __forceinline__ bool p_beats(unsigned v, unsigned t1x2) { return pmax3(v, t1x2, t1x2) != t1x2; }
// visit the columns of a packed chunk whose value exceeds t1, pruning with the max tree (cf. for_each_ge)
template <class F>
__device__ __forceinline__ void for_each_gt_p(const uint32_t (&v)[32], unsigned t1x2, F f) {
    const int t1 = (int)(short)(t1x2 & 0xffffu);
    auto leaf = [&](int r) {
        if ((int)(short)(v[r] & 0xffffu) > t1) f(2 * r);
        if ((int)(short)(v[r] >> 16) > t1) f(2 * r + 1);
    };
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        unsigned m[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i] = pmax3(v[9 * g + 3 * i], v[9 * g + 3 * i + 1], v[9 * g + 3 * i + 2]);
        if (p_beats(pmax3(m[0], m[1], m[2]), t1x2)) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (p_beats(m[i], t1x2)) {
#pragma unroll
                    for (int e = 0; e < 3; ++e)
                        if (p_beats(v[9 * g + 3 * i + e], t1x2)) leaf(9 * g + 3 * i + e);
                }
            }
        }
    }
    if (p_beats(pmax3(v[27], v[28], v[29]), t1x2)) {
#pragma unroll
        for (int e = 27; e < 30; ++e)
            if (p_beats(v[e], t1x2)) leaf(e);
    }
    if (p_beats(v[30], t1x2)) leaf(30);
    if (p_beats(v[31], t1x2)) leaf(31);
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}


// ---------------------------------------------------------------------------
// pack: row-major f32 (rows x 16) -> 128-row tiles [hi | lo][K chunk 0..1][row] of
// 16-byte granules (8 halves).  x = hi + lo with hi = fp16(x), lo = fp16(x - hi):
// 22 significant bits for |x| in the fp16 normal range and an absolute error below
// 3e-8 otherwise (embedding components are bounded by 1), measured 3.9e-7 max on
// the scores — the same as a float32 sgemv.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_f16_tiles_kernel(const float *__restrict__ src, long long n_rows, long long n_tiles,
                      uint4 *__restrict__ dst, int row_stride) {
    const long long total = n_tiles * (kDTile * 2);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long tile = g / (kDTile * 2);
        const int w = (int)(g - tile * (kDTile * 2));
        const int c = w / kDTile, r = w % kDTile;       // c: which 8-element K chunk
        long long row = tile * kDTile + r;
        if (row_stride > 1) {
            // the sample table of pass 1: every row_stride-th domain, dealt round-robin to the four 64-column
            // groups of a stage, so that the neighbouring samples of one similarity peak do not all land in the
            // same epilogue warp's column group (pass 1 needs no indices, any order will do)
            const long long in_stage = row & (kDStage - 1);
            row = ((row & ~(long long)(kDStage - 1)) + (in_stage & 63) * 4 + (in_stage >> 6)) * row_stride;
        }
        float xs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < n_rows) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(src + row * ED + c * 8));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(src + row * ED + c * 8 + 4));
            xs[0] = a.x; xs[1] = a.y; xs[2] = a.z; xs[3] = a.w; xs[4] = b.x; xs[5] = b.y; xs[6] = b.z; xs[7] = b.w;
        }
        __half2 hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half h0 = __float2half_rn(xs[2 * i]), h1 = __float2half_rn(xs[2 * i + 1]);
            hi[i] = __halves2half2(h0, h1);
            lo[i] = __halves2half2(__float2half_rn(xs[2 * i] - __half2float(h0)),
                                   __float2half_rn(xs[2 * i + 1] - __half2float(h1)));
        }
        uint4 *t = dst + tile * (2 * kDTile * 2);
        // SWIZZLE_32B K-major atom: 8 rows x 32 bytes, 16-byte chunk index XOR (row >> 2) & 1
        const int slot = (r >> 3) * 16 + (r & 7) * 2 + (c ^ ((r >> 2) & 1));
        t[slot] = *reinterpret_cast<const uint4 *>(hi);
        t[kDTile * 2 + slot] = *reinterpret_cast<const uint4 *>(lo);
    }
}

// largest squared row norm of a (rows x 16) table, as a float bit pattern (non-negative floats order like their
// bits); slightly above the exact value whatever the rounding of the sixteen FMAs
__global__ void __launch_bounds__(256)
row_norm2_max_kernel(const float *__restrict__ x, long long n_rows, unsigned *__restrict__ out) {
    float n2 = 0.0f;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        float ss = 0.0f;
#pragma unroll
        for (int k = 0; k < ED; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * ED + k));
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        n2 = fmaxf(n2, ss * (1.0f + 2e-6f));
    }
    const unsigned u = __reduce_max_sync(kFull, __float_as_uint(n2));
    if ((threadIdx.x & 31) == 0 && u) atomicMax(out, u);
}

// ---------------------------------------------------------------------------
// Error bounds of the two filters ("turn the slacks into bounds").  nq = max |q|, ne = max |e| (row_norm2_max_kernel), B = nq * ne >= sum_k |q_k e_k|.
//
// Full split, T = sum_k (hq he + hq le + lq he) on the tensor core, against the canonical float32 score C:
//   * x = h + l + r with h = fp16(x), l = fp16(x - h): |r| <= 2^-22 |x| in the fp16 normal range, <= 2^-25 below it
//     (x - h is exact in float32; the fp16 subnormal spacing is 2^-24).  Representation error of the products plus
//     the dropped lq * le term: <= 3 * 2^-22 * B + 2^-25 * 4 * (nq + ne)          (sum |q_k| <= 4 |q| for 16 components)
//   * accumulation inside the tensor core is not documented.  Model: the products are exact (22-bit mantissas) and
//     each of the 3 x 16 = 48 additions behaves no worse than a float32 addition of operands bounded by B:
//     <= 48 * 2^-23 * B
//   * C itself is a chain of 16 float32 FMAs: <= 16 * 2^-24 * B
//   |T - C| <= (3 * 2^-22 + 48 * 2^-23 + 16 * 2^-24) * B + 2^-23 * (nq + ne)  =  7.39e-6 * B + 1.2e-7 * (nq + ne)
//   (two unit heads, B = 2: 1.5e-5; measured maximum on config 2: 3.9e-7).
// hi*hi term alone, T1 = sum_k hq he: |hq he - q e| <= (2 * 2^-11 + 2^-22) |q e| per component, plus 16 additions and
// the chain of C:  |T1 - C| <= (2^-10 + 2^-22 + 16 * 2^-23 + 16 * 2^-24) * B + 2^-23 * (nq + ne)  =  9.796e-4 * B + ...
//   (B = 2: 1.96e-3; measured 1.1e-3).
// ---------------------------------------------------------------------------
// hi*hi with HALF-PRECISION accumulators (A16): on top of the above the result is delivered as an fp16 number, 2^-10
// relative at worst (truncation; rounding to nearest would be 2^-11), |result| <= B:  |T16 - C| <= 1.957e-3 * B + ...
//   (B = 2: 3.9e-3; measured with scripts/umma_f16acc_probe acc: see DESIGN.md 4.3).
constexpr float kFullSlackPerB = 7.39e-6f, kHiSlackPerB = 9.796e-4f, kA16SlackPerB = 1.957e-3f, kSlackPerNorm = 1.2e-7f;

// norms: largest squared row norm of the queries [0] and of the domains [1] (row_norm2_max_kernel)
// filter: 0 full split, 1 hi*hi term alone, 2 hi*hi with half-precision accumulators
__device__ __forceinline__ float score_slack(const unsigned *__restrict__ norms, int filter) {
    const float nq = sqrtf(__uint_as_float(norms[0])) * (1.0f + 1e-6f), ne = sqrtf(__uint_as_float(norms[1])) * (1.0f + 1e-6f);
    return (filter == 2 ? kA16SlackPerB : filter == 1 ? kHiSlackPerB : kFullSlackPerB) * nq * ne + kSlackPerNorm * (nq + ne);
}

// ---------------------------------------------------------------------------
// Compact split (validated on B200 in round 2: equal to the FFMA kernel on every route, tests/test_gpu_parity.py).
// With range_size 4 (the reference's default tile_size 1024, BASELINE config 4) only 7 of the 16 embedding
// dimensions are ever non-zero (3 tonal + 4 transient, fractal.py:154-208).  Eight live dimensions leave room for
// the hi AND the lo part in one K = 16 operand, so the full split needs two MMAs per stage instead of three:
//     queries   part 0 = [q_hi | q_lo]   part 1 = [q_hi | 0]
//     domains   part 0 = [e_hi | e_hi]   part 1 = [e_lo | 0]
//     part 0 . part 0 = hi*hi + lo*hi        part 1 . part 1 = hi*lo
// The parts sit where the hi / lo parts of pack_f16_tiles_kernel sit, so the hi*hi-only variant (part 0 alone)
// works unchanged and is a little more accurate (q . e_hi instead of q_hi . e_hi).
// ---------------------------------------------------------------------------
struct LivePerm { int n; int dim[8]; };

__global__ void live_dims_kernel(const float *__restrict__ x, long long n_rows, unsigned *__restrict__ mask) {
    unsigned m = 0;
    const long long total = n_rows * (ED / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
        const int k = (int)(i % (ED / 4)) * 4;
        m |= (v.x != 0.f ? 1u : 0u) << k | (v.y != 0.f ? 1u : 0u) << (k + 1) | (v.z != 0.f ? 1u : 0u) << (k + 2) |
             (v.w != 0.f ? 1u : 0u) << (k + 3);
    }
    m = __reduce_or_sync(kFull, m);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(mask, m);
}

// role 0: query tiles, role 1: domain tiles (see above); same tiling, swizzle and sample dealing as pack_f16_tiles_kernel
__global__ void __launch_bounds__(256)
pack_compact_tiles_kernel(const float *__restrict__ src, long long n_rows, long long n_tiles, uint4 *__restrict__ dst,
                          int row_stride, LivePerm perm, int role) {
    const long long total = n_tiles * (kDTile * 2);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const long long tile = g / (kDTile * 2);
        const int w = (int)(g - tile * (kDTile * 2));
        const int c = w / kDTile, r = w % kDTile;       // c: which 16-byte chunk of the 32-byte operand row
        long long row = tile * kDTile + r;
        if (row_stride > 1) {
            const long long in_stage = row & (kDStage - 1);
            row = ((row & ~(long long)(kDStage - 1)) + (in_stage & 63) * 4 + (in_stage >> 6)) * row_stride;
        }
        float xs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < n_rows) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < perm.n) xs[k] = __ldg(src + row * ED + perm.dim[k]);
        }
        __half2 hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half h0 = __float2half_rn(xs[2 * i]), h1 = __float2half_rn(xs[2 * i + 1]);
            hi[i] = __halves2half2(h0, h1);
            lo[i] = __halves2half2(__float2half_rn(xs[2 * i] - __half2float(h0)),
                                   __float2half_rn(xs[2 * i + 1] - __half2float(h1)));
        }
        const uint4 vhi = *reinterpret_cast<const uint4 *>(hi), vlo = *reinterpret_cast<const uint4 *>(lo);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        uint4 *t = dst + tile * (2 * kDTile * 2);
        const int slot = (r >> 3) * 16 + (r & 7) * 2 + (c ^ ((r >> 2) & 1));
        if (role == 0) {
            t[slot] = c == 0 ? vhi : vlo;                       // part 0 = [hi | lo]
            t[kDTile * 2 + slot] = c == 0 ? vhi : zero;         // part 1 = [hi | 0]
        } else {
            t[slot] = vhi;                                      // part 0 = [hi | hi]
            t[kDTile * 2 + slot] = c == 0 ? vlo : zero;         // part 1 = [lo | 0]
        }
    }
}

// ---------------------------------------------------------------------------
// Candidate bookkeeping.  A candidate is one 64-bit key
//     [ order-preserving bits of the score | 0xFFFFFFFF - domain index ]
// so "ranks before" (score descending, index ascending) is a plain unsigned
// compare.  Every query row owns kKeep keys in shared memory, kept sorted
// best-first; the row's threshold is the score of the last one.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_bits(float s) {
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
constexpr uint32_t kNegInfBits = 0x007FFFFFu;   // order_bits(-inf)
__device__ __forceinline__ unsigned long long empty_key(int slot) {
    return ((unsigned long long)kNegInfBits << 32) | (uint32_t)(kCap - 1 - slot);
}
__device__ __forceinline__ unsigned long long make_key(float s, int id) {
    return ((unsigned long long)order_bits(s) << 32) | (0xFFFFFFFFu - (uint32_t)id);
}

// Warp-cooperative merge of one 32-column chunk into a row's sorted list: lane l
// brings column l's key (`pass` says whether it beat the threshold).  Every list key
// and every passing key computes its rank in the union from ballots — no shared
// memory traffic and no barrier inside the loop over the passing columns — and is
// stored at that rank if it is below kKeep.  Keys are unique, so the ranks are a
// permutation.  Returns the new threshold (score of the last kept key).
__device__ __forceinline__ float merge_chunk(volatile unsigned long long *keys, unsigned long long kx, bool pass,
                                             int lane) {
    const bool has1 = lane + 32 < kKeep;
    const unsigned long long k0 = keys[lane];
    const unsigned long long k1 = has1 ? keys[lane + 32] : 0ull;
    int r0 = lane, r1 = lane + 32, rx = 0;
    unsigned pm = __ballot_sync(kFull, pass);
    while (pm) {
        const int j = __ffs(pm) - 1;
        pm &= pm - 1;
        const unsigned long long kj = __shfl_sync(kFull, kx, j);
        const bool g0 = k0 > kj, g1 = has1 && k1 > kj;
        r0 += g0 ? 0 : 1;                     // list keys behind column j's slide down one slot
        r1 += g1 ? 0 : 1;
        const unsigned b0 = __ballot_sync(kFull, g0), b1 = __ballot_sync(kFull, g1);
        const unsigned bx = __ballot_sync(kFull, pass && kx > kj);
        if (lane == j) rx = __popc(b0) + __popc(b1) + __popc(bx);
    }
    __syncwarp();
    if (r0 < kKeep) keys[r0] = k0;
    if (has1 && r1 < kKeep) keys[r1] = k1;
    if (pass && rx < kKeep) keys[rx] = kx;
    __syncwarp();
    return unorder_bits((uint32_t)(keys[kKeep - 1] >> 32));
}

// Largest of the 32 scores of one chunk (3-input max tree).
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
    float m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        m[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    m[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float a = max3(m[0], m[1], m[2]), b = max3(m[3], m[4], m[5]), c = max3(m[6], m[7], m[8]);
    return max3(max3(a, b, c), m[9], m[10]);
}

__device__ __forceinline__ void dump_chunk(uint32_t *dst, const uint32_t (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4 *>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}

// Rare path, once per stage and warp: rows whose threshold was beaten ("owners") are
// served one at a time.  The owner spills the chunks that hit to the warp's scratch
// line; then all 32 lanes take one column each and the chunk is merged into the row's
// list in one go.  `lists` are this warp's own 32 lists (its rows x its column half),
// so nothing here is shared with another warp.
__device__ __forceinline__ void absorb_stage(const uint32_t (&v)[kChunks][32], unsigned hits, float &tau,
                                             long long base, long long n_d, unsigned long long *lists,
                                             uint32_t *scratch, int lane) {
    unsigned owners = __ballot_sync(kFull, hits != 0);
    while (owners) {
        const int bl = __ffs(owners) - 1;
        owners &= owners - 1;
        const unsigned hb = __shfl_sync(kFull, hits, bl);
        if (lane == bl) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c)
                if (hb >> c & 1) dump_chunk(scratch + 32 * c, v[c]);
        }
        __syncwarp();
        float tb = __shfl_sync(kFull, tau, bl);
        unsigned long long *keys = lists + (size_t)bl * (2 * kCap);
        for (int c = 0; c < kChunks; ++c) {
            if (!(hb >> c & 1)) continue;
            const float x = __uint_as_float(scratch[32 * c + lane]);
            const long long id = base + 32 * c + lane;
            const bool pass = x > tb && id < n_d;
            if (__any_sync(kFull, pass)) tb = merge_chunk(keys, make_key(x, (int)id), pass, lane);
        }
        if (lane == bl) tau = tb;
        __syncwarp();
    }
}

// insert x into a descending register list (compare-exchange chain, no memory)
template <int L>
__device__ __forceinline__ void insert_desc(float (&t)[L], float x) {
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const float hi = fmaxf(t[i], x);
        x = fminf(t[i], x);
        t[i] = hi;
    }
}

// Visit the columns j of one 32-column chunk with v[j] >= thr, pruning with the same 3-input max
// tree chunk_max uses (nine-column groups, then triples, then single columns): a chunk that holds
// one short run of passing columns costs ~15 comparisons instead of 32.  f(j) is called per thread.
template <class F>
__device__ __forceinline__ void for_each_ge(const uint32_t (&v)[32], float thr, F f) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        float m[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            m[i] = max3(__uint_as_float(v[9 * g + 3 * i]), __uint_as_float(v[9 * g + 3 * i + 1]), __uint_as_float(v[9 * g + 3 * i + 2]));
        if (max3(m[0], m[1], m[2]) >= thr) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (m[i] >= thr) {
#pragma unroll
                    for (int e = 0; e < 3; ++e)
                        if (__uint_as_float(v[9 * g + 3 * i + e]) >= thr) f(9 * g + 3 * i + e);
                }
            }
        }
    }
    if (max3(__uint_as_float(v[27]), __uint_as_float(v[28]), __uint_as_float(v[29])) >= thr) {
#pragma unroll
        for (int e = 27; e < 30; ++e)
            if (__uint_as_float(v[e]) >= thr) f(e);
    }
    if (__uint_as_float(v[30]) >= thr) f(30);
    if (__uint_as_float(v[31]) >= thr) f(31);
}

// One kernel skeleton, three epilogues:
//   MODE_LISTS    exact threshold top-K with sorted per-row lists in shared memory (any table; the
//                 fallback of the fast path below and the path for small tables);
//   MODE_THETA    pass 1 of the fast path: scans a packed sample of the table (every 16th domain) and
//                 keeps, per row and in registers, the kTheta largest scores; theta = the kTheta-th of them, so about
//                 16 * kTheta domains of the whole table reach it (Gamma(kTheta) spread: fewer than 32
//                 with probability ~1e-8);
//   MODE_COLLECT  pass 2: visits every stage with theta fixed and appends the index of every column
//                 that reaches it to the row's buffer in global memory — no ordering work at all.
constexpr int kTraceFrom = 256, kTraceStages = 64;
#define FWAV_TRACE(slot, t)                                                                           \
    do {                                                                                              \
        if ((dbg & 64) && blockIdx.x == 0 && (t) >= kTraceFrom && (t) < kTraceFrom + kTraceStages)    \
            a.trace[((t) - kTraceFrom) * 8 + (slot)] = clock64();                                     \
    } while (0)

// HI: only the hi*hi term (one MMA per stage, half the operand bytes).  Its scores are off by up to
// score_slack(hi_only), which a FILTER can afford when the data leave that much room between the top_k-th score and the
// threshold (decided per launch from pass 1, verified per query by finalize_kernel); MODE_LISTS never uses it.
// CG: CTAs per tensor-core group.  2 = the pair described above (a domain tile is read once per 256 queries).
// 1 = every CTA on its own (M = 128, the whole 256-domain stage in its shared memory, all barriers local: no relay,
// no remote arrive, no multicast commit in the accumulator hand-over chain that bounds the streaming modes).
#ifndef FWAV_ALT_SETS
#define FWAV_ALT_SETS 1
#endif

template <int MODE, bool HI, int CG, bool COMPACT = false>
__global__ void __cluster_dims__(CG, 1, 1) __launch_bounds__(n_threads(MODE), 1) scan_kernel(const ScanArgs a) {
    static_assert(!(HI && COMPACT), "the hi*hi-only variant reads part 0 alone: no separate compact form");
    // hi*hi-only collect pass: the sixteen epilogue warps form two sets, one per TMEM buffer.  A set takes every
    // other stage (128 columns per warp), so while one set waits for its tcgen05.ld the other one is reducing;
    // with all warps on every stage they move in lockstep and the load time adds to the ALU time (collect pass of
    // config 2: 73.0 -> 68.7 ms).  A set hands its buffer back only after its second pair of loads, which costs
    // more than it gains when a stage takes three MMAs (threshold pass 7.8 -> 10.1 ms, full-split collect +3.5 %):
    // those keep all warps on every stage.
    constexpr bool kAlt = FWAV_ALT_SETS && MODE == MODE_COLLECT && HI;
    constexpr int kQGroup = kQTile * CG;                             // queries per tensor-core group
    constexpr uint32_t kStageBytes = stage_bytes(CG);                // B bytes per stage in this CTA's shared memory
    constexpr uint32_t kOffB = off_ring(MODE);
    constexpr uint32_t kLoOff = kStageBytes / 2;                     // hi parts first, lo parts behind them
    static_assert(!(HI && MODE == MODE_LISTS), "the exact list kernel needs the full split");
    constexpr uint32_t kOpBytes = HI ? kPartBytes : kTileBytes;      // bytes of a tile this launch reads
    constexpr int kEpi = epi_warps(MODE), kThreads = n_threads(MODE);
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_q = a.n_q, n_d = a.n_d;
    const int top_k = a.top_k, dbg = a.dbg;
    const uint8_t *__restrict__ active = a.active;
    uint32_t cta_rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const int pair_id = (int)(blockIdx.x / CG) / a.n_split, split = (int)(blockIdx.x / CG) % a.n_split;
    const long long pair_base = (long long)pair_id * kQGroup;              // first query of the group
    const long long q_base = pair_base + (long long)cta_rank * kQTile;    // first query of this CTA
    unsigned long long *rows = reinterpret_cast<unsigned long long *>(smem + kOffMode);
    const uint32_t bars = smem_u32(smem + kOffBars);
    // barrier slots (8 bytes each)
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tfull = bars + 16 * kStages,
                   bar_tempty = bar_tfull + 16, bar_a = bar_tempty + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffBars + kBarBytes - 8);

    // whole-pair early out (energy-pruned stretch).  Both CTAs look at all 256 rows so that they
    // take the same decision.
    {
        int any = 0;
        for (int i = threadIdx.x; i < kQGroup; i += kThreads) {
            const long long q = pair_base + i;
            if (q < n_q && (!active || active[q])) any = 1;
        }
        if (!__syncthreads_or(any)) {
            if (MODE == MODE_LISTS && a.n_split > 1) {
                // merge_parts_kernel writes the -1 rows of pruned queries itself
            } else if (MODE == MODE_LISTS) {
                for (int i = threadIdx.x; i < kQTile * top_k; i += kThreads) {
                    const long long q = q_base + i / top_k;
                    if (q < n_q) {
                        a.cand[q * top_k + i % top_k] = -1;
                        if (a.scores) a.scores[q * top_k + i % top_k] = -INFINITY;
                    }
                }
            } else {
                for (int i = threadIdx.x; i < kQTile; i += kThreads) {
                    const long long q = q_base + i;
                    if (q >= n_q) continue;
                    if (MODE == MODE_THETA) { a.theta[q] = INFINITY; a.theta_hi[q] = INFINITY; }
                    if (MODE == MODE_COLLECT) { for (int g = 0; g < 4; ++g) a.ccount[(q * a.n_split + split) * 4 + g] = 0; }
                }
            }
            return;
        }
    }

    if (MODE == MODE_LISTS)
        for (int i = threadIdx.x; i < kQTile * 2 * kCap; i += kThreads) rows[i] = empty_key(i % kCap);
    if (threadIdx.x == 0) {
        // the leader's "stage has landed" barriers collect its own copy and the peer's relay
        const uint32_t n_land = (CG == 2 && cta_rank == 0) ? 2u : 1u;
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, n_land); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, CG * kEpi / (kAlt ? 2 : 1)); }
        mbar_init(bar_a, n_land);
        mbar_init(bar_a + 8, 1);
        mbar_init(bar_a + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpi) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();           // barriers of both CTAs are initialised before anyone arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Full scans visit the stages starting at the pair's own rows and wrapping around: when the
    // queries are rows of the same table (the reference's aliasing) their best matches sit next to
    // them, so MODE_LISTS thresholds tighten within the first few stages; and neighbouring pairs
    // are always one stage apart, so every stage is read from HBM once and from L2 by the rest.
    // With n_split > 1 this pair only covers its share [s_lo, s_hi) of the stages.
    const int s_lo = (int)((long long)split * a.n_stages / a.n_split), s_hi = (int)((long long)(split + 1) * a.n_stages / a.n_split);
    const int n_visit = s_hi - s_lo;
    const int t_first = s_lo + (int)((pair_base / kDStage) % n_visit);
    constexpr uint32_t kStageStride = kStageBytes;

    if (warp == kEpi) {
        // ===== producer: bulk copies (TMA engine) of this CTA's query tile and its half of every stage =====
        // (the whole warp runs the loop, one elected lane issues: operands of the copy instructions stay in uniform
        // registers instead of going through an elect-and-broadcast loop each -- see collect_hi_kernel)
        {
            if (elect_one()) {
                mbar_expect_tx(bar_a, kOpBytes);
                bulk_g2s(smem_u32(smem + kOffA), a.q_tiles + ((long long)CG * pair_id + cta_rank) * (kTileBytes / 16), kOpBytes, bar_a);
            }
            __syncwarp();
            int tt = t_first;
            for (int t = 0; t < n_visit; ++t) {
                const int s = t & (kStages - 1);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const uint32_t dst = smem_u32(smem + kOffB + s * kStageStride);
                if (elect_one()) {
                    if (CG == 2) {
                        // this CTA's half of the stage: its hi part, and right behind it its lo part
                        mbar_expect_tx(bar_full + 8 * s, kOpBytes);
                        bulk_g2s(dst, a.e_tiles + (2ll * tt + cta_rank) * (kTileBytes / 16), kOpBytes, bar_full + 8 * s);
                    } else {
                        // both tiles of the stage: [hi 0 | hi 1] and, for the full split, [lo 0 | lo 1] behind them
                        const uint4 *t0 = a.e_tiles + (2ll * tt) * (kTileBytes / 16), *t1 = t0 + kTileBytes / 16;
                        mbar_expect_tx(bar_full + 8 * s, 2 * kOpBytes);
                        bulk_g2s(dst, t0, kPartBytes, bar_full + 8 * s);
                        bulk_g2s(dst + kPartBytes, t1, kPartBytes, bar_full + 8 * s);
                        if (!HI) {
                            bulk_g2s(dst + kLoOff, t0 + kPartBytes / 16, kPartBytes, bar_full + 8 * s);
                            bulk_g2s(dst + kLoOff + kPartBytes, t1 + kPartBytes / 16, kPartBytes, bar_full + 8 * s);
                        }
                    }
                }
                __syncwarp();
                if (++tt == s_hi) tt = s_lo;
            }
        }
    } else if (CG == 2 && cta_rank != 0 && warp > kEpi) {
        // ===== peer CTA: tell the leader when this CTA's operands have landed =====
        if (warp == kEpi + 1 && lane == 0) {
            mbar_wait(bar_a, 0);
            mbar_arrive_remote(bar_a, 0);
            for (int t = 0; t < n_visit; ++t) {
                const int s = t & (kStages - 1);
                mbar_wait(bar_full + 8 * s, (uint32_t)((t / kStages) & 1));
                mbar_arrive_remote(bar_full + 8 * s, 0);
            }
        }
    } else if (warp > kEpi) {
        // ===== leader CTA: two MMA issuers, one thread each, for the pair =====
        // One stage costs the issuing thread ~600 cycles of serial latency (two mbarrier waits,
        // three tcgen05.mma, two commits) against ~520 cycles of tensor-pipe time, so the stages
        // alternate between two threads: the first issuer warp owns the even ones (TMEM buffer 0), the second the odd
        // ones (buffer 1).  The buffers and shared-memory stages are disjoint and a commit covers
        // its own thread's MMAs, so the two instruction streams need no ordering between them.
        // (whole warp in the loop, one elected lane issues -- as the producer above)
        {
            mbar_wait(bar_a, 0);
            const uint32_t a_hi = smem_u32(smem + kOffA), a_lo = a_hi + kPartBytes;
            const uint64_t da_hi = smem_desc(a_hi), da_lo = smem_desc(a_lo);
            const int buf = __shfl_sync(kFull, warp - (kEpi + 1), 0);
            const uint32_t d = tmem_base + (uint32_t)(buf * kDStage);
            for (int t = buf; t < n_visit; t += 2) {
                const int s = t & (kStages - 1);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                const uint32_t tph = (uint32_t)((t >> 1) & 1);
                const uint32_t b_hi = smem_u32(smem + kOffB + s * kStageStride), b_lo = b_hi + kLoOff;
                const uint64_t db_hi = smem_desc(b_hi), db_lo = smem_desc(b_lo);
                mbar_wait(bar_full + 8 * s, ph);                               // landed long ago, as a rule
                if (lane == 0) FWAV_TRACE(0, t);
                if (!(dbg & 8)) mbar_wait(bar_tempty + 8 * buf, tph ^ 1);    // dbg 8: free-running MMA (profiling)
                if (lane == 0) FWAV_TRACE(1, t);
                tc_fence_after();
                if (elect_one()) {
                    // small cross terms first, the hi*hi term last; one K=16 instruction each
                    if (HI) {
                        umma_f16<CG>(d, da_hi, db_hi, 0);
                    } else if (COMPACT) {
                        umma_f16<CG>(d, da_lo, db_lo, 0);      // part 1 . part 1 = hi*lo
                        umma_f16<CG>(d, da_hi, db_hi, 1);      // part 0 . part 0 = hi*hi + lo*hi
                    } else {
                        umma_f16<CG>(d, da_hi, db_lo, 0);
                        umma_f16<CG>(d, da_lo, db_hi, 1);
                        umma_f16<CG>(d, da_hi, db_hi, 1);
                    }
                    umma_commit<CG>(bar_tfull + 8 * buf);      // accumulators ready for the epilogue warps of both CTAs
                    umma_commit<CG>(bar_empty + 8 * s);        // stage free (both CTAs) once these MMAs have read it
                }
                __syncwarp();
                if (lane == 0) FWAV_TRACE(2, t);
            }
            if (dbg & 8) {   // free-running profiling mode: drain the tensor pipe before leaving
                if (elect_one()) umma_commit<CG>(bar_a + 8 + 8 * buf);
                __syncwarp();
                mbar_wait(bar_a + 8 + 8 * buf, 0);
            }
        }
    } else if (warp < kEpi) {
        // ===== epilogue: one query row per thread; a warp owns 32 rows x (256 / (kEpi / 4)) columns of every stage =====
        const int quad = warp & 3, half = warp >> 2;      // half: which column group (2 of 128 or 4 of 64)
        const int row0 = quad * 32;                       // first row of this warp inside the CTA tile
        const long long q = q_base + row0 + lane;
        const bool live = q < n_q && (!active || active[q]) && !(dbg & 4);
        constexpr int kCols = kDStage / (kEpi / 4);       // columns per warp and stage
        // two sets: `half` = set + 2 * (which 128 columns); the set's buffer never changes
        const int set = half & 1, colhalf = half >> 1;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                (uint32_t)(kAlt ? set * kDStage + colhalf * 128 : half * kCols);
        // mode state
        float tau = live ? -INFINITY : INFINITY;                    // LISTS: running threshold; +inf: never a candidate
        unsigned long long *lists = rows + ((size_t)row0 * 2 + half) * kCap;     // row r of the warp: lists + r * 2 * kCap
        uint32_t *scratch = reinterpret_cast<uint32_t *>(smem + kOffScratch) + warp * 128;
        float t8[kThetaPart];                                       // THETA: largest scores of this column group, descending
#pragma unroll
        for (int i = 0; i < kThetaPart; ++i) t8[i] = live ? -INFINITY : INFINITY;
        int cnt = 0;                                                // COLLECT
        int32_t *cbuf = nullptr;
        if (MODE == MODE_COLLECT) {
            tau = (q < n_q && !(dbg & 4)) ? a.theta[q] : INFINITY;  // +inf for pruned rows (written by pass 1)
            cbuf = a.cbuf + (((q < n_q ? q : 0) * a.n_split + split) * 4 + half) * (long long)a.cap;
        }
        int tt = t_first;
        if (MODE == MODE_LISTS) {
            for (int t = 0; t < ((dbg & 8) ? 0 : n_visit); ++t) {
                const int buf = t & 1;
                const uint32_t tph = (uint32_t)((t >> 1) & 1);
                mbar_wait(bar_tfull + 8 * buf, tph);
                tc_fence_after();
                const uint32_t ta = t_lane + (uint32_t)(buf * kDStage);
                uint32_t v[kChunks][32];
                float m0 = -INFINITY, m1 = -INFINITY;
                if (!(dbg & 2)) {
                    tmem_ld32(ta, v[0]);
                    tmem_ld32(ta + 32, v[1]);
                    tmem_wait_ld2(v[0], v[1]);
                    tmem_ld32(ta + 64, v[2]);        // in flight while the first two chunks are reduced
                    tmem_ld32(ta + 96, v[3]);
                    m0 = chunk_max(v[0]);
                    m1 = chunk_max(v[1]);
                    tmem_wait_ld2(v[2], v[3]);
                }
                // the whole stage sits in registers: hand the TMEM buffer back before looking at the rest
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_remote(bar_tempty + 8 * buf, 0); else mbar_arrive_local(bar_tempty + 8 * buf); }
                const long long base = (long long)tt * kDStage + half * 128;
                if (++tt == s_hi) tt = s_lo;
                if (dbg & 3) continue;
                const float m2 = chunk_max(v[2]), m3 = chunk_max(v[3]);
                const unsigned hits = (m0 > tau ? 1u : 0u) | (m1 > tau ? 2u : 0u) | (m2 > tau ? 4u : 0u) | (m3 > tau ? 8u : 0u);
                if (__any_sync(kFull, hits != 0)) absorb_stage(v, hits, tau, base, n_d, lists, scratch, lane);
            }
        } else if (kAlt && !(dbg & 8)) {
            uint32_t x0[32], x1[32];
            int it = 0;
            // one 32-column chunk with its maximum m: THETA keeps the group's best scores, COLLECT appends indices
            auto look = [&](const uint32_t (&x)[32], float m, int col) {
                if (MODE == MODE_THETA) {
                    if (it < kThetaWarm / 2) {
                        if (m > t8[kThetaPart - 1]) insert_desc(t8, m);      // warm-up: chunk maxima only (see below)
                    } else if (m > t8[kThetaPart - 1]) {
                        for_each_ge(x, nextafterf(t8[kThetaPart - 1], INFINITY), [&](int j) {
                            const float v = __uint_as_float(x[j]);
                            if (v > t8[kThetaPart - 1]) insert_desc(t8, v);
                        });
                    }
                } else if (m >= tau) {
                    for_each_ge(x, tau, [&](int j) {
                        if (cnt < a.cap) cbuf[cnt] = col + j;
                        ++cnt;
                    });
                }
            };
            tt = t_first + set;
            if (tt >= s_hi) tt -= n_visit;
            const uint32_t bar_f = bar_tfull + 8 * set, bar_e = bar_tempty + 8 * set;
            for (int t = set; t < n_visit; t += 2, ++it) {
                mbar_wait_hot(bar_f, (uint32_t)(it & 1));
                tc_fence_after();
                tmem_ld32(t_lane, x0);
                tmem_ld32(t_lane + 32, x1);
                tmem_wait_ld2(x0, x1);
                const int col0 = tt * kDStage + colhalf * 128;
                tt += 2;
                if (tt >= s_hi) tt -= n_visit;
                {
                    const float ma = chunk_max(x0), mb = chunk_max(x1);
                    look(x0, ma, col0);
                    look(x1, mb, col0 + 32);
                }
                tmem_ld32(t_lane + 64, x0);
                tmem_ld32(t_lane + 96, x1);
                tmem_wait_ld2(x0, x1);
                // the warp's share of the buffer has been read: hand it back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_remote(bar_e, 0); else mbar_arrive_local(bar_e); }
                {
                    const float ma = chunk_max(x0), mb = chunk_max(x1);
                    look(x0, ma, col0 + 64);
                    look(x1, mb, col0 + 96);
                }
            }
        } else if (!(dbg & 8)) {
            // Streaming epilogue: 64 columns per warp and stage, four warps per scheduler interleave their chains.
            for (int t = 0; t < n_visit; ++t) {
                const int buf = t & 1;
                mbar_wait_hot(bar_tfull + 8 * buf, (uint32_t)((t >> 1) & 1));
                tc_fence_after();
                const uint32_t ta = t_lane + (uint32_t)(buf * kDStage);
                uint32_t x0[32], x1[32];
                tmem_ld32(ta, x0);
                tmem_ld32(ta + 32, x1);
                tmem_wait_ld2(x0, x1);
                // this warp's share of the buffer is in registers: hand it back before looking at it (the hand-over
                // chain MMA -> commit -> load -> arrive -> MMA is what bounds the scan, so nothing else goes in it)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_remote(bar_tempty + 8 * buf, 0); else mbar_arrive_local(bar_tempty + 8 * buf); }
                const float ma = chunk_max(x0);
                const int col0 = tt * kDStage + half * kCols;
                if (++tt == s_hi) tt = s_lo;
                const float mb = chunk_max(x1);
                if (MODE == MODE_THETA) {
                    if (t < kThetaWarm) {
                        // warm-up: while the thresholds are still loose nearly every chunk passes; taking only the
                        // chunk maxima (each of them a real score) tightens them at one insertion per chunk.  What
                        // the other columns of those chunks might have contributed can only lower theta.
                        if (ma > t8[kThetaPart - 1]) insert_desc(t8, ma);
                        if (mb > t8[kThetaPart - 1]) insert_desc(t8, mb);
                    } else {
                        // every sampled score that beats this column group's current kThetaPart-th best
                        if (ma > t8[kThetaPart - 1])
                            for_each_ge(x0, nextafterf(t8[kThetaPart - 1], INFINITY), [&](int j) {
                                const float x = __uint_as_float(x0[j]);
                                if (x > t8[kThetaPart - 1]) insert_desc(t8, x);
                            });
                        if (mb > t8[kThetaPart - 1])
                            for_each_ge(x1, nextafterf(t8[kThetaPart - 1], INFINITY), [&](int j) {
                                const float x = __uint_as_float(x1[j]);
                                if (x > t8[kThetaPart - 1]) insert_desc(t8, x);
                            });
                    }
                } else {
                    if (ma >= tau)
                        for_each_ge(x0, tau, [&](int j) {
                            if (cnt < a.cap) cbuf[cnt] = col0 + j;
                            ++cnt;
                        });
                    if (mb >= tau)
                        for_each_ge(x1, tau, [&](int j) {
                            if (cnt < a.cap) cbuf[cnt] = col0 + 32 + j;
                            ++cnt;
                        });
                }
            }
        }
        if (MODE == MODE_COLLECT) {
            if (q < n_q) a.ccount[(q * a.n_split + split) * 4 + half] = cnt;
        } else if (MODE == MODE_THETA) {
            // theta of a row = kTheta-th largest sampled score over the four column groups.  A group keeps
            // its own kThetaPart best, so if more than that many of the row's best sit in one group the
            // merged value comes out lower than the true one: more candidates, never fewer.
            float *th = reinterpret_cast<float *>(smem + kOffMode);
            if (half != 0) {
#pragma unroll
                for (int i = 0; i < kThetaPart; ++i) th[((row0 + lane) * 3 + half - 1) * kThetaPart + i] = t8[i];
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (half == 0 && q < n_q) {
                constexpr int kMerged = 4 * kThetaPart;          // all four groups' lists, merged and sorted
                float tm[kMerged];
#pragma unroll
                for (int i = 0; i < kMerged; ++i) tm[i] = i < kThetaPart ? t8[i] : (live ? -INFINITY : INFINITY);
                for (int i = 0; i < 3 * kThetaPart; ++i) {
                    const float x = th[(row0 + lane) * 3 * kThetaPart + i];
                    if (x > tm[kMerged - 1]) insert_desc(tm, x);
                }
                // theta = the theta_rank-th best sampled score (16th for top_k <= 32: ~256 domains reach it;
                // 24th for top_k <= 64: ~384).  +inf for pruned rows, -inf if the sample was too small.
                float tsel = tm[0];
#pragma unroll
                for (int i = 1; i < kMerged; ++i) tsel = (i == a.theta_rank - 1) ? tm[i] : tsel;
                a.theta[q] = tsel;
                int hi_rank = a.hi_rank - 1;
                hi_rank = hi_rank < 0 ? 0 : hi_rank > kTheta - 1 ? kTheta - 1 : hi_rank;
                float th = tm[0];
#pragma unroll
                for (int i = 1; i < kTheta; ++i) th = i == hi_rank ? tm[i] : th;
                a.theta_hi[q] = hi_rank == 0 ? tm[0] : th;
            }
        } else {
            // both warps of a quadrant are done with their lists before the rows are finalised
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (a.n_split > 1) {
                // partial result of this share of the table: the raw keys, merged by merge_parts_kernel
                for (int r = half * 16; r < half * 16 + 16; ++r) {
                    const long long qq = q_base + row0 + r;
                    if (qq >= n_q) break;
                    const unsigned long long *keys = rows + (size_t)(row0 + r) * 2 * kCap;
                    unsigned long long *dst = a.parts + (qq * a.n_split + split) * (2 * kCap);
                    for (int i = lane; i < 2 * kCap; i += 32) dst[i] = keys[i];
                }
            } else {
            // ---- exact float32 re-score of the 2 x kKeep kept candidates of a row and best-first write-out ----
            constexpr int kPerLane = 2 * kKeep / 32;
            static_assert(2 * kKeep % 32 == 0, "final merge handles whole warps of keys");
            for (int r = half * 16; r < half * 16 + 16; ++r) {
                const long long qq = q_base + row0 + r;
                if (qq >= n_q) break;
                unsigned long long *keys = rows + (size_t)(row0 + r) * 2 * kCap;      // both column halves, contiguous
                const float *qv = a.Q + qq * ED;
                // what an EVICTED candidate of either half can have scored at most on the tensor cores: the half's
                // 48th kept key (-inf while that list is not full: nothing was evicted from it)
                const float evict_max = fmaxf(unorder_bits((uint32_t)(keys[kKeep - 1] >> 32)),
                                              unorder_bits((uint32_t)(keys[kCap + kKeep - 1] >> 32)));
                __syncwarp();
                // canonical score (ascending-k float32 FMA chain) of every kept candidate
                unsigned long long k[kPerLane];
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) {
                    k[i] = keys[lane + 32 * i];
                    if ((uint32_t)(k[i] >> 32) != kNegInfBits) {
                        const int id = (int)(0xFFFFFFFFu - (uint32_t)k[i]);
                        k[i] = make_key(fwm::score_chain(qv, a.E + (long long)id * ED, ED), id);
                    } else {
                        k[i] = (unsigned long long)kNegInfBits << 32 | (uint32_t)(2 * kCap - 1 - (lane + 32 * i));   // distinct empties
                    }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) keys[lane + 32 * i] = k[i];
                __syncwarp();
                int rk[kPerLane];
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) rk[i] = 0;
                for (int o = 0; o < 2 * kKeep; ++o) {
                    const unsigned long long ko = keys[o];
#pragma unroll
                    for (int i = 0; i < kPerLane; ++i) rk[i] += ko > k[i] ? 1 : 0;
                }
                float kth = INFINITY;        // canonical score of the last returned candidate (-inf: fewer than top_k exist)
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) {
                    if (rk[i] < top_k) {
                        const bool has = (uint32_t)(k[i] >> 32) != kNegInfBits;
                        a.cand[qq * top_k + rk[i]] = has ? (int)(0xFFFFFFFFu - (uint32_t)k[i]) : -1;
                        if (a.scores) a.scores[qq * top_k + rk[i]] = has ? unorder_bits((uint32_t)(k[i] >> 32)) : -INFINITY;
                        if (rk[i] == top_k - 1) kth = has ? unorder_bits((uint32_t)(k[i] >> 32)) : -INFINITY;
                    }
                }
                kth = fminf(kth, __shfl_xor_sync(kFull, kth, 16));
                kth = fminf(kth, __shfl_xor_sync(kFull, kth, 8));
                kth = fminf(kth, __shfl_xor_sync(kFull, kth, 4));
                kth = fminf(kth, __shfl_xor_sync(kFull, kth, 2));
                kth = fminf(kth, __shfl_xor_sync(kFull, kth, 1));
                // The lists are kept by TENSOR-CORE score.  An evicted candidate scored at most evict_max there, hence
                // at most evict_max + slack canonically: unless the top_k-th canonical score beats that, a true member
                // of the top_k may have been evicted (more than 16 candidates crowding the boundary within the filter's
                // error) and the row goes to the FFMA kernel.
                if (lane == 0 && a.lfail_list && evict_max != -INFINITY && !(kth > evict_max + score_slack(a.norms, 0)))
                    a.lfail_list[atomicAdd(a.lfail_count, 1)] = (int)qq;
                __syncwarp();
            }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the other may still signal it
    if (warp == kEpi) {
        tc_fence_after();
        if (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------
// collect_hi_kernel (round 2): the hi*hi-only collect pass -- the dominant kernel -- with HALF-PRECISION ACCUMULATORS
// and FOUR issuing threads.
//
// (1) Every score of the pass goes through a 3-input max on the half-rate ALU pipe.  Half-precision accumulators
// (kind::f16 with D=F16: one value per 32-bit TMEM column) come out of TMEM two per register with
// tcgen05.ld ...pack::16b, and the DPX packed maximum (SASS VIMNMX3.S16x2; fp16 bit patterns of non-negative values
// order like signed 16-bit integers) handles both halves in one instruction: half the registers and half the ALU work
// per score, and a warp's whole 128-column share of a stage fits ONE round of loads, so the accumulator goes back
// before anything is reduced.  The price is the filter's error -- the result is delivered as an fp16 number -- which
// score_slack accounts for (measured: scripts/umma_f16acc_probe.cu acc); batches whose queries lack the room for it
// keep float32 accumulators (scan_kernel).
// (2) Measured with the epilogue switched off (debug build, FWAV_UMMA_DEBUG=8): two issuer threads put one M128 N256
// K16 instruction per ~300 cycles on the tensor pipe (38 ms for config 2) although the instruction takes 171: one
// issue is a serial chain in its thread (the "operands landed" wait, ~90 cycles even when long complete; the
// "accumulator free" wait; the instruction, which blocks its thread until the pipe takes it; two commits).  Here four
// threads issue, round-robin over the stages (thread i: stages t = i (mod 4), accumulator t & 1), and each is also the
// producer of its own stages' operands (bulk copies one own stage ahead), so the CTA stays at 20 warps -- five per
// scheduler, what 96 registers allow.  Two threads take turns on one accumulator; a parity wait of the one that runs
// ahead would pass on the phase BEFORE the one it needs, so every (accumulator, thread) pair has its own "free" barrier
// and the epilogue arrives on the one of the thread that issues the NEXT use.
// Epilogue layout, inputs, outputs and candidate buffers as scan_kernel<MODE_COLLECT, true, 1> (set = accumulator,
// 128 columns per warp).  Config 2 on B200: 68.4 -> 60-62 ms (profiles/r02_collect_acc16_timings.txt).
// ---------------------------------------------------------------------------
constexpr int kHiIss = 4;
constexpr int kHiThreads = (16 + kHiIss) * 32;
#ifdef FWAV_DEBUG_KNOBS
// clock64 stamps of one CTA (FWAV_UMMA_DEBUG bit 64; bits 16-23: first stage / 64, bits 24-30: CTA / 32): issuer 0 loop top, 1 operands landed,
// 2 accumulator free, 3 issued and committed | epilogue (quadrant 0) 4 accumulator full, 5 loaded (column half 0),
// 6 reduced, 7 loaded (column half 1)
#define FWAV_HI_TRACE(slot, t)                                                                               \
    do {                                                                                                     \
        if ((a.dbg & 64) && (int)blockIdx.x == ((a.dbg >> 24) & 0x7f) * 32 && (t) >= ((a.dbg >> 16) & 0xff) * 64 &&   \
            (t) < ((a.dbg >> 16) & 0xff) * 64 + kTraceStages)                                                \
            a.trace[((t) - ((a.dbg >> 16) & 0xff) * 64) * 8 + (slot)] = clock64();                           \
    } while (0)
#else
#define FWAV_HI_TRACE(slot, t) do { } while (0)
#endif
// stamps inside the epilogue warps cost them registers (the loop runs at the 96-register limit) and slow the kernel
// down by a third even when switched off at run time: compiled only with -DFWAV_TRACE_EPILOGUE (bit 128 switches them on)
#if defined(FWAV_DEBUG_KNOBS) && defined(FWAV_TRACE_EPILOGUE)
#define FWAV_HI_TRACE_EPI(cond, slot, t) do { if ((a.dbg & 128) && (cond)) FWAV_HI_TRACE(slot, t); } while (0)
#else
#define FWAV_HI_TRACE_EPI(cond, slot, t) do { } while (0)
#endif
constexpr uint32_t kHiOffBars = kPartBytes;                       // after the hi part of the query tile
constexpr uint32_t kHiOffRing = 8192;                             // 1024-aligned
constexpr uint32_t kHiStageBytes = 2 * kPartBytes;                // the two hi tiles of a 256-domain stage
constexpr uint32_t kHiOffScratch = kHiOffRing + kStages * kHiStageBytes;   // per epilogue warp: a 32 x 33-word transpose buffer
constexpr uint32_t kHiScratchWords = 32 * 33;
constexpr uint32_t kHiSmem = kHiOffScratch + 16 * kHiScratchWords * 4;

// THETA = true: the same skeleton as pass 1 of the fast path (thresholds from the sampled table, a.theta / a.theta_hi
// out).  No indices, no hits: a lane keeps, per parity of the column (the two halves of a packed register), the three
// largest CHUNK MAXIMA of its (accumulator, column half) group -- a packed insertion of six 2-input min / max
// instructions per 64-column chunk, unconditional -- and at the end the 4 x 6 values of a row are merged.  The r-th
// largest of such maxima can only be at or below the r-th largest sampled score (more candidates, never fewer); the
// sample tiles are dealt so that neighbouring samples of one similarity peak land in different (group, parity)
// streams (pack_f16_tiles_kernel).  Scores are fp16 here: a threshold is an estimate, and no proof depends on how it
// was obtained.
template <bool THETA>
__global__ void __launch_bounds__(kHiThreads, 1) collect_hi_kernel(const ScanArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_q = a.n_q;
    const uint8_t *__restrict__ active = a.active;
    const int group_id = (int)blockIdx.x / a.n_split, split = (int)blockIdx.x % a.n_split;
    const long long q_base = (long long)group_id * kQTile;
    const uint32_t bars = smem_u32(smem + kHiOffBars);
    // barrier slots (8 bytes each): full[8] empty[8] tfull[2] tfree[2][2] a
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tfull = bars + 16 * kStages,
                   bar_tfree = bar_tfull + 16, bar_a = bar_tfree + 32;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kHiOffBars + 16 * kStages + 64);

    {   // energy-pruned stretch: nothing to scan
        int any = 0;
        for (int i = threadIdx.x; i < kQTile; i += kHiThreads) {
            const long long q = q_base + i;
            if (q < n_q && (!active || active[q])) any = 1;
        }
        if (!__syncthreads_or(any)) {
            for (int i = threadIdx.x; i < kQTile; i += kHiThreads) {
                const long long q = q_base + i;
                if (q < n_q) {
                    if (THETA) { a.theta[q] = INFINITY; a.theta_hi[q] = INFINITY; }
                    else for (int g = 0; g < 4; ++g) a.ccount[(q * a.n_split + split) * 4 + g] = 0;
                }
            }
            return;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) mbar_init(bar_tfull + 8 * b, 1);
        for (int b = 0; b < 4; ++b) mbar_init(bar_tfree + 8 * b, 8);          // the eight epilogue warps of a set
        mbar_init(bar_a, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    int *front_slot = reinterpret_cast<int *>(tmem_slot + 1);
    if (threadIdx.x == 32) *front_slot = a.front ? *reinterpret_cast<volatile int *>(a.front + split) : -1;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int s_lo = (int)((long long)split * a.n_stages / a.n_split), s_hi = (int)((long long)(split + 1) * a.n_stages / a.n_split);
    const int n_visit = s_hi - s_lo;
    // any starting stage is valid (every CTA visits all n_visit stages of its split, wrapping; the order in which a
    // query's candidates are collected does not matter to finalize_kernel): start where the other CTAs are
    const int front0 = *front_slot;
    const int t_first = s_lo + (front0 >= 0 ? front0 % n_visit : (int)((q_base / kDStage) % n_visit));

    if (warp >= 16) {
        // ===== issuer + producer warps: warp i owns the stages t = i (mod 4) and their ring slots i, i + 4 =====
        // The whole warp runs the loop and waits; one elected lane issues.  (With a single lane inside a divergent
        // branch the compiler cannot know that descriptors and barrier addresses are warp-uniform and wraps every
        // tcgen05 / bulk-copy instruction in an elect-and-broadcast loop -- ~200 cycles per stage on the chain that
        // bounds this kernel.)
        const int i = __shfl_sync(kFull, warp - 16, 0), b = i & 1, j = i >> 1;
        auto load_stage = [&](int t) {             // the two hi tiles of the t-th stage this CTA visits
            int tt = t_first + t;
            if (tt >= s_hi) tt -= n_visit;
            const int s = t & (kStages - 1);
            const uint32_t dst = smem_u32(smem + kHiOffRing + s * kHiStageBytes);
            const uint4 *t0 = a.e_tiles + (2ll * tt) * (kTileBytes / 16), *t1 = t0 + kTileBytes / 16;
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * s, 2 * kPartBytes);
                bulk_g2s(dst, t0, kPartBytes, bar_full + 8 * s);
                bulk_g2s(dst + kPartBytes, t1, kPartBytes, bar_full + 8 * s);
            }
        };
        if (i == 0 && elect_one()) {
            mbar_expect_tx(bar_a, kPartBytes);
            bulk_g2s(smem_u32(smem + kOffA), a.q_tiles + (long long)group_id * (kTileBytes / 16), kPartBytes, bar_a);
        }
        __syncwarp();
        if (i < n_visit) load_stage(i);
        mbar_wait(bar_a, 0);
        const uint64_t da_hi = smem_desc(smem_u32(smem + kOffA));
        const uint32_t d = tmem_base + (uint32_t)(b * kDStage);
        int k = 0;
        for (int t = i; t < n_visit; t += kHiIss, ++k) {
            if (lane == 0) FWAV_HI_TRACE(0, t);
            if (i == 0 && (k & 7) == 0 && a.front && elect_one()) {       // every 32nd stage: where this CTA is
                int rel = t_first - s_lo + t;
                if (rel >= n_visit) rel -= n_visit;
                *reinterpret_cast<volatile int *>(a.front + split) = rel;
            }
            const int tn = t + kHiIss;
            if (tn < n_visit) {                    // next own stage: its slot was freed by this warp's own MMA of t - 4
                mbar_wait(bar_empty + 8 * (tn & (kStages - 1)), (uint32_t)(((tn / kStages) & 1) ^ 1));
                if (lane == 0 && !(a.dbg & 128)) FWAV_HI_TRACE(4, t);
                load_stage(tn);
                if (lane == 0 && !(a.dbg & 128)) FWAV_HI_TRACE(5, t);
            }
            const int s = t & (kStages - 1);
            const uint64_t db_hi = smem_desc(smem_u32(smem + kHiOffRing + s * kHiStageBytes));
            mbar_wait(bar_full + 8 * s, (uint32_t)((t / kStages) & 1));
            if (lane == 0) FWAV_HI_TRACE(1, t);
            // use u = t >> 1 = 2 k + j of accumulator b: its previous use has been read (nothing to wait for at u = 0)
            if (j == 1 || k > 0) mbar_wait(bar_tfree + 8 * (2 * b + j), (uint32_t)((j == 1 ? k : k - 1) & 1));
            if (lane == 0) FWAV_HI_TRACE(2, t);
            tc_fence_after();
            if (elect_one()) {
                umma_f16<1, true>(d, da_hi, db_hi, 0);
                umma_commit<1>(bar_tfull + 8 * b);
                umma_commit<1>(bar_empty + 8 * s);
            }
            __syncwarp();
            if (lane == 0) FWAV_HI_TRACE(3, t);
        }
    } else {
        // ===== epilogue: one query row per thread; warp = (TMEM lane quadrant, set = accumulator, column half) =====
        const int quad = warp & 3, half = warp >> 2, set = half & 1, colhalf = half >> 1;
        const long long q = q_base + quad * 32 + lane;
#ifdef FWAV_DEBUG_KNOBS
        const float tau = (!THETA && q < n_q && !(a.dbg & 256)) ? a.theta[q] : INFINITY;      // dbg 256: no hits in this kernel (profiling)
#else
        const float tau = (!THETA && q < n_q) ? a.theta[q] : INFINITY;      // +inf for pruned rows (written by pass 1)
#endif
        // THETA: the three largest chunk maxima per parity, descending (signed 16-bit order, see pmax3)
        unsigned top0 = 0x80008000u, top1 = 0x80008000u, top2 = 0x80008000u;
        auto keep = [&](unsigned v) {
            unsigned h = __vmaxs2(top0, v); v = __vmins2(top0, v); top0 = h;
            h = __vmaxs2(top1, v); v = __vmins2(top1, v); top1 = h;
            top2 = __vmaxs2(top2, v);
        };
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * kDStage + colhalf * 128);
        const uint32_t bar_f = bar_tfull + 8 * set;
        int cnt = 0;
        int tt = t_first + set;
        if (tt >= s_hi) tt -= n_visit;
        uint32_t x0[32], x1[32];
        // the row's threshold in the accumulators' own format, rounded DOWN (nothing that reaches theta is lost),
        // minus one unit so that "exceeds" means "reaches"; +inf (pruned rows) stays out of reach
        const float tpos = tau > 0.0f ? tau : 0.0f;
        const unsigned t1x2 = (unsigned)(((int)__half_as_ushort(__float2half_rd(tpos)) - 1) & 0xffff) * 0x10001u;
        int it = 0;
        // Hits.  A row's hits sit in registers only its own lane can index, and only with constant indices: walking
        // them lane by lane costs a tree of divergent branches per hit (and 30 KB of unrolled code), which stalls the
        // whole set where hits are dense (a query's own neighbourhood).  Instead the warp turns a chunk with hits by
        // 90 degrees through shared memory: every lane stores its 32 registers (conflict-free, 33-word rows), then for
        // each row with a hit lane j reads that row's register j -- one load -- and all 64 columns of the row are
        // compared at once; ballots give every hit its place in the row's buffer, so the indices of a row leave as
        // neighbouring stores.  No divergence, no data-dependent register index.
        uint32_t *scr = reinterpret_cast<uint32_t *>(smem + kHiOffScratch) + warp * kHiScratchWords;
        int32_t *cbuf_warp = a.cbuf + (((q_base + quad * 32) * a.n_split + split) * 4 + half) * (long long)a.cap;   // row 0 of the warp
        const long long cbuf_row = (long long)a.n_split * 4 * a.cap;
        const unsigned lane_lt = (1u << lane) - 1u;
        auto extract = [&](const uint32_t (&x)[32], unsigned m, int colbase) {
            unsigned rows_hit = __ballot_sync(kFull, p_beats(m, t1x2));
            if (rows_hit == 0) return;
#pragma unroll
            for (int j = 0; j < 32; ++j) scr[j * 33 + lane] = x[j];
            __syncwarp();
            do {
                const int r = __ffs(rows_hit) - 1;
                rows_hit &= rows_hit - 1;
                const unsigned v = scr[lane * 33 + r];                  // register `lane` of row r: its columns 2 lane, 2 lane + 1
                const unsigned tr = __shfl_sync(kFull, t1x2, r);
                const unsigned gt = __vcmpgts2(v, tr);                  // 0xffff per half that exceeds the row's threshold
                const unsigned b_lo = __ballot_sync(kFull, (gt & 0xffffu) != 0), b_hi = __ballot_sync(kFull, (gt >> 16) != 0);
                const int cnt_r = __shfl_sync(kFull, cnt, r);
                int32_t *cb = cbuf_warp + r * cbuf_row;
                const int p_lo = cnt_r + __popc(b_lo & lane_lt), p_hi = cnt_r + __popc(b_lo) + __popc(b_hi & lane_lt);
                if ((gt & 0xffffu) && p_lo < a.cap) cb[p_lo] = colbase + 2 * lane;
                if ((gt >> 16) && p_hi < a.cap) cb[p_hi] = colbase + 2 * lane + 1;
                if (lane == r) cnt = cnt_r + __popc(b_lo) + __popc(b_hi);
            } while (rows_hit);
            __syncwarp();
        };
        uint32_t bar_back = bar_tfree + 8 * (2 * set + 1);      // the issuers of an accumulator take turns: 2 set + ((it + 1) & 1)
        for (int t = set; t < n_visit; t += 2, ++it) {
            mbar_wait_hot(bar_f, (uint32_t)(it & 1));
            FWAV_HI_TRACE_EPI(quad == 0 && colhalf == 0 && lane == 0, 4, t);
            tc_fence_after();
            tmem_ld32_pack16(t_lane, x0);            // this warp's 128 columns of the stage, in one round
            tmem_ld32_pack16(t_lane + 64, x1);
            tmem_wait_ld2(x0, x1);
            FWAV_HI_TRACE_EPI(quad == 0 && lane == 0, colhalf ? 7 : 5, t);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_local(bar_back);     // to the warp that issues the next use
            bar_back ^= 8u;
            // The hand-back is on the chain that bounds the kernel and nothing in this warp depends on it, so the
            // instruction scheduler sinks it below the first levels of the max tree (~30 instructions).  A wait that
            // always passes (the query tile's barrier completed its only phase at the start) is a loop the scheduler
            // does not move code across: the hand-back stays up here.
            mbar_wait_hot(bar_a, 0);
            const int col0 = tt * kDStage + colhalf * 128;
            tt += 2;
            if (tt >= s_hi) tt -= n_visit;
            const unsigned m0 = chunk_max_p(x0), m1 = chunk_max_p(x1);
            if (THETA) {
                keep(m0);
                keep(m1);
            } else if (__any_sync(kFull, p_beats(pmax3(m0, m1, m1), t1x2))) {
                extract(x0, m0, col0);
                extract(x1, m1, col0 + 64);
            }
            FWAV_HI_TRACE_EPI(quad == 0 && colhalf == 0 && lane == 0, 6, t);
        }
        if (!THETA) {
            if (q < n_q) a.ccount[(q * a.n_split + split) * 4 + half] = cnt;
        } else {
            // theta of a row = the theta_rank-th largest of its 4 groups x 2 parities x 3 kept maxima (the transpose
            // buffers of the collect pass, unused here, carry them to the group-0 lanes)
            float *th = reinterpret_cast<float *>(smem + kHiOffScratch);      // [row][4][6]
            const int row = quad * 32 + lane;
            const bool live = q < n_q && (!active || active[q]);
            const unsigned tops[3] = {top0, top1, top2};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                // a slot that never saw a score (0x8000 = -0) counts as -inf
                const unsigned short lo16 = (unsigned short)(tops[i] & 0xffffu), hi16 = (unsigned short)(tops[i] >> 16);
                th[(row * 4 + half) * 6 + 2 * i] = lo16 == 0x8000u ? -INFINITY : __half2float(__ushort_as_half(lo16));
                th[(row * 4 + half) * 6 + 2 * i + 1] = hi16 == 0x8000u ? -INFINITY : __half2float(__ushort_as_half(hi16));
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (half == 0 && q < n_q) {
                constexpr int kMerged = 24;
                float tm[kMerged];
#pragma unroll
                for (int i = 0; i < kMerged; ++i) tm[i] = -INFINITY;
                for (int i = 0; i < kMerged; ++i) {
                    const float x = th[row * 24 + i];
                    if (x > tm[kMerged - 1]) insert_desc(tm, x);
                }
                float tsel = tm[0];
#pragma unroll
                for (int i = 1; i < kMerged; ++i) tsel = (i == a.theta_rank - 1) ? tm[i] : tsel;
                int hi_rank = a.hi_rank - 1;
                hi_rank = hi_rank < 0 ? 0 : hi_rank > kTheta - 1 ? kTheta - 1 : hi_rank;
                float thi = tm[0];
#pragma unroll
                for (int i = 1; i < kTheta; ++i) thi = i == hi_rank ? tm[i] : thi;
                a.theta[q] = live ? tsel : INFINITY;           // +inf for pruned rows, -inf if the sample was too small
                a.theta_hi[q] = live ? thi : INFINITY;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------
// Pass 3 of the fast path: one warp per query.  Every collected candidate is
// re-scored with the canonical float32 chain, the best top_k are selected
// best-first (score descending, index ascending) and the result is VERIFIED:
// a domain outside the collected set has a tensor-core score below theta, hence
// a canonical score below theta + slack (score_slack: a bound, not a measurement); if the top_k-th selected score
// reaches theta + slack nothing outside can belong to the top_k.  Queries
// that fail (too few candidates, buffer overflow, boundary within the slack) go
// on the list for the exact MODE_LISTS kernel.
// ---------------------------------------------------------------------------
constexpr int kFinWarps = 4;
constexpr int kFinKeys = 768;          // keys per query the shared-memory path of the first finalize holds (the second chance: all)
constexpr int kFinRegs = 10;           // candidates per lane the register path of finalize_kernel holds (320 per query)

// two hardware reductions (REDUX.MAX.U32): the largest high word, then the largest low word among its holders
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long x) {
    const unsigned hi = __reduce_max_sync(kFull, (unsigned)(x >> 32));
    const unsigned lo = __reduce_max_sync(kFull, (unsigned)(x >> 32) == hi ? (unsigned)x : 0u);
    return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(kFinWarps * 32, 8)       // 64 registers: the row gathers need the warps (20 bytes spilled)
finalize_kernel(const float *__restrict__ Q, const float *__restrict__ E, long long n_q, long long n_d, int top_k,
                const uint8_t *__restrict__ active, const float *theta, const int32_t *__restrict__ cbuf,
                const int *__restrict__ ccount, int cap, int parts, int q_index0, const unsigned *__restrict__ norms,
                int hi_only, int32_t *__restrict__ cand,
                float *__restrict__ scores, int *__restrict__ fail_list, int *__restrict__ fail_count,
                float *theta_retry /* may alias theta: a failed query's threshold for the second pass */, int key_cap,
                float *theta_retry16 /* the same for a second pass that filters with fp16 accumulators (may be NULL) */) {
    extern __shared__ unsigned long long fin_keys[];       // [kFinWarps][key_cap]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * kFinWarps + warp;
    if (q >= n_q) return;
    if (active && !active[q]) {
        for (int i = lane; i < top_k; i += 32) {
            cand[q * top_k + i] = -1;
            if (scores) scores[q * top_k + i] = -INFINITY;
        }
        return;
    }
    unsigned long long *keys = fin_keys + (size_t)warp * key_cap;
    // how far the filter's score may sit from the canonical one (score_slack): the margin of the proof below
    const float slack = score_slack(norms, hi_only);
    const float slack_full = score_slack(norms, 0);
    int c = 0;
    bool ok = true, overflow = false;
    for (int p = 0; p < parts; ++p) {                  // parts = 4 column groups x table splits
        const int cn = ccount[q * parts + p];
        ok = ok && cn <= cap;
        c += cn;
    }
    ok = ok && c <= key_cap;                           // more than the shared-memory path holds: treated as an overflow
    overflow = !ok;
    int n_sel = 0;
    float last = -INFINITY, first = -INFINITY;
    const int want = (long long)top_k < n_d ? top_k : (int)n_d;
    // canonical score of candidate `id` as a key (0: a padded column past the end of the table)
    float qv[ED];
#pragma unroll
    for (int k = 0; k < ED; k += 4) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(Q + q * ED + k));
        qv[k] = f.x; qv[k + 1] = f.y; qv[k + 2] = f.z; qv[k + 3] = f.w;
    }
    const bool wide_rows = (reinterpret_cast<uintptr_t>(E) & 31) == 0;     // the contract asks for 16-byte alignment only
    auto key_of = [&](int id) -> unsigned long long {
        if (id >= n_d) return 0ull;
        float ev[ED];
        if (wide_rows) {                               // two 32-byte requests per row instead of four 16-byte ones
            ldg_f8(E + (long long)id * ED, ev);
            ldg_f8(E + (long long)id * ED + 8, ev + 8);
        } else {
#pragma unroll
            for (int k = 0; k < ED; k += 4) {
                const float4 f = __ldg(reinterpret_cast<const float4 *>(E + (long long)id * ED + k));
                ev[k] = f.x; ev[k + 1] = f.y; ev[k + 2] = f.z; ev[k + 3] = f.w;
            }
        }
        return make_key(fwm::score_chain(qv, ev, ED), id);
    };
    if (ok && c <= 32 * kFinRegs) {
        // the usual case: at most kFinRegs candidates per lane, kept in registers.  Every lane sorts its own keys
        // (descending); the selection below then works on the lanes' heads.
        unsigned long long k[kFinRegs];
#pragma unroll
        for (int j = 0; j < kFinRegs; ++j) k[j] = 0ull;
        if (parts == 4) {
            // the main launch: four column groups.  Candidate number g = lane + 32 j of the query sits in the group
            // its prefix sums name; every lane finds its ten indices first (independent loads), then scores them --
            // no part-by-part walk with a divergent block per (part, j) pair
            const int o1 = ccount[q * 4], o2 = o1 + ccount[q * 4 + 1], o3 = o2 + ccount[q * 4 + 2];
            const int32_t *b = cbuf + q * 4 * (long long)cap;
            int ids[kFinRegs];
#pragma unroll
            for (int j = 0; j < kFinRegs; ++j) {
                const int g = lane + 32 * j;
                const int part = (g >= o1) + (g >= o2) + (g >= o3);
                const int first = part == 0 ? 0 : part == 1 ? o1 : part == 2 ? o2 : o3;
                ids[j] = g < c ? __ldg(b + part * cap + (g - first)) : 0x7fffffff;     // past the end: key_of gives 0
            }
#pragma unroll
            for (int j = 0; j < kFinRegs; ++j) k[j] = key_of(ids[j]);
        } else {
            int off = 0;
            for (int p = 0; p < parts; ++p) {          // candidate number g of the query = part p, entry g - off
                const int cn = ccount[q * parts + p];
                const int32_t *b = cbuf + (q * parts + p) * (long long)cap;
#pragma unroll
                for (int j = 0; j < kFinRegs; ++j) {
                    const int g = lane + 32 * j;
                    if (g >= off && g < off + cn) k[j] = key_of(b[g - off]);
                }
                off += cn;
            }
        }
        // odd-even transposition sort of kFinRegs keys (descending)
#pragma unroll
        for (int pass = 0; pass < kFinRegs; ++pass) {
#pragma unroll
            for (int j = pass & 1; j + 1 < kFinRegs; j += 2) {
                const unsigned long long hi = k[j] > k[j + 1] ? k[j] : k[j + 1], lo = k[j] > k[j + 1] ? k[j + 1] : k[j];
                k[j] = hi;
                k[j + 1] = lo;
            }
        }
        // Selection, several keys per round: a head (largest key of a lane's sorted list) that beats every lane's
        // SECOND key beats everything that is not a head, so all such heads are the next keys of the global order
        // (the largest head always is one of them).  They are ranked among themselves (one shuffle per emitting
        // lane), written in parallel and popped together: ~5 rounds for 32 keys instead of 32 rounds of warp maximum
        // + pop (each 65 instructions, 60 % of this kernel).
        while (n_sel < want) {
            const unsigned long long h = k[0];
            const unsigned long long s2 = warp_max_u64(k[1]);
            const bool emit = h > s2;                              // (h == 0: an empty list never emits)
            const unsigned em = __ballot_sync(kFull, emit);
            if (em == 0) break;                                    // nothing left
            const unsigned h_hi = (unsigned)(h >> 32), h_lo = (unsigned)h;
            int rank = 0;
            if (em & (em - 1)) {                                   // more than one: rank among the emitted heads
                for (unsigned m = em; m; m &= m - 1) {
                    const int src = __ffs(m) - 1;
                    const unsigned o_hi = __shfl_sync(kFull, h_hi, src), o_lo = __shfl_sync(kFull, h_lo, src);
                    rank += (o_hi > h_hi || (o_hi == h_hi && o_lo > h_lo)) ? 1 : 0;
                }
            }
            const int pos = n_sel + rank;
            if (emit && pos < want) {
                cand[q * top_k + pos] = (int)(0xFFFFFFFFu - h_lo);
                if (scores) scores[q * top_k + pos] = unorder_bits(h_hi);
            }
            if (n_sel == 0) {                                      // the best score of all (short queries use it)
                const unsigned src = __ballot_sync(kFull, emit && rank == 0);
                first = unorder_bits(__shfl_sync(kFull, h_hi, __ffs(src) - 1));
            }
            const int e = __popc(em);
            if (n_sel + e >= want) {                               // the top_k-th key is among these
                const unsigned src = __ballot_sync(kFull, emit && pos == want - 1);
                last = unorder_bits(__shfl_sync(kFull, h_hi, __ffs(src) - 1));
            }
            if (emit) {
#pragma unroll
                for (int j = 0; j + 1 < kFinRegs; ++j) k[j] = k[j + 1];
                k[kFinRegs - 1] = 0ull;
            }
            n_sel = n_sel + e < want ? n_sel + e : want;
        }
        ok = n_sel == want && last >= theta[q] + slack;
        for (int i = want + lane; i < top_k; i += 32) {      // table smaller than top_k: pad like the reference
            cand[q * top_k + i] = -1;
            if (scores) scores[q * top_k + i] = -INFINITY;
        }
    } else if (ok) {
        int off = 0;
        for (int p = 0; p < parts; ++p) {
            const int cn = ccount[q * parts + p];
            const int32_t *b = cbuf + (q * parts + p) * (long long)cap;
            for (int i = lane; i < cn; i += 32) keys[off + i] = key_of(b[i]);
            off += cn;
        }
        __syncwarp();
        // top_k rounds of "largest key below the previous one"
        unsigned long long prev = ~0ull;
        for (int r = 0; r < want; ++r) {
            unsigned long long best = 0ull;
            for (int i = lane; i < c; i += 32) {
                const unsigned long long kk = keys[i];
                if (kk < prev && kk > best) best = kk;
            }
            best = warp_max_u64(best);
            if (best == 0ull) break;
            if (lane == 0) {
                cand[q * top_k + r] = (int)(0xFFFFFFFFu - (uint32_t)best);
                if (scores) scores[q * top_k + r] = unorder_bits((uint32_t)(best >> 32));
            }
            prev = best;
            last = unorder_bits((uint32_t)(best >> 32));
            if (r == 0) first = last;
            ++n_sel;
        }
        ok = n_sel == want && last >= theta[q] + slack;
        for (int i = want + lane; i < top_k; i += 32) {      // table smaller than top_k: pad like the reference
            cand[q * top_k + i] = -1;
            if (scores) scores[q * top_k + i] = -INFINITY;
        }
    }
    if (!ok && lane == 0) {
        fail_list[atomicAdd(fail_count, 1)] = (int)q + q_index0;
        atomicAdd(fail_count + (overflow ? 1 : n_sel < top_k ? 2 : 3), 1);      // diagnostics: why
        // "boundary": enough candidates, but the last one sits within the slack of theta, so a better one may have
        // been filtered out.  Every domain that beats `last` scores at least last - slack_full on the tensor
        // cores with the full split: a second collect pass with this threshold finds them all and verifies.
        // A second pass on the fp16-accumulator filter needs the threshold below last - THAT filter's bound; a quarter
        // of the bound on top keeps its proof (the top_k-th score can only rise) clear of its own threshold.
        float t16 = theta[q];
        if (theta_retry && !overflow && n_sel == want && want > 0) {
            t16 = fminf(theta[q], last - 1.25f * score_slack(norms, 2));
            theta_retry[q] = fminf(theta[q], last - 2.0f * slack_full);
        }
        // "short": fewer than top_k domains reach theta at all (the tail of the sampled estimate: a handful of queries
        // per half million).  The same threshold would fail the same way, and the exact list kernel costs ~0.8 ms
        // however few queries it gets: the second pass has many times the buffer room, so it looks twice as far below
        // the best score found (any threshold is valid: the proof is checked against whatever was used).
        else if (theta_retry && !overflow && n_sel < want) {
            const float t = theta[q];
            t16 = t - fmaxf(n_sel > 0 ? first - t : 0.0f, 0.02f);
            theta_retry[q] = t16;
        }
        if (theta_retry16) theta_retry16[q] = t16;
    }
}

// After pass 1: how many live queries leave less than `room` between their estimated top_k-th score and theta?
// Few: the collect pass may filter with the hi*hi term alone.  (A wrong guess costs fallbacks, never correctness.)
__global__ void count_flat_kernel(const float *__restrict__ theta, const float *__restrict__ theta_hi, long long n_q,
                                  const unsigned *__restrict__ norms, int *__restrict__ counts) {
    // counts[0]: queries without room for the hi*hi filter (twice its error bound: theta_hi is only an estimate of the
    // top_k-th score); counts[2]: without room for the half-precision-accumulator filter (1.5 x its bound: the bound
    // assumes truncation, and a query that lacks the room only costs a second chance); counts[1]: live queries
    const float room = 2.0f * score_slack(norms, 1), room16 = 1.5f * score_slack(norms, 2);
    int flat = 0, live = 0, flat16 = 0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n_q; q += (long long)gridDim.x * blockDim.x) {
        const float t = theta[q];
        if (t == INFINITY) continue;            // pruned
        ++live;
        flat += (theta_hi[q] - t < room) ? 1 : 0;
        flat16 += (theta_hi[q] - t < room16) ? 1 : 0;
    }
    flat = __reduce_add_sync(0xffffffffu, flat);
    live = __reduce_add_sync(0xffffffffu, live);
    flat16 = __reduce_add_sync(0xffffffffu, flat16);
    if ((threadIdx.x & 31) == 0) { atomicAdd(counts, flat); atomicAdd(counts + 1, live); atomicAdd(counts + 2, flat16); }
}

// fallback plumbing: the failed queries as a dense table, and their results back in place
__global__ void gather_rows_kernel(const float *__restrict__ Q, const int *__restrict__ list, int n, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (ED / 4)) return;
    const int r = i / (ED / 4), k = i % (ED / 4);
    reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(Q + (long long)list[r] * ED) + k);
}
__global__ void gather_f32_kernel(const float *__restrict__ src, const int *__restrict__ list, int n, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[list[i]];
}
__global__ void scatter_cand_kernel(const int32_t *__restrict__ src, const float *__restrict__ src_scores,
                                    const int *__restrict__ list, int n, int top_k, int32_t *__restrict__ cand,
                                    float *__restrict__ scores) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * top_k) return;
    const long long dst = (long long)list[i / top_k] * top_k + i % top_k;
    cand[dst] = src[i];
    if (scores) scores[dst] = src_scores[i];
}

// MODE_LISTS with n_split > 1: one warp per query merges the partial lists (n_split x 2 * kCap keys):
// exact re-score of every kept candidate in place, then top_k rounds of "largest key below the previous one".
__global__ void __launch_bounds__(128)
merge_parts_kernel(const float *__restrict__ Q, const float *__restrict__ E, long long n_q, int n_split, int top_k,
                   const uint8_t *__restrict__ active, unsigned long long *__restrict__ parts, int32_t *__restrict__ cand,
                   float *__restrict__ scores, const unsigned *__restrict__ norms, int *__restrict__ lfail_list,
                   int *__restrict__ lfail_count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * 4 + warp;
    if (q >= n_q) return;
    if (active && !active[q]) {
        for (int i = lane; i < top_k; i += 32) {
            cand[q * top_k + i] = -1;
            if (scores) scores[q * top_k + i] = -INFINITY;
        }
        return;
    }
    unsigned long long *keys = parts + q * n_split * (2 * kCap);
    const int c = n_split * 2 * kCap;
    const float *qv = Q + q * ED;
    // the most an evicted candidate of any (split, column half) list can have scored on the tensor cores
    float evict_max = -INFINITY;
    for (int l = lane; l < 2 * n_split; l += 32)
        evict_max = fmaxf(evict_max, unorder_bits((uint32_t)(keys[(size_t)l * kCap + kKeep - 1] >> 32)));
#pragma unroll
    for (int o = 16; o; o >>= 1) evict_max = fmaxf(evict_max, __shfl_xor_sync(kFull, evict_max, o));
    __syncwarp();
    for (int i = lane; i < c; i += 32) {
        unsigned long long k = keys[i];
        if ((uint32_t)(k >> 32) != kNegInfBits) {
            const int id = (int)(0xFFFFFFFFu - (uint32_t)k);
            k = make_key(fwm::score_chain(qv, E + (long long)id * ED, ED), id);
        } else {
            k = 0ull;
        }
        keys[i] = k;
    }
    __syncwarp();
    unsigned long long prev = ~0ull;
    int r = 0;
    for (; r < top_k; ++r) {
        unsigned long long best = 0ull;
        for (int i = lane; i < c; i += 32) {
            const unsigned long long k = keys[i];
            if (k < prev && k > best) best = k;
        }
        best = warp_max_u64(best);
        if (best == 0ull) break;
        if (lane == 0) {
            cand[q * top_k + r] = (int)(0xFFFFFFFFu - (uint32_t)best);
            if (scores) scores[q * top_k + r] = unorder_bits((uint32_t)(best >> 32));
        }
        prev = best;
    }
    for (int i = r + lane; i < top_k; i += 32) {
        cand[q * top_k + i] = -1;
        if (scores) scores[q * top_k + i] = -INFINITY;
    }
    // same proof as in scan_kernel<MODE_LISTS>: the top_k-th canonical score must beat what an evicted candidate
    // could have reached
    const float kth = r == top_k ? unorder_bits((uint32_t)(prev >> 32)) : -INFINITY;
    if (lane == 0 && lfail_list && evict_max != -INFINITY && !(kth > evict_max + score_slack(norms, 0)))
        lfail_list[atomicAdd(lfail_count, 1)] = (int)q;
}

}  // namespace

// The list kernel keeps 48 candidates per row (top_k <= 32); the fast path collects ~256 per query and selects any
// top_k up to 64 from them (BASELINE.json config 4), handing its rare failures to the FFMA kernel when top_k > 32.
bool fwav_topk_umma_supported(int emb_dim, int top_k, int64_t n_q, int64_t n_d) {
    if (emb_dim != ED || top_k < 1 || n_q <= 0 || n_d <= 0) return false;
    return top_k <= 32 || (top_k <= 64 && n_d >= (1 << 16));
}

namespace {

constexpr int kCollectCap = 256;              // candidate indices kept per (query, column group of 64)
constexpr int kCollectCapWide = 320;          // the same for top_k > 32 (theta is the 24th best sampled score there)
constexpr long long kFastMinDomains = 1 << 16; // below this the sample is too small for a useful threshold
constexpr double kFailWeight = 2.5;           // route choice: passes' worth of second chance per fraction of queries without room
constexpr int kSampleStride = 16;             // pass 1 looks at every 16th domain
constexpr bool kDefaultTheta16 = true;        // threshold pass through collect_hi_kernel<true> (fp16 accumulators, chunk maxima)
constexpr bool kDefaultAcc16 = true;          // half-precision accumulators in the hi*hi-only collect pass (FWAV_UMMA_ACC16=0: off)
constexpr long long kBatchQueries = 1 << 20;  // queries per fast-path batch (bounds the candidate buffers: 3 GB)

inline int grid_for(const fwav_ctx *ctx, long long work) {
    long long need = (work + 255) / 256, cap = (long long)ctx->num_sms * 8;
    return (int)(need < cap ? need : cap);
}

// one launch of the scan skeleton: `groups` tensor-core groups of 128 * CG queries, each scanned by `split` of them
template <int MODE, bool HI, int CG, bool COMPACT = false>
int launch_scan(fwav_ctx *ctx, const ScanArgs &a, long long groups, long long split, cudaStream_t st) {
    constexpr int smem = (int)smem_bytes(MODE, CG);
    // function attributes are per device: set before every launch (a process may hold contexts on several GPUs)
    FWAV_CUDA(ctx, cudaFuncSetAttribute(scan_kernel<MODE, HI, CG, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    scan_kernel<MODE, HI, CG, COMPACT><<<(unsigned)(CG * groups * split), n_threads(MODE), smem, st>>>(a);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

// the four-issuer hi*hi-only collect pass: one CTA per 128 queries and table share
template <bool THETA = false>
int launch_collect_hi(fwav_ctx *ctx, const ScanArgs &a, long long groups, long long split, cudaStream_t st) {
    FWAV_CUDA(ctx, cudaFuncSetAttribute(collect_hi_kernel<THETA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHiSmem));
    collect_hi_kernel<THETA><<<(unsigned)(groups * split), kHiThreads, kHiSmem, st>>>(a);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

// record phase boundary k of timed batch `slot` (events are created on first use)
int mark(fwav_ctx *ctx, int slot, int k, cudaStream_t st) {
    if (slot >= fwav_ctx::kSearchSlots) return FWAV_OK;
    if (!ctx->search_ev[slot][k]) FWAV_CUDA(ctx, cudaEventCreate(&ctx->search_ev[slot][k]));
    FWAV_CUDA(ctx, cudaEventRecord(ctx->search_ev[slot][k], st));
    return FWAV_OK;
}

// exact list kernel over packed tiles (small tables, forced mode, fallback of the fast path)
int launch_lists(fwav_ctx *ctx, const uint4 *d_qt, const uint4 *d_et, const float *d_q, const float *d_emb, long long n_q,
                 long long n_d, int n_stages, int top_k, const uint8_t *d_active, int32_t *d_cand, float *d_scores, int dbg,
                 cudaStream_t st, const unsigned *d_norms, bool compact = false) {
    ScanArgs a = {};
    a.q_tiles = d_qt; a.e_tiles = d_et; a.Q = d_q; a.E = d_emb; a.n_q = n_q; a.n_d = n_d;
    a.n_stages = n_stages; a.top_k = top_k; a.active = d_active; a.cand = d_cand; a.scores = d_scores;
    a.dbg = dbg;
    // rows whose lists cannot prove their top_k (a boundary crowded within the filter's error) go to the FFMA kernel
    int *d_lfail = nullptr;
    {
        int rc0;
        if ((rc0 = fwav_ws_reserve(ctx, WS_UMMA_LFAIL, (size_t)(n_q + 4) * sizeof(int), (void **)&d_lfail))) return rc0;
        FWAV_CUDA(ctx, cudaMemsetAsync(d_lfail + n_q, 0, sizeof(int), st));
    }
    a.lfail_list = d_lfail; a.lfail_count = d_lfail + n_q; a.norms = d_norms;
    const long long q_pairs = (n_q + kQPair - 1) / kQPair;
    // few queries: split the table between several CTA pairs per 256 queries so that the machine is full
    const long long resident = ctx->num_sms / 2;
    long long split = 1;
    if (q_pairs < resident && !dbg) {
        split = resident / q_pairs;
        if (split > 64) split = 64;
        if (split > n_stages / 8) split = n_stages / 8;
        if (split < 1) split = 1;
    }
    a.n_split = (int)split;
    if (split > 1) {
        int rc;
        if ((rc = fwav_ws_reserve(ctx, WS_UMMA_PARTS, (size_t)n_q * split * 2 * kCap * 8, (void **)&a.parts))) return rc;
    }
    {
        int rc = compact ? launch_scan<MODE_LISTS, false, 2, true>(ctx, a, q_pairs, split, st)
                         : launch_scan<MODE_LISTS, false, 2>(ctx, a, q_pairs, split, st);
        if (rc) return rc;
    }
    if (split > 1) {
        merge_parts_kernel<<<(unsigned)((n_q + 3) / 4), 128, 0, st>>>(d_q, d_emb, n_q, (int)split, top_k, d_active, a.parts,
                                                                    d_cand, d_scores, d_norms, a.lfail_list, a.lfail_count);
        FWAV_LAUNCH_CHECK(ctx);
    }
    int n_lfail = 0;
    FWAV_CUDA(ctx, cudaMemcpyAsync(&n_lfail, d_lfail + n_q, sizeof(int), cudaMemcpyDeviceToHost, st));
    FWAV_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_lfail > 0 && !dbg) {
        if (getenv("FWAV_UMMA_VERBOSE"))
            fprintf(stderr, "[fwav] list kernel: %d of %lld rows could not prove their top_k: FFMA kernel\n", n_lfail, n_q);
        ctx->umma_ffma_queries += n_lfail;
        unsigned char *blk = nullptr;
        const size_t sz_q = (((size_t)n_lfail * ED * sizeof(float)) + 255) & ~(size_t)255,
                     sz_c = (((size_t)n_lfail * top_k * sizeof(int32_t)) + 255) & ~(size_t)255;
        int rc1;
        if ((rc1 = fwav_ws_reserve(ctx, WS_UMMA_PARTS, sz_q + 2 * sz_c, (void **)&blk))) return rc1;     // the parts are consumed
        float *d_fq = reinterpret_cast<float *>(blk);
        int32_t *d_fc = reinterpret_cast<int32_t *>(blk + sz_q);
        float *d_fs = reinterpret_cast<float *>(blk + sz_q + sz_c);
        gather_rows_kernel<<<(n_lfail * (ED / 4) + 255) / 256, 256, 0, st>>>(d_q, d_lfail, n_lfail, d_fq);
        FWAV_LAUNCH_CHECK(ctx);
        if ((rc1 = fwav_launch_topk_ffma(ctx, d_fq, n_lfail, d_emb, n_d, ED, top_k, nullptr, d_fc, d_fs, st))) return rc1;
        scatter_cand_kernel<<<(n_lfail * top_k + 255) / 256, 256, 0, st>>>(d_fc, d_fs, d_lfail, n_lfail, top_k, d_cand, d_scores);
        FWAV_LAUNCH_CHECK(ctx);
    }
    return FWAV_OK;
}

}  // namespace

int fwav_launch_topk_umma(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb, int64_t n_d,
                          int emb_dim, int top_k, const uint8_t *d_active, int32_t *d_cand, float *d_scores,
                          cudaStream_t st) {
    FWAV_REQUIRE(ctx, fwav_topk_umma_supported(emb_dim, top_k, n_q, n_d),
                 "tensor-core search is built for emb_dim=16 and top_k<=32 (<=64 from 65536 domains up) (got %d, %d)",
                 emb_dim, top_k);
    FWAV_REQUIRE(ctx, n_d < (1ll << 31) - kDStage, "n_domains %lld does not fit the int32 match index", (long long)n_d);
    FWAV_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_emb)) & 15) == 0,
                 "embedding tables must be 16-byte aligned");
    const long long n_stages = (n_d + kDStage - 1) / kDStage;      // 256 domains each: two packed tiles
    const long long e_tiles = 2 * n_stages;
    const long long q_pairs = (n_q + kQPair - 1) / kQPair;
    const long long q_tiles = q_pairs * 2;
    FWAV_REQUIRE(ctx, 2 * q_pairs < (1ll << 31), "too many queries for one launch (%lld)", (long long)n_q);
    uint4 *d_et = nullptr, *d_qt = nullptr;
    int rc;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_E, (size_t)e_tiles * kTileBytes, (void **)&d_et))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_Q, (size_t)q_tiles * kTileBytes, (void **)&d_qt))) return rc;
    ctx->search_slots_used = 0;
    if ((rc = mark(ctx, 0, 0, st))) return rc;
    // Compact split (see pack_compact_tiles_kernel): at most eight live embedding dimensions leave room for the hi AND
    // the lo part in one K = 16 operand -- two MMAs per stage instead of three for the full split (range_size 4, the
    // reference's default tile_size: 3 tonal + 4 transient dimensions; config-4 shape collect pass 1 824 -> 1 706 ms).
    // Which dimensions can be non-zero is known without looking at the data when the caller states the range_size
    // the tables were embedded for (the pipeline entry points do; fwav_ctx_set_search_range_size for fwav_topk);
    // FWAV_UMMA_COMPACT=1 probes arbitrary tables on the device (one round trip), =0 switches the split off.
    LivePerm perm = {};
    bool compact = false;
    {
        const char *c_env = getenv("FWAV_UMMA_COMPACT"), *cg2_env = getenv("FWAV_UMMA_CG");
        const bool allowed = !(c_env && atoi(c_env) == 0) && !(cg2_env && atoi(cg2_env) == 2) && n_d >= (1 << 16);
        unsigned live = 0xffffu;
        bool known = false;
        if (allowed && c_env && atoi(c_env) == 1) {
            unsigned *d_mask = nullptr;
            if ((rc = fwav_ws_reserve(ctx, WS_UMMA_THETA, 2 * sizeof(unsigned), (void **)&d_mask))) return rc;
            FWAV_CUDA(ctx, cudaMemsetAsync(d_mask, 0, 2 * sizeof(unsigned), st));
            live_dims_kernel<<<grid_for(ctx, n_d * (ED / 4)), 256, 0, st>>>(d_emb, n_d, d_mask);
            FWAV_LAUNCH_CHECK(ctx);
            live_dims_kernel<<<grid_for(ctx, n_q * (ED / 4)), 256, 0, st>>>(d_q, n_q, d_mask + 1);
            FWAV_LAUNCH_CHECK(ctx);
            unsigned h_mask[2] = {0, 0};
            FWAV_CUDA(ctx, cudaMemcpyAsync(h_mask, d_mask, sizeof h_mask, cudaMemcpyDeviceToHost, st));
            FWAV_CUDA(ctx, cudaStreamSynchronize(st));
            live = h_mask[0] & h_mask[1];    // a dimension dead on either side adds exactly 0 to every score
            known = true;
        } else if (allowed && ctx->search_range_size > 0 && ctx->embed_kind == FWAV_EMBED_TWO_HEAD) {
            // tables built by fwav_embed for this range_size: min(8, N - 1) tonal and min(8, N) transient dimensions
            const int N = ctx->search_range_size;
            live = 0;
            for (int k = 0; k < ED / 2 && k < N - 1; ++k) live |= 1u << k;
            for (int k = 0; k < ED / 2 && k < N; ++k) live |= 1u << (ED / 2 + k);
            known = true;
        }
        if (known && __builtin_popcount(live) <= 8) {
            compact = true;
            for (int k = 0; k < ED; ++k)
                if (live >> k & 1) perm.dim[perm.n++] = k;
        }
        if (getenv("FWAV_UMMA_VERBOSE"))
            fprintf(stderr, "[fwav] live embedding dimensions %04x (%s): %s split\n", live, known ? "known" : "unknown",
                    compact ? "compact (2 MMAs per stage)" : "plain");
    }
    // largest row norms of both tables: the error bounds of the filter scores scale with them (score_slack)
    unsigned *d_norms = nullptr;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_NORMS, 4 * sizeof(unsigned), (void **)&d_norms))) return rc;
    FWAV_CUDA(ctx, cudaMemsetAsync(d_norms, 0, 2 * sizeof(unsigned), st));
    row_norm2_max_kernel<<<grid_for(ctx, n_q), 256, 0, st>>>(d_q, n_q, d_norms);
    FWAV_LAUNCH_CHECK(ctx);
    row_norm2_max_kernel<<<grid_for(ctx, n_d), 256, 0, st>>>(d_emb, n_d, d_norms + 1);
    FWAV_LAUNCH_CHECK(ctx);
    // row-major float32 -> packed fp16 tiles (role 0: queries, 1: domains; stride > 1: the sample table of pass 1)
    auto pack = [&](const float *src, long long rows, long long tiles, uint4 *dst, int stride, int role) -> int {
        if (compact)
            pack_compact_tiles_kernel<<<grid_for(ctx, tiles * kDTile * 2), 256, 0, st>>>(src, rows, tiles, dst, stride, perm, role);
        else
            pack_f16_tiles_kernel<<<grid_for(ctx, tiles * kDTile * 2), 256, 0, st>>>(src, rows, tiles, dst, stride);
        FWAV_LAUNCH_CHECK(ctx);
        return FWAV_OK;
    };
    if ((rc = pack(d_emb, n_d, e_tiles, d_et, 1, 1))) return rc;
    if ((rc = pack(d_q, n_q, q_tiles, d_qt, 1, 0))) return rc;
#ifdef FWAV_DEBUG_KNOBS
    const char *dbg_env = getenv("FWAV_UMMA_DEBUG");   // profiling aid (results are WRONG when set): debug builds only
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
#else
    const int dbg = 0;
#endif
    const char *mode_env = getenv("FWAV_UMMA_MODE");   // "lists": force the exact list kernel
    const bool fast = n_d >= kFastMinDomains && (top_k > 32 || !(mode_env && !strcmp(mode_env, "lists")));
    ctx->search_fast_path = fast;
    ctx->search_route = 0;
    if (!fast) {
        for (int k = 1; k <= 4; ++k)
            if ((rc = mark(ctx, 0, k, st))) return rc;
        if ((rc = launch_lists(ctx, d_qt, d_et, d_q, d_emb, n_q, n_d, (int)n_stages, top_k, d_active, d_cand, d_scores, dbg, st, d_norms, compact)))
            return rc;
        if ((rc = mark(ctx, 0, 5, st))) return rc;
        ctx->search_slots_used = 1;
        return FWAV_OK;
    }

    // ---- fast path: sampled threshold, collect, finalize + verify, exact fallback for the failures ----
    // pass 1 scans a strided sample of the table (every 16th domain), packed like the table itself:
    // neighbouring domains are near-duplicates of each other, a strided sample is not
    // top_k <= 32: every 32nd domain and the 8th best sampled score (256 candidates expected, like 16 x 16, at
    // half the cost of pass 1: 7.75 -> 4.4 ms on config 2; the wider spread costs ~40 instead of ~1 second chances)
    int sample_stride = top_k > 32 ? kSampleStride : 2 * kSampleStride;
    if (const char *stride_env = getenv("FWAV_UMMA_STRIDE")) {     // tuning knob (with FWAV_UMMA_RANK: stride x rank candidates)
        const int v = atoi(stride_env);
        if (v >= 2 && v <= 256) sample_stride = v;
    }
    const long long n_samp = (n_d + sample_stride - 1) / sample_stride;
    const long long s_stages = (n_samp + kDStage - 1) / kDStage;
    uint4 *d_es = nullptr;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_MISC, (size_t)s_stages * 2 * kTileBytes, (void **)&d_es))) return rc;
    if ((rc = pack(d_emb, n_d, s_stages * 2, d_es, sample_stride, 1))) return rc;
    const char *cg_env = getenv("FWAV_UMMA_CG");
    const bool single = !(cg_env && atoi(cg_env) == 2);
    long long batch_cap = kBatchQueries;
    if (const char *batch_env = getenv("FWAV_UMMA_BATCH")) {  // test knob: small batches exercise the multi-batch loop
        const long long v = atoll(batch_env);
        if (v >= kQPair) batch_cap = v / kQPair * kQPair;
    }
    const long long batch = n_q < batch_cap ? n_q : batch_cap;
    float *d_theta = nullptr;
    int32_t *d_cbuf = nullptr;
    int *d_cnt = nullptr, *d_fail = nullptr;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_THETA, (size_t)(2 * n_q + 32) * sizeof(float), (void **)&d_theta))) return rc;
    float *d_theta_hi = d_theta + n_q;
    int *d_flat = reinterpret_cast<int *>(d_theta + 2 * n_q);
    // collect_hi_kernel's shared table position (ScanArgs::front): one int for the main launch, one per split of the tail
    int *d_front = d_flat + 8;
    const char *front_env = getenv("FWAV_UMMA_FRONT");         // 0: every CTA starts at its own static position
    const bool use_front = !(front_env && atoi(front_env) == 0);

    int collect_cap = top_k > 32 ? kCollectCapWide : kCollectCap;
    if (const char *cap_env = getenv("FWAV_UMMA_CAP")) {       // test knob: small buffers force the failure paths
        const int v = atoi(cap_env);
        if (v >= 2 && v <= collect_cap) collect_cap = v & ~1;
    }
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_CBUF, (size_t)batch * 4 * collect_cap * sizeof(int32_t), (void **)&d_cbuf))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_CNT, (size_t)batch * 4 * sizeof(int), (void **)&d_cnt))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_FAIL, (size_t)(n_q + 4) * sizeof(int), (void **)&d_fail))) return rc;
    int *d_fail_count = d_fail + n_q;
    FWAV_CUDA(ctx, cudaMemsetAsync(d_fail_count, 0, 4 * sizeof(int), st));
    // stride x rank candidates expected per query.  P(fewer than top_k reach theta) = P(Bin(top_k, 1/stride) >= rank):
    // 1e-5 for (32; 1/32, 8) -- a handful of second chances per half million queries -- and 5e-8 for (64; 1/16, 18).
    // Fewer candidates do not pay for top_k <= 32: they leave too little room between the top_k-th score and theta
    // for the hi*hi-only collect pass (config 2 at 16 x 12: 1 % of the queries under 4e-3).
    // threshold pass through collect_hi_kernel<true> (fp16 accumulators, chunk maxima) where the CTAs run on their own and
    // the table is not compact; its r-th largest chunk maximum sits a little below the r-th largest sampled score, so
    // rank 7 there collects what rank 8 collects with scan_kernel (config 2: 53.2 ms per search against 54.3 at rank 8)
    const char *theta16_env = getenv("FWAV_UMMA_THETA16");      // 0: float32 accumulators (scan_kernel<MODE_THETA>)
    const bool theta16 = single && !compact && !dbg && !getenv("FWAV_UMMA_THETA_FULL") &&
                         (theta16_env ? atoi(theta16_env) != 0 : kDefaultTheta16);
    int theta_rank = top_k > 32 ? 18 : theta16 ? 7 : kTheta / 2;
    if (const char *rank_env = getenv("FWAV_UMMA_RANK")) {     // tuning knob: 16 x rank candidates expected per query
        const int v = atoi(rank_env);
        if (v >= 1 && v <= 4 * kThetaPart) theta_rank = v;
    }
    int slot = 0;
    for (long long q0 = 0; q0 < n_q; q0 += batch, ++slot) {
        if (slot > 0 && (rc = mark(ctx, slot, 0, st))) return rc;      // later batches: nothing to pack
        if ((rc = mark(ctx, slot, 1, st))) return rc;
        const long long nq = n_q - q0 < batch ? n_q - q0 : batch;      // batch is a multiple of kQPair unless it is everything
        const long long pairs = (nq + kQPair - 1) / kQPair;
        ScanArgs a = {};
        a.q_tiles = d_qt + (q0 / kQTile) * (kTileBytes / 16);
        a.e_tiles = d_et; a.Q = d_q + q0 * ED; a.E = d_emb; a.n_q = nq; a.n_d = n_d;
        a.n_stages = (int)n_stages; a.top_k = top_k; a.active = d_active ? d_active + q0 : nullptr;
        a.theta = d_theta + q0; a.theta_hi = d_theta_hi + q0; a.theta_rank = theta_rank; a.hi_rank = top_k / sample_stride; a.cbuf = d_cbuf; a.ccount = d_cnt; a.cap = collect_cap; a.dbg = dbg;
        a.n_split = 1;
        a.e_tiles = d_es; a.n_stages = (int)s_stages;
        // pass 1 keeps the full split: on data whose scores crowd together a threshold that is off by the hi*hi
        // error (1e-3) lands hundreds of ranks away from where it should
        // streaming modes: CTAs on their own by default (FWAV_UMMA_CG=2: CTA pairs)
        const long long groups = single ? (nq + kQTile - 1) / kQTile : pairs;
        if ((rc = compact ? launch_scan<MODE_THETA, false, 1, true>(ctx, a, groups, 1, st)
                  // (on its own CTA the threshold pass takes the hi*hi term alone: a threshold is an estimate anyway, and
                  // the proof of a query never relies on how it was obtained; FWAV_UMMA_THETA_FULL=1: the full split)
                  : single ? (getenv("FWAV_UMMA_THETA_FULL") ? launch_scan<MODE_THETA, false, 1>(ctx, a, groups, 1, st)
                              : theta16 ? launch_collect_hi<true>(ctx, a, groups, 1, st)
                                                             : launch_scan<MODE_THETA, true, 1>(ctx, a, groups, 1, st))
                           : launch_scan<MODE_THETA, false, 2>(ctx, a, groups, 1, st)))
            return rc;
        // may pass 2 filter with the hi*hi term too?  Only if (nearly) every query has room for its error bound
        bool hi_only = false, acc16 = false;
        if (!(mode_env && !strcmp(mode_env, "precise"))) {
            FWAV_CUDA(ctx, cudaMemsetAsync(d_flat, 0, 3 * sizeof(int), st));
            count_flat_kernel<<<grid_for(ctx, nq), 256, 0, st>>>(d_theta + q0, d_theta_hi + q0, nq, d_norms, d_flat);
            FWAV_LAUNCH_CHECK(ctx);
            int h_flat[3] = {0, 0, 0};
            FWAV_CUDA(ctx, cudaMemcpyAsync(h_flat, d_flat, sizeof h_flat, cudaMemcpyDeviceToHost, st));
            FWAV_CUDA(ctx, cudaStreamSynchronize(st));
            // (a query without that room is not lost: it fails verification and takes the second chance below,
            // which is cheap next to the +35 % of a full-split pass)
            const char *a16_env = getenv("FWAV_UMMA_ACC16");
            const bool a16_ok = single && !(dbg & 0xffff & ~(64 | 128 | 256)) && (a16_env ? atoi(a16_env) != 0 : kDefaultAcc16);
            if (compact || h_flat[1] == 0) {
                // (compact tiles carry hi and lo in one part: no hi*hi-only form of them)
                hi_only = h_flat[1] > 0 && (double)h_flat[0] <= 0.02 * h_flat[1];
                acc16 = hi_only && a16_ok && (double)h_flat[2] <= 0.02 * h_flat[1];
            } else {
                // The cheapest of three, in units of the fp16-accumulator pass over the whole batch (measured on config 2:
                // float32 accumulators 1.35, full split 1.85).  A query that lacks the room MAY fail its proof (on config 2
                // one in four does) and then costs a second chance: full split, own launch, partial waves -- kFailWeight
                // passes' worth per failing fraction, a deliberately pessimistic figure.
                const double f16 = (double)h_flat[2] / h_flat[1], f32 = (double)h_flat[0] / h_flat[1];
                const double c16 = a16_ok ? 1.0 + kFailWeight * f16 : 1e30, c32 = 1.35 + kFailWeight * f32, cfull = 1.85;
                acc16 = c16 <= c32 && c16 <= cfull;
                hi_only = acc16 || c32 <= cfull;
            }
            if (mode_env && (!strcmp(mode_env, "hionly") || !strcmp(mode_env, "acc16"))) hi_only = true;
            if (mode_env && !strcmp(mode_env, "acc16") && single && !dbg) acc16 = true;
            if (mode_env && !strcmp(mode_env, "hionly")) acc16 = false;
            if (getenv("FWAV_UMMA_VERBOSE"))
                fprintf(stderr, "[fwav] search batch at %lld: of %d live queries %d lack the room for the hi*hi filter and %d for half-precision accumulators: %s collect pass\n",
                        q0, h_flat[1], h_flat[0], h_flat[2], acc16 ? "hi*hi-only, fp16 accumulators" : hi_only ? "hi*hi-only" : "full-split");
        }
        ctx->search_hi_only = hi_only;
        ctx->search_route = acc16 ? 3 : hi_only ? 2 : 1;
        if ((rc = mark(ctx, slot, 2, st))) return rc;
        a.e_tiles = d_et; a.n_stages = (int)n_stages;
        if (dbg & 64) {
            if ((rc = fwav_ws_reserve(ctx, WS_UMMA_FB, kTraceStages * 8 * sizeof(long long), (void **)&a.trace))) return rc;
            FWAV_CUDA(ctx, cudaMemsetAsync(a.trace, 0, kTraceStages * 8 * sizeof(long long), st));
        }
        // The last, partial wave of CTAs would take as long as a full one: split the table between several CTAs
        // per 128 queries there (own candidate buffers, `parts` = 4 x splits), so that the tail takes 1/splits.
        long long main_groups = groups, tail_groups = 0, tail_split = 1;
        if (single && !dbg && groups > ctx->num_sms) {
            const long long rem = groups % ctx->num_sms;
            if (rem > 0 && rem <= ctx->num_sms / 2) {
                tail_split = ctx->num_sms / rem;
                if (tail_split > 8) tail_split = 8;
                if (tail_split >= 2) { tail_groups = rem; main_groups = groups - rem; } else tail_split = 1;
            }
        }
        const long long main_q = tail_groups ? main_groups * kQTile : nq;
        const int tail_cap = collect_cap / 2;      // a split sees 1/tail_split of the table: ~256 / (4 * tail_split) hits per part expected
        if (use_front && acc16) {
            FWAV_CUDA(ctx, cudaMemsetAsync(d_front, 0, 16 * sizeof(int), st));
            a.front = d_front;
        }
        ScanArgs at = a;
        if (tail_groups) {
            int32_t *d_tbuf = nullptr;
            const long long tq = nq - main_q;
            const size_t nb = (size_t)tq * tail_split * 4 * tail_cap * sizeof(int32_t), nc = (size_t)tq * tail_split * 4 * sizeof(int);
            if ((rc = fwav_ws_reserve(ctx, WS_UMMA_TAIL, nb + nc, (void **)&d_tbuf))) return rc;
            a.n_q = main_q;
            at.q_tiles = a.q_tiles + main_groups * (kTileBytes / 16);
            at.Q = a.Q + main_q * ED;
            at.active = a.active ? a.active + main_q : nullptr;
            at.theta = a.theta + main_q;
            at.n_q = tq;
            at.n_split = (int)tail_split;
            at.cbuf = d_tbuf;
            at.ccount = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(d_tbuf) + nb);
            at.cap = tail_cap;
            if (a.front) at.front = d_front + 8;       // tail_split <= 8
        }
        for (int part = 0; part < (tail_groups ? 2 : 1); ++part) {
            const ScanArgs &ax = part ? at : a;
            const long long g = part ? tail_groups : main_groups, sp = part ? tail_split : 1;
            if (compact && !hi_only)
                rc = launch_scan<MODE_COLLECT, false, 1, true>(ctx, ax, g, sp, st);
            else if (acc16)
                rc = launch_collect_hi(ctx, ax, g, sp, st);
            else if (hi_only)
                rc = single ? launch_scan<MODE_COLLECT, true, 1>(ctx, ax, g, sp, st) : launch_scan<MODE_COLLECT, true, 2>(ctx, ax, g, sp, st);
            else
                rc = single ? launch_scan<MODE_COLLECT, false, 1>(ctx, ax, g, sp, st) : launch_scan<MODE_COLLECT, false, 2>(ctx, ax, g, sp, st);
            if (rc) return rc;
        }
        if ((rc = mark(ctx, slot, 3, st))) return rc;
#ifdef FWAV_DEBUG_KNOBS
        if (dbg & 64) {      // clock64 stamps of CTA 0, printed to stderr (debug builds only; no file I/O in the library)
            static long long h_trace[kTraceStages * 8];
            FWAV_CUDA(ctx, cudaMemcpyAsync(h_trace, a.trace, sizeof h_trace, cudaMemcpyDeviceToHost, st));
            FWAV_CUDA(ctx, cudaStreamSynchronize(st));
            fprintf(stderr, acc16 ? "# stage: loop_top operands_ok acc_free issued | acc_full loaded_h0 reduced loaded_h1 (cycles rel. to first)\n"
                                  : "# stage: full_ok tempty_ok issued | ld_a_done released early_try look_b_done late_wait_done (cycles rel. to first)\n");
            const long long t0 = h_trace[0];
            for (int i = 0; i < kTraceStages; ++i) {
                fprintf(stderr, "%4d:", (acc16 ? ((dbg >> 16) & 0xff) * 64 : kTraceFrom) + i);
                for (int k = 0; k < 8; ++k) fprintf(stderr, " %8lld", h_trace[i * 8 + k] ? h_trace[i * 8 + k] - t0 : -1ll);
                fprintf(stderr, "\n");
            }
        }
#endif
        // A batch that took the fp16-accumulator route gives its failed queries their second chance on the same
        // kernel (2.3 x cheaper per stage than the full split): a boundary case only needs its threshold below
        // (K-th score - the filter's bound), whatever the filter; the table split gives it the room.
        const bool retry16_ok = acc16 && !compact && single && !(mode_env && !strcmp(mode_env, "noretry")) &&
                                !(getenv("FWAV_UMMA_RETRY16") && atoi(getenv("FWAV_UMMA_RETRY16")) == 0);
        for (int part = 0; part < (tail_groups ? 2 : 1); ++part) {
            const ScanArgs &ax = part ? at : a;
            const long long qoff = part ? main_q : 0;
            const int parts = 4 * ax.n_split;
            const int key_cap = parts * ax.cap < kFinKeys ? parts * ax.cap : kFinKeys;
            const size_t fin_smem = (size_t)kFinWarps * key_cap * sizeof(unsigned long long);
            if (fin_smem > 48 * 1024)
                FWAV_CUDA(ctx, cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
            finalize_kernel<<<(unsigned)((ax.n_q + kFinWarps - 1) / kFinWarps), kFinWarps * 32,
                              fin_smem, st>>>(
                ax.Q, d_emb, ax.n_q, n_d, top_k, ax.active, ax.theta, ax.cbuf, ax.ccount, ax.cap, parts, (int)qoff,
                d_norms, acc16 ? 2 : hi_only ? 1 : 0, d_cand + (q0 + qoff) * top_k,
                d_scores ? d_scores + (q0 + qoff) * top_k : nullptr, d_fail, d_fail_count,
                d_theta + q0 + qoff, key_cap, d_theta_hi + q0 + qoff);     // (theta_hi has served its purpose: count_flat_kernel)
            FWAV_LAUNCH_CHECK(ctx);
        }
        if ((rc = mark(ctx, slot, 4, st))) return rc;
        // NOTE: fail_list holds batch-local indices; resolve this batch's failures before the next one
        int h_fail[4] = {0, 0, 0, 0};
        FWAV_CUDA(ctx, cudaMemcpyAsync(h_fail, d_fail_count, sizeof h_fail, cudaMemcpyDeviceToHost, st));
        FWAV_CUDA(ctx, cudaStreamSynchronize(st));
        const int n_fail = h_fail[0];
        ctx->umma_fallback_queries += n_fail;
        if (getenv("FWAV_UMMA_VERBOSE"))
            fprintf(stderr, "[fwav] search batch at %lld: %d of %lld queries to the exact list kernel (overflow %d, short %d, boundary %d)\n",
                    q0, n_fail, nq, h_fail[1], h_fail[2], h_fail[3]);
        if (n_fail > 0 && dbg) FWAV_CUDA(ctx, cudaMemsetAsync(d_fail_count, 0, 4 * sizeof(int), st));   // profiling modes: wrong anyway
        if (n_fail > 0 && !dbg) {
            const long long fp = (n_fail + kQPair - 1) / kQPair;
            unsigned char *blk = nullptr;
            const size_t sz_q = (size_t)fp * kQPair * ED * sizeof(float), sz_t = (size_t)fp * 2 * kTileBytes,
                         sz_c = (((size_t)n_fail * top_k * sizeof(int32_t)) + 255) & ~(size_t)255,
                         sz_n = (((size_t)n_fail + 4) * sizeof(int) + 255) & ~(size_t)255;
            if ((rc = fwav_ws_reserve(ctx, WS_UMMA_FB, 2 * sz_q + 2 * sz_t + 4 * sz_c + 2 * sz_n, (void **)&blk))) return rc;
            float *d_fq = reinterpret_cast<float *>(blk);
            uint4 *d_fqt = reinterpret_cast<uint4 *>(blk + sz_q);
            int32_t *d_fc = reinterpret_cast<int32_t *>(blk + sz_q + sz_t);
            float *d_fs = reinterpret_cast<float *>(blk + sz_q + sz_t + sz_c);
            float *d_ftheta = reinterpret_cast<float *>(blk + sz_q + sz_t + 2 * sz_c);
            int *d_fail2 = reinterpret_cast<int *>(blk + sz_q + sz_t + 2 * sz_c + sz_n);
            float *d_fq2 = reinterpret_cast<float *>(blk + sz_q + sz_t + 2 * sz_c + 2 * sz_n);
            int32_t *d_fc2 = reinterpret_cast<int32_t *>(blk + 2 * sz_q + sz_t + 2 * sz_c + 2 * sz_n);
            float *d_fs2 = reinterpret_cast<float *>(blk + 2 * sz_q + sz_t + 3 * sz_c + 2 * sz_n);
            uint4 *d_fqt2 = reinterpret_cast<uint4 *>(blk + 2 * sz_q + sz_t + 4 * sz_c + 2 * sz_n);
            gather_rows_kernel<<<(n_fail * (ED / 4) + 255) / 256, 256, 0, st>>>(d_q + q0 * ED, d_fail, n_fail, d_fq);
            FWAV_LAUNCH_CHECK(ctx);
            if ((rc = pack(d_fq, n_fail, fp * 2, d_fqt, 1, 0))) return rc;
            {
                // The exact kernels are expensive for a handful of queries (list kernel: ~3 ms for 40 queries of
                // config 2; FFMA for top_k > 32: milliseconds per query on a large table).  Second chance on the
                // tensor cores first: the failed queries alone, full split (slack 4e-6 instead of 2e-3), thresholds
                // as before or lowered to the K-th score found (boundary cases), the table split between up to
                // sixteen CTAs per 128 queries, each with its own candidate buffers: many times the room per query.
                // What fails again goes to the list kernel (top_k <= 32) or the FFMA kernel.
                const long long fg = (n_fail + kQTile - 1) / kQTile;
                long long rs = ctx->num_sms / fg;
                if (rs > 16) rs = 16;
                if (rs > n_stages / 8) rs = n_stages / 8;
                if (rs < 1) rs = 1;
                // (rs == 1 with a full-split first pass still pays: boundary cases -- the bulk of the failures on crowded
                // scores, 2 % of config 4's queries -- only need their threshold lowered, not more room; without the
                // second pass they would all go to the FFMA kernel: 477 ms per million queries of the config-4 shape)
                const bool retry = single && !(mode_env && !strcmp(mode_env, "noretry"));
                // The fp16 second chance lowers a boundary case's threshold by 1.25 x ITS bound (5e-3 for two unit heads),
                // which on crowded scores multiplies the candidates (a 15-minute signal with the first pass's 256 entries
                // per column group: 11-20 % of the retried queries overflowed): it gets at least 4 096 entries per query,
                // from the table split where there are few failures, from larger buffers where there are many.
                // (Degenerate data -- pure tones, digital silence: most of a batch tied within the bound -- would ask
                // for gigabytes of such buffers: past 2 GB the full split and its tight margin take over, as before.)
                const bool retry16 = retry16_ok &&
                                     (rs >= 4 || (size_t)n_fail * 4096 * sizeof(int32_t) <= ((size_t)2 << 30));
                int n_fail2 = n_fail;
                const int *d_list2 = nullptr;      // FFMA input rows: indices into the gathered table (nullptr: all of it)
                if (retry) {
                    int rcap = rs > 8 ? collect_cap / 4 : rs >= 2 ? collect_cap / 2 : collect_cap;
                    if (retry16 && rs < 4) rcap = ((int)(4096 / (4 * rs)) + 1) & ~1;        // 1 024 / 512 / 342 per column group
                    // (a part of a finely split table still sees a whole cluster of neighbouring domains: with 64 entries
                    // per part the retried queries of an eighth of config 2 -- 16 CTAs per 128 of them -- overflowed single
                    // parts and paid for the list kernel, 0.8-2 ms on the slowest of 8 ranks)
                    if (retry16 && rcap < 128) rcap = 128;
                    int32_t *d_rbuf = nullptr;
                    const size_t nb = (size_t)n_fail * rs * 4 * rcap * sizeof(int32_t), nc = (size_t)n_fail * rs * 4 * sizeof(int);
                    if ((rc = fwav_ws_reserve(ctx, WS_UMMA_TAIL, nb + nc, (void **)&d_rbuf))) return rc;
                    gather_f32_kernel<<<(n_fail + 255) / 256, 256, 0, st>>>((retry16 ? d_theta_hi : d_theta) + q0, d_fail, n_fail, d_ftheta);
                    FWAV_LAUNCH_CHECK(ctx);
                    FWAV_CUDA(ctx, cudaMemsetAsync(d_fail2 + n_fail, 0, 4 * sizeof(int), st));
                    ScanArgs ar = a;
                    ar.q_tiles = d_fqt; ar.Q = d_fq; ar.n_q = n_fail; ar.active = nullptr; ar.theta = d_ftheta;
                    ar.n_split = (int)rs; ar.cbuf = d_rbuf; ar.cap = rcap;
                    ar.ccount = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(d_rbuf) + nb);
                    if (retry16) {
                        if (ar.front) FWAV_CUDA(ctx, cudaMemsetAsync(d_front, 0, 16 * sizeof(int), st));      // rs <= 16
                        rc = launch_collect_hi(ctx, ar, fg, rs, st);
                    } else {
                        rc = compact ? launch_scan<MODE_COLLECT, false, 1, true>(ctx, ar, fg, rs, st)
                                     : launch_scan<MODE_COLLECT, false, 1>(ctx, ar, fg, rs, st);
                    }
                    if (rc) return rc;
                    const int parts = 4 * (int)rs;
                    const int key_cap2 = parts * rcap < 4096 ? parts * rcap : 4096;      // keys per query the verification holds (more: an overflow)
                    const size_t fin_smem = (size_t)kFinWarps * key_cap2 * sizeof(unsigned long long);
                    if (fin_smem > 48 * 1024)
                        FWAV_CUDA(ctx, cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
                    finalize_kernel<<<(unsigned)((n_fail + kFinWarps - 1) / kFinWarps), kFinWarps * 32, fin_smem, st>>>(
                        d_fq, d_emb, n_fail, n_d, top_k, nullptr, d_ftheta, ar.cbuf, ar.ccount, rcap, parts, 0, d_norms,
                        retry16 ? 2 : 0, d_fc, d_fs, d_fail2, d_fail2 + n_fail, nullptr, key_cap2, nullptr);
                    FWAV_LAUNCH_CHECK(ctx);
                    int h_fail2[4] = {0, 0, 0, 0};
                    FWAV_CUDA(ctx, cudaMemcpyAsync(h_fail2, d_fail2 + n_fail, sizeof h_fail2, cudaMemcpyDeviceToHost, st));
                    FWAV_CUDA(ctx, cudaStreamSynchronize(st));
                    n_fail2 = h_fail2[0];
                    d_list2 = d_fail2;
                    if (getenv("FWAV_UMMA_VERBOSE"))
                        fprintf(stderr, "[fwav] search batch at %lld: second chance (%s, table split %lld ways): %d of %d fail again (overflow %d, short %d, boundary %d)\n",
                                q0, retry16 ? "hi*hi, fp16 accumulators" : "full split", rs, n_fail2, n_fail, h_fail2[1], h_fail2[2], h_fail2[3]);
                }
                ctx->umma_ffma_queries += n_fail2;     // (the list kernel's, for top_k <= 32)
                if (n_fail2 > 0) {
                    const float *d_in = d_fq;
                    if (d_list2) {
                        gather_rows_kernel<<<(n_fail2 * (ED / 4) + 255) / 256, 256, 0, st>>>(d_fq, d_list2, n_fail2, d_fq2);
                        FWAV_LAUNCH_CHECK(ctx);
                        d_in = d_fq2;
                    }
                    if (top_k <= 32) {
                        const uint4 *d_in_t = d_fqt;
                        if (d_list2) {
                            const long long fp2 = (n_fail2 + kQPair - 1) / kQPair;
                            if ((rc = pack(d_fq2, n_fail2, fp2 * 2, d_fqt2, 1, 0))) return rc;
                            d_in_t = d_fqt2;
                        }
                        rc = launch_lists(ctx, d_in_t, d_et, d_in, d_emb, n_fail2, n_d, (int)n_stages, top_k, nullptr,
                                          d_list2 ? d_fc2 : d_fc, d_list2 ? d_fs2 : d_fs, dbg, st, d_norms, compact);
                    } else {
                        rc = fwav_launch_topk_ffma(ctx, d_in, n_fail2, d_emb, n_d, ED, top_k, nullptr, d_list2 ? d_fc2 : d_fc,
                                                   d_list2 ? d_fs2 : d_fs, st);
                    }
                    if (rc) return rc;
                    if (d_list2) {
                        scatter_cand_kernel<<<(n_fail2 * top_k + 255) / 256, 256, 0, st>>>(d_fc2, d_fs2, d_list2, n_fail2, top_k, d_fc, d_fs);
                        FWAV_LAUNCH_CHECK(ctx);
                    }
                }
            }
            scatter_cand_kernel<<<(n_fail * top_k + 255) / 256, 256, 0, st>>>(d_fc, d_fs, d_fail, n_fail, top_k,
                                                                               d_cand + q0 * top_k,
                                                                               d_scores ? d_scores + q0 * top_k : nullptr);
            FWAV_LAUNCH_CHECK(ctx);
            FWAV_CUDA(ctx, cudaMemsetAsync(d_fail_count, 0, 4 * sizeof(int), st));
        }
        if ((rc = mark(ctx, slot, 5, st))) return rc;
        if (slot < fwav_ctx::kSearchSlots) ctx->search_slots_used = slot + 1;
    }
    return FWAV_OK;
}

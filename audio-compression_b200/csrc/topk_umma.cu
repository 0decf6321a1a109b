// topk_umma.cu — A4 on the 5th-generation tensor cores: exact cosine-similarity
// candidate search as a fused split-fp16 tcgen05 contraction with an in-kernel
// threshold top-K (replaces range_candidates_from_embedding_emb,
// /root/reference/fractal.py:535-552, for all ranges at once).
//
// Shape of the work: scores = Q (n_q x 16) . E^T (16 x n_d), n_d ~ 2e6..4e7, top
// 32 per row.  K = 16 is tiny, so the contraction is "all epilogue": every score
// has to be looked at once, and a tcgen05.mma with so little K is bound by its
// per-instruction cost, not by the tensor array (scripts/umma_microbench.cu on B200:
// 118 cycles for M=128 N=128 K=16 whatever the operand layout, 173 for the CTA-pair
// M=256 N=256 one that does four times the work).  The design therefore uses the
// largest instruction there is and keeps the tensor pipe and the epilogue warps
// both busy without ever writing a score to memory:
//
//   * a CLUSTER OF TWO CTAs (one TPC) owns 256 queries, 128 per CTA, and issues
//     tcgen05.mma.cta_group::2 with M=256, N=256: each CTA keeps its own 128 query
//     rows (A) and HALF of every 256-domain stage (B) in shared memory, so every
//     domain tile is read from L2 once per 256 queries;
//   * operands are pre-split into fp16 hi/lo parts (x = hi + lo, 22 significant
//     bits) and pre-tiled by a pack kernel into the K-major SWIZZLE_32B UMMA layout,
//     so one plain bulk copy (cp.async.bulk, the TMA engine, no tensor map) lands a
//     128-domain tile in shared memory ready for the tensor core;
//   * per stage the leader CTA's elected thread issues three K=16 instructions
//     (hi*lo, lo*hi, hi*hi) into one of two 256-column TMEM accumulator buffers and
//     commits to mbarriers in BOTH CTAs (multicast); the peer CTA relays "my half
//     of the stage has landed" to the leader with a remote mbarrier arrive;
//   * 8 epilogue warps per CTA (two per TMEM lane quadrant, 128 columns each) read
//     the accumulators with tcgen05.ld 32x32b.x32 — one query row per thread — in two
//     halves so that the second half's loads overlap the first half's reduction,
//     hand the buffer back to the MMA thread as soon as it is in registers, reduce
//     each 32-column chunk to its maximum with 3-input max instructions and compare
//     it with the row's running threshold; only chunks that beat it take the
//     warp-cooperative insertion path into the row's candidate list in shared memory
//     (two warps share a row's list, under a per-quadrant lock);
//   * a row keeps its 48 best candidates (top_k <= 32 plus a 16-entry margin) as
//     sorted 64-bit keys; at the end every kept candidate is re-scored with the
//     canonical float32 FMA chain and the best top_k are written best-first, so the
//     result equals the FFMA kernel's unless more than 16 domains tie with the K-th
//     score at the split-fp16 rounding level (~1e-6).
//
// Bound: tensor pipe, 2*16 algorithmic flop per pair (the split issues 3x that);
// see DESIGN.md for the epilogue budget.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "fwav_math.cuh"

namespace {

constexpr int ED = 16;                 // embedding dim this kernel is built for
constexpr int kQTile = 128;            // queries per CTA (UMMA M = 256 across the CTA pair)
constexpr int kQPair = 2 * kQTile;     // queries per cluster
constexpr int kDTile = 128;            // domain rows per packed tile = one CTA's half of a stage
constexpr int kDStage = 2 * kDTile;    // domains per stage (UMMA N)
constexpr int kStages = 8;             // 64 KB of domain tiles in flight per CTA (power of two)
constexpr int kThreads = 352;          // warps 0-7 epilogue, 8 producer + TMEM alloc, 9-10 MMA issuers (leader) / 9 relay (peer)
constexpr int kChunks = 4;             // 32-column TMEM chunks per warp and stage (128 columns)
constexpr int kKeep = 48;              // candidates kept per query (top_k <= 32 plus a 16-entry margin)
constexpr int kCap = kKeep;            // eight-byte keys per query row in shared memory
constexpr uint32_t kPartBytes = kDTile * ED * 2;       // 4 KB: one 128-row hi or lo tile (fp16)
constexpr uint32_t kTileBytes = 2 * kPartBytes;        // 8 KB: hi | lo
constexpr uint32_t kSBO = 256;                         // bytes between 8-row groups (8 rows x 32 B)
constexpr unsigned kFull = 0xffffffffu;

// shared memory map (dynamic, 1024-aligned), identical in both CTAs of a pair
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + kTileBytes;
constexpr uint32_t kOffList = kOffB + kStages * kTileBytes;
constexpr uint32_t kOffScratch = kOffList + kQTile * 2 * kCap * 8;   // [row][column half][kCap] keys; then one owner's 128 scores per warp
constexpr uint32_t kOffBars = kOffScratch + 8 * 128 * 4;
constexpr uint32_t kBarBytes = 16 * kStages + 72;      // full[], empty[], tfull[2], tempty[2], a, done[2], tmem slot
constexpr uint32_t kSmemBytes = kOffBars + kBarBytes;

// UMMA instruction descriptor: D=F32, A=B=F16, both K-major, N=256, M=256 (cta_group::2)
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kDStage >> 3) << 17) | ((uint32_t)(kQPair >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
// completion of every MMA issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
// K-major SWIZZLE_32B matrix descriptor (cute::UMMA::SmemDescriptor, version 1): 8-row groups
// kSBO bytes apart, the two 16-byte K chunks of a row adjacent (leading offset 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(kSBO >> 4) << 32) |
           (1ull << 46) | (6ull << 61);
}
// D[tmem] (+)= A[smem] . B[smem]^T over the CTA pair, fp16 operands, K = 16, f32 accumulate
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_wait_ld2(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : FWAV_RW32(a), FWAV_RW32(b)::"memory");
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
                      uint4 *__restrict__ dst) {
    const long long total = n_tiles * (kDTile * 2);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long tile = g / (kDTile * 2);
        const int w = (int)(g - tile * (kDTile * 2));
        const int c = w / kDTile, r = w % kDTile;       // c: which 8-element K chunk
        const long long row = tile * kDTile + r;
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
topk_umma_kernel(const uint4 *__restrict__ q_tiles, const uint4 *__restrict__ e_tiles,
                 const float *__restrict__ Q, const float *__restrict__ E, long long n_q, long long n_d,
                 int n_stages, int top_k, const uint8_t *__restrict__ active, int32_t *__restrict__ cand,
                 float *__restrict__ scores, int dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const long long pair_base = (long long)(blockIdx.x >> 1) * kQPair;     // first query of the pair
    const long long q_base = pair_base + (long long)cta_rank * kQTile;    // first query of this CTA
    unsigned long long *rows = reinterpret_cast<unsigned long long *>(smem + kOffList);
    const uint32_t bars = smem_u32(smem + kOffBars);
    // barrier slots (8 bytes each)
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tfull = bars + 16 * kStages,
                   bar_tempty = bar_tfull + 16, bar_a = bar_tempty + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffBars + kBarBytes - 8);

    // whole-pair early out (energy-pruned stretch): every row gets -1.  Both CTAs look at all
    // 256 rows so that they take the same decision.
    {
        int any = 0;
        for (int i = threadIdx.x; i < kQPair; i += kThreads) {
            const long long q = pair_base + i;
            if (q < n_q && (!active || active[q])) any = 1;
        }
        if (!__syncthreads_or(any)) {
            for (int i = threadIdx.x; i < kQTile * top_k; i += kThreads) {
                const long long q = q_base + i / top_k;
                if (q < n_q) {
                    cand[q * top_k + i % top_k] = -1;
                    if (scores) scores[q * top_k + i % top_k] = -INFINITY;
                }
            }
            return;
        }
    }

    for (int i = threadIdx.x; i < kQTile * 2 * kCap; i += kThreads) rows[i] = empty_key(i % kCap);
    if (threadIdx.x == 0) {
        // the leader's "stage has landed" barriers collect its own copy and the peer's relay
        const uint32_t n_land = cta_rank == 0 ? 2u : 1u;
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, n_land); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 16); }
        mbar_init(bar_a, n_land);
        mbar_init(bar_a + 8, 1);
        mbar_init(bar_a + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();           // barriers of both CTAs are initialised before anyone arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Stages are visited starting at the pair's own rows and wrapping around: when the
    // queries are rows of the same table (the reference's aliasing) their best matches
    // sit next to them, so the thresholds tighten within the first few stages.
    // Neighbouring pairs start one stage apart, so they still share every stage in L2.
    const int t_first = (int)((pair_base / kDStage) % n_stages);

    if (warp == 8) {
        // ===== producer: bulk copies (TMA engine) of this CTA's query tile and its half of every stage =====
        if (lane == 0) {
            mbar_expect_tx(bar_a, kTileBytes);
            bulk_g2s(smem_u32(smem + kOffA), q_tiles + ((long long)blockIdx.x) * (kTileBytes / 16), kTileBytes, bar_a);
            int tt = t_first;
            for (int t = 0; t < n_stages; ++t) {
                const int s = t & (kStages - 1);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * s, kTileBytes);
                bulk_g2s(smem_u32(smem + kOffB + s * kTileBytes),
                         e_tiles + (2ll * tt + cta_rank) * (kTileBytes / 16), kTileBytes, bar_full + 8 * s);
                if (++tt == n_stages) tt = 0;
            }
        }
    } else if (cta_rank != 0 && warp >= 9) {
        // ===== peer CTA: tell the leader when this CTA's operands have landed =====
        if (warp == 9 && lane == 0) {
            mbar_wait(bar_a, 0);
            mbar_arrive_remote(bar_a, 0);
            for (int t = 0; t < n_stages; ++t) {
                const int s = t & (kStages - 1);
                mbar_wait(bar_full + 8 * s, (uint32_t)((t / kStages) & 1));
                mbar_arrive_remote(bar_full + 8 * s, 0);
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ===== leader CTA: two MMA issuers, one thread each, for the pair =====
        // One stage costs the issuing thread ~600 cycles of serial latency (two mbarrier waits,
        // three tcgen05.mma, two commits) against ~520 cycles of tensor-pipe time, so the stages
        // alternate between two threads: warp 9 owns the even ones (TMEM buffer 0), warp 10 the odd
        // ones (buffer 1).  The buffers and shared-memory stages are disjoint and a commit covers
        // its own thread's MMAs, so the two instruction streams need no ordering between them.
        if (lane == 0) {
            mbar_wait(bar_a, 0);
            const uint32_t a_hi = smem_u32(smem + kOffA), a_lo = a_hi + kPartBytes;
            const uint64_t da_hi = smem_desc(a_hi), da_lo = smem_desc(a_lo);
            const int buf = warp - 9;
            const uint32_t d = tmem_base + (uint32_t)(buf * kDStage);
            for (int t = buf; t < n_stages; t += 2) {
                const int s = t & (kStages - 1);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                const uint32_t tph = (uint32_t)((t >> 1) & 1);
                if (!(dbg & 8)) mbar_wait(bar_tempty + 8 * buf, tph ^ 1);    // dbg 8: free-running MMA (profiling)
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(smem + kOffB + s * kTileBytes), b_lo = b_hi + kPartBytes;
                const uint64_t db_hi = smem_desc(b_hi), db_lo = smem_desc(b_lo);
                // small cross terms first, the hi*hi term last; one K=16 instruction each
                umma_f16_pair(d, da_hi, db_lo, 0);
                umma_f16_pair(d, da_lo, db_hi, 1);
                umma_f16_pair(d, da_hi, db_hi, 1);
                umma_commit_pair(bar_empty + 8 * s);        // stage free (both CTAs) once these MMAs have read it
                umma_commit_pair(bar_tfull + 8 * buf);      // accumulators ready for the epilogue warps of both CTAs
            }
            if (dbg & 8) {   // free-running profiling mode: drain the tensor pipe before leaving
                umma_commit_pair(bar_a + 8 + 8 * buf);
                mbar_wait(bar_a + 8 + 8 * buf, 0);
            }
        }
    } else if (warp < 8) {
        // ===== epilogue: one query row per thread, 128 of the stage's 256 columns per warp =====
        const int quad = warp & 3, half = warp >> 2;
        const int row0 = quad * 32;                       // first row of this warp inside the CTA tile
        const long long q = q_base + row0 + lane;
        float tau = (q < n_q && (!active || active[q]) && !(dbg & 4)) ? -INFINITY : INFINITY;   // +inf: never a candidate
        unsigned long long *lists = rows + ((size_t)row0 * 2 + half) * kCap;     // row r of the warp: lists + r * 2 * kCap
        uint32_t *scratch = reinterpret_cast<uint32_t *>(smem + kOffScratch) + warp * 128;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
        int tt = t_first;
        for (int t = 0; t < ((dbg & 8) ? 0 : n_stages); ++t) {
            const int buf = t & 1;
            const uint32_t tph = (uint32_t)((t >> 1) & 1);
            mbar_wait(bar_tfull + 8 * buf, tph);
            tc_fence_after();
            const uint32_t ta = t_lane + (uint32_t)(buf * kDStage);
            uint32_t v[kChunks][32];
            unsigned hits = 0;
            if (!(dbg & 2)) {
                tmem_ld32(ta, v[0]);
                tmem_ld32(ta + 32, v[1]);
                tmem_wait_ld2(v[0], v[1]);
                tmem_ld32(ta + 64, v[2]);        // in flight while the first two chunks are reduced
                tmem_ld32(ta + 96, v[3]);
                hits = (chunk_max(v[0]) > tau ? 1u : 0u) | (chunk_max(v[1]) > tau ? 2u : 0u);
                tmem_wait_ld2(v[2], v[3]);
            }
            // the whole stage sits in registers: hand the TMEM buffer back before looking at the rest
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);
            const long long base = (long long)tt * kDStage + half * 128;
            if (++tt == n_stages) tt = 0;
            if (dbg & 3) continue;
            hits |= (chunk_max(v[2]) > tau ? 4u : 0u) | (chunk_max(v[3]) > tau ? 8u : 0u);
            if (__any_sync(kFull, hits != 0)) absorb_stage(v, hits, tau, base, n_d, lists, scratch, lane);
        }
        // both warps of a quadrant are done with their lists before the rows are finalised
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // ---- exact float32 re-score of the 2 x kKeep kept candidates of a row and best-first write-out ----
        constexpr int kPerLane = 2 * kKeep / 32;
        static_assert(2 * kKeep % 32 == 0, "final merge handles whole warps of keys");
        for (int r = half * 16; r < half * 16 + 16; ++r) {
            const long long qq = q_base + row0 + r;
            if (qq >= n_q) break;
            unsigned long long *keys = rows + (size_t)(row0 + r) * 2 * kCap;      // both column halves, contiguous
            const float *qv = Q + qq * ED;
            // canonical score (ascending-k float32 FMA chain) of every kept candidate
            unsigned long long k[kPerLane];
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) {
                k[i] = keys[lane + 32 * i];
                if ((uint32_t)(k[i] >> 32) != kNegInfBits) {
                    const int id = (int)(0xFFFFFFFFu - (uint32_t)k[i]);
                    k[i] = make_key(fwm::score_chain(qv, E + (long long)id * ED, ED), id);
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
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) {
                if (rk[i] < top_k) {
                    const bool live = (uint32_t)(k[i] >> 32) != kNegInfBits;
                    cand[qq * top_k + rk[i]] = live ? (int)(0xFFFFFFFFu - (uint32_t)k[i]) : -1;
                    if (scores) scores[qq * top_k + rk[i]] = live ? unorder_bits((uint32_t)(k[i] >> 32)) : -INFINITY;
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the other may still signal it
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

bool fwav_topk_umma_supported(int emb_dim, int top_k, int64_t n_q, int64_t n_d) {
    return emb_dim == ED && top_k >= 1 && top_k <= 32 && n_q > 0 && n_d > 0;
}

int fwav_launch_topk_umma(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb, int64_t n_d,
                          int emb_dim, int top_k, const uint8_t *d_active, int32_t *d_cand, float *d_scores,
                          cudaStream_t st) {
    FWAV_REQUIRE(ctx, fwav_topk_umma_supported(emb_dim, top_k, n_q, n_d),
                 "tensor-core search is built for emb_dim=16 and top_k<=32 (got %d, %d)", emb_dim, top_k);
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
    auto grid_for = [&](long long work) {
        long long need = (work + 255) / 256, cap = (long long)ctx->num_sms * 8;
        return (int)(need < cap ? need : cap);
    };
    pack_f16_tiles_kernel<<<grid_for(e_tiles * kDTile * 2), 256, 0, st>>>(d_emb, n_d, e_tiles, d_et);
    FWAV_LAUNCH_CHECK(ctx);
    pack_f16_tiles_kernel<<<grid_for(q_tiles * kDTile * 2), 256, 0, st>>>(d_q, n_q, q_tiles, d_qt);
    FWAV_LAUNCH_CHECK(ctx);
    FWAV_CUDA(ctx, cudaFuncSetAttribute(topk_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    const char *dbg_env = getenv("FWAV_UMMA_DEBUG");   // profiling aid (results are wrong when set)
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    topk_umma_kernel<<<(unsigned)(2 * q_pairs), kThreads, kSmemBytes, st>>>(d_qt, d_et, d_q, d_emb, n_q, n_d, (int)n_stages,
                                                                           top_k, d_active, d_cand, d_scores, dbg);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

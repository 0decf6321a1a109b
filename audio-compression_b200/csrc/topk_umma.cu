// topk_umma.cu — A4 on the 5th-generation tensor cores: exact cosine-similarity
// candidate search as a fused 3xTF32 tcgen05 contraction with an in-kernel
// threshold top-K (replaces range_candidates_from_embedding_emb,
// /root/reference/fractal.py:535-552, for all ranges at once).
//
// Shape of the work: scores = Q (n_q x 16) . E^T (16 x n_d), n_d ~ 2e6..4e7, top
// 32 per row.  K = 16 is tiny, so the contraction is "all epilogue": every score
// has to be looked at once.  The design therefore keeps the tensor pipe and the
// epilogue warps both busy and never writes a score to memory:
//
//   * operands are pre-split into TF32 hi/lo parts and pre-tiled by a pack kernel
//     into the canonical K-major no-swizzle UMMA layout ([16-byte K chunk][row]),
//     so one plain bulk copy (cp.async.bulk, the TMA engine, no tensor map) lands
//     a 128-domain stage in shared memory ready for tcgen05.mma;
//   * a CTA owns 256 queries (two M=128 operand tiles, resident in shared memory)
//     and streams all domains in 128-row stages through a 3-deep mbarrier ring;
//   * one elected thread issues 2 x 6 tcgen05.mma.kind::tf32 (lo*hi, hi*lo, hi*hi;
//     K = 8 per instruction) per stage into TMEM; the accumulators are
//     double-buffered (2 halves x 2 buffers x 128 columns = all 512 columns);
//   * 8 epilogue warps (2 per TMEM lane quadrant) read the accumulators with
//     tcgen05.ld 32x32b.x32 — one query row per thread — reduce each 32-column
//     chunk to its maximum with 3-input max instructions and compare it with the
//     row's running threshold; only chunks that beat it take the warp-cooperative
//     insertion path into the row's candidate list in shared memory;
//   * a row keeps its 48 best candidates (top_k <= 32 plus a 16-entry margin) as
//     sorted 64-bit keys; an insertion is one ballot for the position and a
//     one-slot shift, and the accumulator buffer is handed back to the MMA warp
//     BEFORE a stage is examined, so a warp in the rare path does not stall the
//     tensor pipe; at the end every kept candidate is
//     re-scored with the canonical float32 FMA chain and the best top_k are written
//     best-first, so the result equals the FFMA kernel's unless more than 16
//     domains tie with the K-th score at the 3xTF32 rounding level (~1e-6).
//
// Bound: tensor pipe (TF32), 2*16 algorithmic flop per pair (3xTF32 issues 3x
// that); see DESIGN.md for the epilogue budget.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "fwav_math.cuh"

namespace {

constexpr int ED = 16;                 // embedding dim this kernel is built for
constexpr int kQTile = 256;            // queries per CTA (2 x M=128)
constexpr int kDTile = 128;            // domains per stage (UMMA N)
constexpr int kStages = 12;           // 96 KB of domain stages in flight: the ring has to cover the L2/HBM latency
constexpr int kThreads = 320;          // 8 epilogue warps + producer/alloc warp + MMA warp
constexpr int kChunks = kDTile / 32;   // 32-column TMEM chunks per stage
constexpr int kKeep = 48;              // candidates kept per query (top_k <= 32 plus a 16-entry margin)
constexpr int kCap = kKeep;            // eight-byte keys per query row in shared memory
constexpr uint32_t kPartBytes = kDTile * ED * 2;       // 4 KB: one 128-row hi or lo tile (fp16)
constexpr uint32_t kTileBytes = 2 * kPartBytes;        // 8 KB: hi | lo
constexpr uint32_t kABytes = 2 * kTileBytes;           // 16 KB: two query halves
constexpr uint32_t kLBO = kDTile * 16;                 // bytes between the two 16-byte K chunks
constexpr uint32_t kSBO = 256;                         // bytes between 8-row groups (8 rows x 32 B)
constexpr unsigned kFull = 0xffffffffu;

// shared memory map (dynamic, 1024-aligned)
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + kABytes;
constexpr uint32_t kOffList = kOffB + kStages * kTileBytes;
constexpr uint32_t kOffScratch = kOffList + kQTile * kCap * 8;       // one owner's stage of scores per warp
constexpr uint32_t kOffBars = kOffScratch + 8 * kDTile * 4;
constexpr uint32_t kBarBytes = 16 * kStages + 64;      // full[], empty[], tfull[2], tempty[2], a, done, tmem slot
constexpr uint32_t kSmemBytes = kOffBars + kBarBytes;

// UMMA instruction descriptor: D=F32, A=B=F16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((kDTile >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_32B matrix descriptor (cute::UMMA::SmemDescriptor, version 1): 8-row groups
// kSBO bytes apart, the two 16-byte K chunks of a row adjacent (leading offset 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(kSBO >> 4) << 32) |
           (1ull << 46) | (6ull << 61);
}
// D[tmem] (+)= A[smem] . B[smem]^T, fp16 operands, K = 16 per instruction, f32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : FWAV_RW32(v)::"memory");
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

// Warp-cooperative sorted insertion of x (known to beat the row's last key): a
// ballot gives its position, every lane moves its two keys one slot down, the
// last key falls off.  ~25 instructions, no loops.  Returns the new threshold.
__device__ __forceinline__ float insert_sorted(unsigned long long *keys, unsigned long long x, int lane) {
    const bool has1 = lane + 32 < kKeep;
    const unsigned long long k0 = keys[lane];
    const unsigned long long k1 = has1 ? keys[lane + 32] : 0ull;
    const int pos = __popc(__ballot_sync(kFull, k0 > x)) + __popc(__ballot_sync(kFull, has1 && k1 > x));
    if (lane >= pos && lane + 1 < kKeep) keys[lane + 1] = k0;
    if (has1 && lane + 32 >= pos && lane + 33 < kKeep) keys[lane + 33] = k1;
    if (lane == 0) keys[pos] = x;
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

// Rare path, once per stage and warp: rows whose threshold was beaten ("owners")
// are served one at a time.  The owner spills the chunks that hit to the warp's
// scratch line; then all 32 lanes test one column each and the passing columns
// are inserted in index order.
__device__ __forceinline__ void absorb_stage(const uint32_t (&v)[kChunks][32], unsigned hits, float &tau,
                                             long long base, long long n_d, unsigned long long *rows, int row0,
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
        unsigned long long *keys = rows + (size_t)(row0 + bl) * kCap;
        for (int c = 0; c < kChunks; ++c) {
            if (!(hb >> c & 1)) continue;
            const float x = __uint_as_float(scratch[32 * c + lane]);
            const long long id0 = base + 32 * c;
            unsigned pm = __ballot_sync(kFull, x > tb && id0 + lane < n_d);
            while (pm) {
                const int j = __ffs(pm) - 1;
                pm &= pm - 1;
                const float xs = __shfl_sync(kFull, x, j);
                if (xs > tb) tb = insert_sorted(keys, make_key(xs, (int)(id0 + j)), lane);
            }
        }
        if (lane == bl) tau = tb;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kThreads, 1)
topk_umma_kernel(const uint4 *__restrict__ q_tiles, const uint4 *__restrict__ e_tiles,
                 const float *__restrict__ Q, const float *__restrict__ E, long long n_q, long long n_d,
                 int top_k, const uint8_t *__restrict__ active, int32_t *__restrict__ cand,
                 float *__restrict__ scores, int dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q_base = (long long)blockIdx.x * kQTile;
    unsigned long long *rows = reinterpret_cast<unsigned long long *>(smem + kOffList);
    const uint32_t bars = smem_u32(smem + kOffBars);
    // barrier slots (8 bytes each)
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_tfull = bars + 16 * kStages,
                   bar_tempty = bar_tfull + 16, bar_a = bar_tempty + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffBars + kBarBytes - 8);

    // whole-CTA early out (energy-pruned stretch): every row gets -1
    {
        int any = 0;
        for (int i = threadIdx.x; i < kQTile; i += kThreads) {
            const long long q = q_base + i;
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

    for (int i = threadIdx.x; i < kQTile * kCap; i += kThreads) rows[i] = empty_key(i % kCap);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 8); }
        mbar_init(bar_a, 1);
        mbar_init(bar_a + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long n_tiles = (n_d + kDTile - 1) / kDTile;
    // Stages are visited starting at the CTA's own rows and wrapping around: when the
    // queries are rows of the same table (the reference's aliasing) their best matches
    // sit next to them, so the thresholds tighten within the first few stages.
    // Neighbouring CTAs start two stages apart, so they still share every stage in L2.
    const long long t_first = (q_base / kDTile) % n_tiles;

    if (warp == 8) {
        // ===== producer: bulk copies (TMA engine) =====
        if (lane == 0) {
            mbar_expect_tx(bar_a, kABytes);
            bulk_g2s(smem_u32(smem + kOffA), q_tiles + (long long)blockIdx.x * (kABytes / 16), kABytes, bar_a);
            for (long long t = 0; t < n_tiles; ++t) {
                const int s = (int)(t % kStages);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * s, kTileBytes);
                const long long tt = (t + t_first) % n_tiles;
                bulk_g2s(smem_u32(smem + kOffB + s * kTileBytes), e_tiles + tt * (kTileBytes / 16), kTileBytes,
                         bar_full + 8 * s);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            mbar_wait(bar_a, 0);
            const uint32_t a_addr = smem_u32(smem + kOffA);
            for (long long t = 0; t < n_tiles; ++t) {
                const int s = (int)(t % kStages);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                const int buf = (int)(t & 1);
                const uint32_t tph = (uint32_t)((t >> 1) & 1);
                if (!(dbg & 8)) mbar_wait(bar_tempty + 8 * buf, tph ^ 1);    // dbg 8: free-running MMA (profiling)
                if (!(dbg & 16)) mbar_wait(bar_full + 8 * s, ph);            // dbg 16: no producer (profiling)
                tc_fence_after();
                const uint32_t b_hi = smem_u32(smem + kOffB + s * kTileBytes), b_lo = b_hi + kPartBytes;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t a_hi = a_addr + h * kTileBytes, a_lo = a_hi + kPartBytes;
                    const uint32_t d = tmem_base + (uint32_t)(buf * 256 + h * 128);
                    // small cross terms first, the hi*hi term last; one K=16 instruction each
                    umma_f16(d, smem_desc(a_hi), smem_desc(b_lo), 0);
                    umma_f16(d, smem_desc(a_lo), smem_desc(b_hi), 1);
                    umma_f16(d, smem_desc(a_hi), smem_desc(b_hi), 1);
                }
                umma_commit(bar_empty + 8 * s);        // stage free once these MMAs have read it
                umma_commit(bar_tfull + 8 * buf);      // accumulators ready for the epilogue
            }
            if (dbg & 8) {   // free-running profiling mode: drain the tensor pipe before leaving
                umma_commit(bar_a + 8);
                mbar_wait(bar_a + 8, 0);
            }
        }
    } else if (warp < 8) {
        // ===== epilogue: one query row per thread =====
        const int quad = warp & 3, half = warp >> 2;
        const int row0 = half * 128 + quad * 32;          // first row of this warp inside the CTA tile
        const long long q = q_base + row0 + lane;
        float tau = (q < n_q && (!active || active[q])) ? -INFINITY : INFINITY;
        if (dbg & 4) tau = INFINITY;      // profiling aid: fast path only
        uint32_t *scratch = reinterpret_cast<uint32_t *>(smem + kOffScratch) + warp * kDTile;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
        for (long long t = 0; t < ((dbg & 8) ? 0 : n_tiles); ++t) {
            const int buf = (int)(t & 1);
            const uint32_t tph = (uint32_t)((t >> 1) & 1);
            mbar_wait(bar_tfull + 8 * buf, tph);
            tc_fence_after();
            const uint32_t ta = t_lane + (uint32_t)(buf * 256);
            uint32_t v[kChunks][32];
            if (!(dbg & 2)) {
#pragma unroll
                for (int c = 0; c < kChunks; ++c) tmem_ld32(ta + 32 * c, v[c]);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) tmem_wait_ld(v[c]);
            }
            // the whole stage sits in registers: hand the TMEM buffer back before looking at it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            if (dbg & 3) continue;
            unsigned hits = 0;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) hits |= (chunk_max(v[c]) > tau ? 1u : 0u) << c;
            if (__any_sync(kFull, hits != 0))
                absorb_stage(v, hits, tau, ((t + t_first) % n_tiles) * kDTile, n_d, rows, row0, scratch, lane);
        }
        // ---- exact float32 re-score of the kept candidates and best-first write-out ----
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const long long qq = q_base + row0 + r;
            if (qq >= n_q) break;
            unsigned long long *keys = rows + (size_t)(row0 + r) * kCap;
            const float *qv = Q + qq * ED;
            // canonical score (ascending-k float32 FMA chain) of every kept candidate
            unsigned long long k0 = keys[lane], k1 = (lane + 32 < kKeep) ? keys[lane + 32] : 0ull;
            if ((uint32_t)(k0 >> 32) != kNegInfBits) {
                const int id = (int)(0xFFFFFFFFu - (uint32_t)k0);
                k0 = make_key(fwm::score_chain(qv, E + (long long)id * ED, ED), id);
            }
            if (lane + 32 < kKeep && (uint32_t)(k1 >> 32) != kNegInfBits) {
                const int id = (int)(0xFFFFFFFFu - (uint32_t)k1);
                k1 = make_key(fwm::score_chain(qv, E + (long long)id * ED, ED), id);
            }
            __syncwarp();
            keys[lane] = k0;
            if (lane + 32 < kKeep) keys[lane + 32] = k1;
            __syncwarp();
            int r0 = 0, r1 = 0;
            for (int o = 0; o < kKeep; ++o) {
                const unsigned long long ko = keys[o];
                r0 += ko > k0 ? 1 : 0;
                r1 += ko > k1 ? 1 : 0;
            }
            if (r0 < top_k) {
                const bool live = (uint32_t)(k0 >> 32) != kNegInfBits;
                cand[qq * top_k + r0] = live ? (int)(0xFFFFFFFFu - (uint32_t)k0) : -1;
                if (scores) scores[qq * top_k + r0] = live ? unorder_bits((uint32_t)(k0 >> 32)) : -INFINITY;
            }
            if (lane + 32 < kKeep && r1 < top_k) {
                const bool live = (uint32_t)(k1 >> 32) != kNegInfBits;
                cand[qq * top_k + r1] = live ? (int)(0xFFFFFFFFu - (uint32_t)k1) : -1;
                if (scores) scores[qq * top_k + r1] = live ? unorder_bits((uint32_t)(k1 >> 32)) : -INFINITY;
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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
    FWAV_REQUIRE(ctx, n_d < (1ll << 31), "n_domains %lld does not fit the int32 match index", (long long)n_d);
    FWAV_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_emb)) & 15) == 0,
                 "embedding tables must be 16-byte aligned");
    const long long e_tiles = (n_d + kDTile - 1) / kDTile;
    const long long q_ctas = (n_q + kQTile - 1) / kQTile;
    const long long q_tiles = q_ctas * 2;
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
    topk_umma_kernel<<<(unsigned)q_ctas, kThreads, kSmemBytes, st>>>(d_qt, d_et, d_q, d_emb, n_q, n_d, top_k, d_active,
                                                                   d_cand, d_scores, dbg);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

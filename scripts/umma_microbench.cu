// umma_microbench.cu — measures, on one B200, what bounds the skinny (K = 16) similarity
// contraction of csrc/topk_umma.cu: tcgen05.mma issue rate for the operand layouts /
// shapes / CTA-pair modes the kernel could use, tcgen05.ld (TMEM -> registers) rate,
// and the two together.  Data are zeros: only timing is of interest.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_microbench umma_microbench.cu
//   ./umma_microbench <variant>          (one variant per process: a bad encoding only kills itself)
//   ./umma_microbench list
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

struct Params {
    int n_mma;            // MMAs issued by the one MMA thread
    uint32_t idesc;
    uint64_t a_desc_hi;   // descriptor without the start address
    uint64_t b_desc_hi;
    uint32_t a_off[4];    // byte offsets (from the A region) cycled through, n_a a power of two
    int n_a;
    uint32_t b_off[16];
    int n_b;
    int d_cols;           // accumulator buffers at column 0 and d_cols, switched every `per_buf` MMAs
    int per_buf;
    int ts;               // A operand from TMEM (column 496)
    int ld_warps;         // 0..8 warps running tcgen05.ld loops
    int ld_iters;
    int ld_mode;          // 0: loads only, 1: + 3-input max tree and threshold compare
    int sync_mode;        // 1: MMA thread and LDTM warps hand accumulator buffers over through mbarriers
    int tf32;             // 1: kind::tf32 (K = 8 per instruction, same 32-byte operand rows) instead of kind::f16 (K = 16)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// kind::tf32, SS form, one CTA: the same operand bytes are read as 8 tf32 values per row
__device__ __forceinline__ void umma_ss_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

#define R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
               "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
               "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
               "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define RW32(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),  \
                "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),          \
                "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),        \
                "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),        \
                "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : R32(v)
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 64 columns, two 16-bit halves per register (half-precision accumulators)
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ unsigned chunk_max_p(const uint32_t (&v)[32]) {
    unsigned m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = __vimax3_s16x2(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    m[10] = __vimax3_s16x2(v[30], v[31], v[31]);
    const unsigned a = __vimax3_s16x2(m[0], m[1], m[2]), b = __vimax3_s16x2(m[3], m[4], m[5]), c = __vimax3_s16x2(m[6], m[7], m[8]);
    return __vimax3_s16x2(__vimax3_s16x2(a, b, c), m[9], m[10]);
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : RW32(v)::"memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
    float m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        m[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    m[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float a = max3(m[0], m[1], m[2]), b = max3(m[3], m[4], m[5]), c = max3(m[6], m[7], m[8]);
    return max3(max3(a, b, c), m[9], m[10]);
}

constexpr uint32_t kOffA = 0, kOffB = 32768, kOffBars = 32768 + 131072, kSmem = kOffBars + 128;

template <int CG, int PER_BUF, int NLDW>
__global__ void __launch_bounds__((NLDW + 1) * 32, 1) bench_kernel(Params p, long long *out, float tau) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int kThreads = (NLDW + 1) * 32;   // warps 0..NLDW-1: tcgen05.ld loops, last warp: TMEM alloc + MMA thread
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    for (int i = threadIdx.x; i < (int)(kOffBars / 16); i += kThreads) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t bars = smem_u32(smem + kOffBars);
    const uint32_t bar_done = bars, bar_tfull = bars + 8, bar_tempty = bars + 24;
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + kOffBars + 64);
    if (threadIdx.x == 0) {
        mbar_init(bar_done, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, p.ld_warps > 0 ? p.ld_warps : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == NLDW) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *slot;
    long long t0 = 0, t1 = 0;
    if (warp == NLDW) {
        if (cta_rank == 0 && p.n_mma > 0) {     // whole warp, converged: MMAs are issued by an elected lane
            const uint32_t a_base = smem_u32(smem + kOffA), b_base = smem_u32(smem + kOffB);
            // one period = two accumulator buffers; everything the loop needs sits in registers
            constexpr int P = 2 * PER_BUF;
            uint64_t ad[P], bd[P];
            uint32_t dd[P];
#pragma unroll
            for (int j = 0; j < P; ++j) {
                ad[j] = p.a_desc_hi | (uint64_t)(((a_base + p.a_off[j & (p.n_a - 1)]) & 0x3FFFFu) >> 4);
                bd[j] = p.b_desc_hi | (uint64_t)(((b_base + p.b_off[j & (p.n_b - 1)]) & 0x3FFFFu) >> 4);
                dd[j] = tmem_base + (uint32_t)((j / PER_BUF) * p.d_cols) + (PER_BUF == 6 ? (uint32_t)(((j % PER_BUF) / 3) * 128) : 0u);
            }
            const uint32_t a_tm = tmem_base + 496u;
            t0 = clock64();
            uint32_t par = 1;
            for (int i = 0; i < p.n_mma; i += P) {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    if (p.sync_mode && j % PER_BUF == 0) {
                        mbar_wait(bar_tempty + 8 * (j / PER_BUF), par);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t acc = (j % 3) != 0;
                    if (elect_one()) {
                        if (p.tf32) umma_ss_tf32(dd[j], ad[j], bd[j], p.idesc, acc);
                        else if (p.ts) umma_ts<CG>(dd[j], a_tm + 8u * (uint32_t)(j & 1), bd[j], p.idesc, acc);
                        else umma_ss<CG>(dd[j], ad[j], bd[j], p.idesc, acc);
                    }
                    __syncwarp();
                    if (p.sync_mode && j % PER_BUF == PER_BUF - 1) { if (elect_one()) umma_commit<CG>(bar_tfull + 8 * (j / PER_BUF)); __syncwarp(); }
                }
                par ^= 1;
            }
            if (elect_one()) umma_commit<CG>(bar_done);
            __syncwarp();
            mbar_wait(bar_done, 0);
            t1 = clock64();
            if (lane == 0) out[blockIdx.x * 4 + 0] = t1 - t0;
        }
    } else if (warp < p.ld_warps && p.ld_iters > 0) {
        // NLDW == 8: a warp owns 32 rows x 128 columns of each 256-column buffer (4 chunks / iteration)
        // NLDW == 16: 32 rows x 64 columns (2 chunks / iteration)
        constexpr int NCH = NLDW == 8 ? 4 : 2;
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * NCH * 32);
        uint32_t sink = 0;
        t0 = clock64();
        if (p.ld_mode >= 3) {
            // half-precision accumulators: pack::16b loads, 64 columns each (mode 4: + the packed 3-input max tree)
            constexpr int NP = NCH / 2;
            for (int it = 0; it < p.ld_iters; ++it) {
                const uint32_t ta = t_lane + (uint32_t)((it & 1) * 256);
                uint32_t v[NP][32];
#pragma unroll
                for (int c = 0; c < NP; ++c) tmem_ld32_pack16(ta + 64 * c, v[c]);
#pragma unroll
                for (int c = 0; c < NP; ++c) tmem_wait_ld(v[c]);
                if (p.ld_mode == 4) {
#pragma unroll
                    for (int c = 0; c < NP; ++c) sink += chunk_max_p(v[c]) != 0x7fff7fffu ? 1u : 0u;
                } else {
                    sink ^= v[0][0] ^ v[NP - 1][31];
                }
            }
        } else if (p.ld_mode <= 1) {
            for (int it = 0; it < p.ld_iters; ++it) {
                const int buf = it & 1;
                if (p.sync_mode) {
                    mbar_wait(bar_tfull + 8 * buf, (uint32_t)((it >> 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t ta = t_lane + (uint32_t)(buf * 256);
                uint32_t v[NCH][32];
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld32(ta + 32 * c, v[c]);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_wait_ld(v[c]);
                if (p.sync_mode) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
                }
                if (p.ld_mode == 1) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) sink += chunk_max(v[c]) > tau ? 1u : 0u;
                } else {
                    sink ^= v[0][0] ^ v[1][7] ^ v[NCH - 1][31];
                }
            }
        } else {
            // software-pipelined: the loads of the next half are in flight while the current half is reduced
            constexpr int H = NCH / 2;
            uint32_t va[H][32], vb[H][32];
#pragma unroll
            for (int c = 0; c < H; ++c) tmem_ld32(t_lane + 32 * c, va[c]);
            for (int it = 0; it < p.ld_iters; ++it) {
                const uint32_t ta = t_lane + (uint32_t)((it & 1) * 256), tn = t_lane + (uint32_t)(((it + 1) & 1) * 256);
#pragma unroll
                for (int c = 0; c < H; ++c) tmem_wait_ld(va[c]);
#pragma unroll
                for (int c = 0; c < H; ++c) tmem_ld32(ta + 32 * (H + c), vb[c]);
#pragma unroll
                for (int c = 0; c < H; ++c) sink += chunk_max(va[c]) > tau ? 1u : 0u;
#pragma unroll
                for (int c = 0; c < H; ++c) tmem_wait_ld(vb[c]);
#pragma unroll
                for (int c = 0; c < H; ++c) tmem_ld32(tn + 32 * c, va[c]);
#pragma unroll
                for (int c = 0; c < H; ++c) sink += chunk_max(vb[c]) > tau ? 1u : 0u;
            }
#pragma unroll
            for (int c = 0; c < H; ++c) tmem_wait_ld(va[c]);
            sink ^= va[0][0];
        }
        t1 = clock64();
        if (lane == 0) out[blockIdx.x * 4 + 1 + (warp == 0 ? 0 : 1)] = t1 - t0;
        if (sink == 0x12345678u) out[blockIdx.x * 4 + 3] = sink;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == NLDW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host ------------------------------------------------------------------
struct Variant {
    std::string name;
    int cg;
    Params p;
};

static uint64_t desc_hi(int layout, uint32_t lbo, uint32_t sbo) {
    return ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
static uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// operand tile of `rows` rows holding hi and lo K=16 fp16 slices, in the given layout:
//   "sw32": two separate 32-byte-row tiles (hi tile, lo tile)     "none": two separate interleaved tiles
//   "sw64": one 64-byte-row tile [hi|lo]                          "sw128": one 128-byte-row tile [hi|lo|pad]
struct Lay { int layout; uint32_t lbo, sbo, hi_off, lo_off, bytes; };
static Lay lay(const char *kind, int rows) {
    Lay l{};
    if (!strcmp(kind, "sw32")) l = {6, 16, 256, 0, (uint32_t)rows * 32, (uint32_t)rows * 64};
    else if (!strcmp(kind, "none")) l = {0, (uint32_t)rows * 16, 128, 0, (uint32_t)rows * 32, (uint32_t)rows * 64};
    else if (!strcmp(kind, "sw64")) l = {4, 16, 512, 0, 32, (uint32_t)rows * 64};
    else l = {2, 16, 1024, 0, 32, (uint32_t)rows * 128};
    return l;
}

static Variant mk(const char *name, int cg, int M, int N, const char *a_kind, const char *b_kind, int ts, int n_mma,
                  int ld_warps, int ld_iters, int ld_mode, int sync_mode, int n_stages = 4) {
    Variant v;
    v.name = name;
    v.cg = cg;
    Params &p = v.p;
    memset(&p, 0, sizeof p);
    p.n_mma = n_mma;
    p.idesc = idesc(M, N);
    const int a_rows = 128, b_rows = N / cg;
    Lay la = lay(a_kind, a_rows), lb = lay(b_kind, b_rows);
    p.a_desc_hi = desc_hi(la.layout, la.lbo, la.sbo);
    p.b_desc_hi = desc_hi(lb.layout, lb.lbo, lb.sbo);
    // per stage three MMAs: (a_hi, b_lo), (a_lo, b_hi), (a_hi, b_hi); emulate with period-4 / period-(4*stages) cycles
    p.n_a = 4;
    p.a_off[0] = la.hi_off; p.a_off[1] = la.lo_off; p.a_off[2] = la.hi_off; p.a_off[3] = la.lo_off;
    if (lb.bytes * n_stages > 131072) n_stages = 131072 / lb.bytes;
    int nb = 1;
    while (nb * 2 <= n_stages * 2 && nb * 2 <= 16) nb *= 2;
    p.n_b = nb;
    for (int i = 0; i < nb; ++i) p.b_off[i] = (uint32_t)(i / 2) * lb.bytes + ((i & 1) ? lb.hi_off : lb.lo_off);
    p.d_cols = 256;
    p.per_buf = 3 * (256 / (N > 256 ? 256 : N));      // MMAs that fill one 256-column accumulator buffer
    if (p.per_buf < 3) p.per_buf = 3;
    p.ts = ts;
    p.ld_warps = ld_warps;
    p.ld_iters = ld_iters;
    p.ld_mode = ld_mode;
    p.sync_mode = sync_mode;
    return v;
}

int main(int argc, char **argv) {
    std::vector<Variant> vs;
    const int NM = 6144;
    // --- MMA alone, SS form, 1 CTA ---
    vs.push_back(mk("ss_m128n128_sw32", 1, 128, 128, "sw32", "sw32", 0, NM, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n128_none", 1, 128, 128, "none", "none", 0, NM, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n128_sw64", 1, 128, 128, "sw64", "sw64", 0, NM, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n128_sw128", 1, 128, 128, "sw128", "sw128", 0, NM, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n256_sw32", 1, 128, 256, "sw32", "sw32", 0, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n256_none", 1, 128, 256, "none", "none", 0, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n256_sw64", 1, 128, 256, "sw64", "sw64", 0, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("ss_m128n256_sw128", 1, 128, 256, "sw128", "sw128", 0, NM / 2, 0, 0, 0, 0));
    // --- kind::tf32 (K = 8): the peak of the other instruction kind a float32-exact score could use ---
    vs.push_back(mk("ss_m128n128_sw32_tf32", 1, 128, 128, "sw32", "sw32", 0, NM, 0, 0, 0, 0));
    vs.back().p.tf32 = 1; vs.back().p.idesc |= (2u << 7) | (2u << 10);
    vs.push_back(mk("ss_m128n256_sw32_tf32", 1, 128, 256, "sw32", "sw32", 0, NM / 2, 0, 0, 0, 0));
    vs.back().p.tf32 = 1; vs.back().p.idesc |= (2u << 7) | (2u << 10);
    // --- A from TMEM ---
    vs.push_back(mk("ts_m128n128_sw32", 1, 128, 128, "sw32", "sw32", 1, NM, 0, 0, 0, 0));
    vs.push_back(mk("ts_m128n128_sw64", 1, 128, 128, "sw32", "sw64", 1, NM, 0, 0, 0, 0));
    vs.push_back(mk("ts_m128n128_sw128", 1, 128, 128, "sw32", "sw128", 1, NM, 0, 0, 0, 0));
    vs.push_back(mk("ts_m128n256_sw32", 1, 128, 256, "sw32", "sw32", 1, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("ts_m128n256_sw64", 1, 128, 256, "sw32", "sw64", 1, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("ts_m128n256_sw128", 1, 128, 256, "sw32", "sw128", 1, NM / 2, 0, 0, 0, 0));
    // --- tcgen05.ld alone ---
    vs.push_back(mk("ld_8w", 1, 128, 128, "sw32", "sw32", 0, 0, 8, 4096, 0, 0));
    vs.push_back(mk("ld_4w", 1, 128, 128, "sw32", "sw32", 0, 0, 4, 4096, 0, 0));
    vs.push_back(mk("ld_8w_max", 1, 128, 128, "sw32", "sw32", 0, 0, 8, 4096, 1, 0));
    vs.push_back(mk("ld_4w_max", 1, 128, 128, "sw32", "sw32", 0, 0, 4, 4096, 1, 0));
    vs.push_back(mk("ld_8w_max_pipelined", 1, 128, 128, "sw32", "sw32", 0, 0, 8, 4096, 2, 0));
    vs.push_back(mk("ld_8w_pack16", 1, 128, 128, "sw32", "sw32", 0, 0, 8, 4096, 3, 0));
    vs.push_back(mk("ld_16w_pack16", 1, 128, 128, "sw32", "sw32", 0, 0, 16, 4096, 3, 0));
    vs.push_back(mk("ld_4w_pack16", 1, 128, 128, "sw32", "sw32", 0, 0, 4, 4096, 3, 0));
    vs.push_back(mk("ld_1w_pack16", 1, 128, 128, "sw32", "sw32", 0, 0, 1, 4096, 3, 0));
    vs.push_back(mk("ld_1w", 1, 128, 128, "sw32", "sw32", 0, 0, 1, 4096, 0, 0));
    vs.push_back(mk("ld_8w_pack16_max", 1, 128, 128, "sw32", "sw32", 0, 0, 8, 4096, 4, 0));
    vs.push_back(mk("ld_16w_pack16_max", 1, 128, 128, "sw32", "sw32", 0, 0, 16, 4096, 4, 0));
    vs.push_back(mk("ld_16w", 1, 128, 128, "sw32", "sw32", 0, 0, 16, 4096, 0, 0));
    vs.push_back(mk("ld_16w_max", 1, 128, 128, "sw32", "sw32", 0, 0, 16, 4096, 1, 0));
    vs.push_back(mk("ld_16w_max_pipelined", 1, 128, 128, "sw32", "sw32", 0, 0, 16, 4096, 2, 0));
    vs.push_back(mk("ss_n128_sw32+ld8max_pipelined_free", 1, 128, 128, "sw32", "sw32", 0, NM, 8, 1024, 2, 0));
    vs.push_back(mk("ss_n128_sw32+ld16max_free", 1, 128, 128, "sw32", "sw32", 0, NM, 16, 1024, 1, 0));
    vs.push_back(mk("ss_n128_sw32+ld16max_sync", 1, 128, 128, "sw32", "sw32", 0, NM, 16, 1024, 1, 1));
    // --- both, free-running (no hand-over) and with the accumulator hand-over ---
    vs.push_back(mk("ss_n128_sw32+ld8max_free", 1, 128, 128, "sw32", "sw32", 0, NM, 8, 1024, 1, 0));
    vs.push_back(mk("ss_n128_sw32+ld8max_sync", 1, 128, 128, "sw32", "sw32", 0, NM, 8, 1024, 1, 1));
    vs.push_back(mk("ss_n128_sw128+ld8max_sync", 1, 128, 128, "sw128", "sw128", 0, NM, 8, 1024, 1, 1));
    vs.push_back(mk("ss_n256_sw128+ld8max_free", 1, 128, 256, "sw128", "sw128", 0, NM / 2, 8, 1024, 1, 0));
    vs.push_back(mk("ts_n256_sw128+ld8max_free", 1, 128, 256, "sw32", "sw128", 1, NM / 2, 8, 1024, 1, 0));
    // --- CTA pair ---
    vs.push_back(mk("cg2_ss_m256n128_sw32", 2, 256, 128, "sw32", "sw32", 0, NM, 0, 0, 0, 0));
    vs.push_back(mk("cg2_ss_m256n256_sw32", 2, 256, 256, "sw32", "sw32", 0, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("cg2_ss_m256n256_sw128", 2, 256, 256, "sw128", "sw128", 0, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("cg2_ts_m256n256_sw128", 2, 256, 256, "sw32", "sw128", 1, NM / 2, 0, 0, 0, 0));
    vs.push_back(mk("cg2_ss_m256n256_sw128+ld8max_free", 2, 256, 256, "sw128", "sw128", 0, NM / 2, 8, 1024, 1, 0));

    if (argc < 2 || !strcmp(argv[1], "list")) {
        for (auto &v : vs) printf("%s\n", v.name.c_str());
        return 0;
    }
    const Variant *sel = nullptr;
    for (auto &v : vs)
        if (v.name == argv[1]) sel = &v;
    if (!sel) { fprintf(stderr, "unknown variant %s\n", argv[1]); return 2; }
    const int grid = 148;
    long long *d_out = nullptr, h_out[148 * 4];
    cudaMalloc(&d_out, sizeof h_out);
    auto launch = [&]() -> cudaError_t {
        cudaMemset(d_out, 0, sizeof h_out);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        const int nldw = sel->p.ld_warps > 8 ? 16 : 8;
        cfg.blockDim = dim3((nldw + 1) * 32);
        cfg.dynamicSmemBytes = kSmem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = sel->cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
#define LAUNCH(CGv, PBv, NWv)                                                                                         \
        if (sel->cg == CGv && sel->p.per_buf == PBv && nldw == NWv) {                                                     \
            cudaFuncSetAttribute(bench_kernel<CGv, PBv, NWv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);   \
            return cudaLaunchKernelEx(&cfg, bench_kernel<CGv, PBv, NWv>, sel->p, d_out, 3.0e38f);                         \
        }
        LAUNCH(1, 3, 8) LAUNCH(1, 6, 8) LAUNCH(2, 3, 8) LAUNCH(2, 6, 8)
        LAUNCH(1, 3, 16) LAUNCH(1, 6, 16) LAUNCH(2, 3, 16) LAUNCH(2, 6, 16)
        return cudaErrorInvalidValue;
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaError_t rc = launch();
        cudaEventRecord(e1);
        cudaError_t rs = cudaDeviceSynchronize();
        if (rc != cudaSuccess || rs != cudaSuccess) {
            printf("{\"variant\": \"%s\", \"error\": \"%s / %s\"}\n", sel->name.c_str(), cudaGetErrorString(rc), cudaGetErrorString(rs));
            return 1;
        }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpy(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost);
    double mma_sum = 0, mma_max = 0, ld_sum = 0, ld_max = 0;
    int n_mma_blocks = 0, n_ld = 0;
    for (int b = 0; b < grid; ++b) {
        if (h_out[b * 4]) { mma_sum += h_out[b * 4]; if (h_out[b * 4] > mma_max) mma_max = h_out[b * 4]; ++n_mma_blocks; }
        if (h_out[b * 4 + 1]) { ld_sum += h_out[b * 4 + 1]; if (h_out[b * 4 + 1] > ld_max) ld_max = h_out[b * 4 + 1]; ++n_ld; }
    }
    const Params &p = sel->p;
    printf("{\"variant\": \"%s\", \"ms\": %.4f", sel->name.c_str(), ms);
    if (n_mma_blocks)
        printf(", \"cyc_per_mma_avg\": %.1f, \"cyc_per_mma_max\": %.1f, \"n_mma\": %d", mma_sum / n_mma_blocks / p.n_mma, mma_max / p.n_mma, p.n_mma);
    if (n_ld)
        printf(", \"cyc_per_ld_iter_avg\": %.1f, \"cyc_per_ld_iter_max\": %.1f, \"ld_warps\": %d, \"bytes_per_iter_per_sm\": %d", ld_sum / n_ld / p.ld_iters,
               ld_max / p.ld_iters, p.ld_warps, 131072 * (p.ld_warps > 8 ? p.ld_warps / 16 : p.ld_warps / 8) + (p.ld_warps == 4 ? 65536 : 0));
    if (n_mma_blocks && !n_ld) {
        // whole-chip rate of this instruction stream: flop per MMA (M x N x K x 2) x MMAs x CTAs (or pairs) / kernel time
        const int M = (int)((p.idesc >> 24) & 0x1f) << 4, N = (int)((p.idesc >> 17) & 0x3f) << 3, K = p.tf32 ? 8 : 16;
        const double flop = 2.0 * M * N * K * (double)p.n_mma * (grid / sel->cg);
        printf(", \"M\": %d, \"N\": %d, \"K\": %d, \"kind\": \"%s\", \"tflops_whole_chip\": %.1f", M, N, K, p.tf32 ? "tf32" : "f16",
               flop / (ms * 1e-3) / 1e12);
    }
    printf("}\n");
    return 0;
}

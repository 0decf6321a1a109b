// topk_ffma.cu — A4: exact cosine-similarity candidate search, FP32 FFMA form
// (replaces range_candidates_from_embedding_emb + pad_candidates,
// /root/reference/fractal.py:535-552, driven per range by cpu_worker :598-623).
//
// scores = E @ q for every (query, domain) pair, top_k indices per query, best
// first, -1 padded.  The similarity matrix is never materialised:
//
//   * a CTA owns 8*TQ queries (one warp owns TQ of them, their embedding
//     vectors live in registers for the whole kernel);
//   * domain embeddings stream through shared memory in 256-row stages filled
//     with cp.async (3 stages), stored transposed ([float4 column][row]) so the
//     per-lane reads are conflict-free; a lane scores one domain per 32-row
//     chunk against the warp's TQ queries with one FMA chain per pair (the
//     canonical float32 score: ascending-k fmaf chain);
//   * each query keeps a running K-th-best threshold in a register; only
//     scores above it take the (rare, warp-cooperative) insertion path into the
//     query's candidate list in shared memory.
//
// Ordering is total and deterministic: score descending, then domain index
// ascending.  Bound: FP32 pipe, 2*emb_dim flop per pair.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kStageRows = 256;
constexpr int kStages = 3;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Is candidate (s1, i1) ranked before (s2, i2)?  score desc, index asc.
__device__ __forceinline__ bool ranks_before(float s1, int i1, float s2, int i2) {
    return s1 > s2 || (s1 == s2 && i1 < i2);
}

// Warp-cooperative insertion of every lane's passing score into one query's
// list (L entries in shared memory, entry p owned by lane p & 31).
__device__ __noinline__ float insert_passing(float acc, float tau, long long base, long long n_d,
                                            float *ls, int *li, int L, int lane) {
    unsigned m = __ballot_sync(kFull, acc > tau && base + lane < n_d);
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const float sc = __shfl_sync(kFull, acc, b);
        if (!(sc > tau)) continue;  // tau rose since the ballot (warp-uniform)
        const int id = (int)(base + b);
        // entry to evict: the one ranked last (lowest score, then highest index, then highest slot)
        float ws = FLT_MAX;
        int wi = -2, wp = -1;
        for (int p = lane; p < L; p += 32) {
            const float s = ls[p];
            const int i = li[p];
            if (wp < 0 || s < ws || (s == ws && i >= wi)) { ws = s; wi = i; wp = p; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const float os = __shfl_xor_sync(kFull, ws, off);
            const int oi = __shfl_xor_sync(kFull, wi, off);
            const int op = __shfl_xor_sync(kFull, wp, off);
            if (os < ws || (os == ws && (oi > wi || (oi == wi && op > wp)))) { ws = os; wi = oi; wp = op; }
        }
        if (lane == (wp & 31)) { ls[wp] = sc; li[wp] = id; }
        __syncwarp();
        float t2 = FLT_MAX;
        for (int p = lane; p < L; p += 32) t2 = fminf(t2, ls[p]);
#pragma unroll
        for (int off = 16; off; off >>= 1) t2 = fminf(t2, __shfl_xor_sync(kFull, t2, off));
        tau = t2;
    }
    return tau;
}

template <int ED, int TQ>
__global__ void __launch_bounds__(kThreads, 1)
topk_ffma_kernel(const float *__restrict__ Q, long long n_q, const float *__restrict__ E,
                 long long n_d, int top_k, int L, const uint8_t *__restrict__ active,
                 int32_t *__restrict__ cand, float *__restrict__ scores,
                 float *__restrict__ part_s, int32_t *__restrict__ part_i) {
    constexpr int C4 = ED / 4;                          // float4 columns per embedding row
    constexpr int PAD = C4 >= 8 ? 1 : 8 / C4;           // keeps the cp.async writes conflict-free
    constexpr int COLS = kStageRows + PAD;              // float4 per column per stage
    constexpr int QPB = kWarps * TQ;                    // queries per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *stage = reinterpret_cast<float4 *>(smem_raw);                     // [kStages][C4][COLS]
    float *list_s = reinterpret_cast<float *>(stage + kStages * C4 * COLS);   // [QPB][L]
    int *list_i = reinterpret_cast<int *>(list_s + QPB * L);                  // [QPB][L]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q0 = (long long)blockIdx.x * QPB + warp * TQ;

    // CTA-level early out: nothing to search (energy-pruned stretch or tail)
    {
        int any = 0;
        for (int i = threadIdx.x; i < QPB; i += kThreads) {
            const long long q = (long long)blockIdx.x * QPB + i;
            if (q < n_q && (!active || active[q])) any = 1;
        }
        if (!__syncthreads_or(any)) {
            if (gridDim.y > 1) return;          // split launch: merge_split_kernel writes the -1 rows
            for (int i = threadIdx.x; i < QPB * top_k; i += kThreads) {
                const long long q = (long long)blockIdx.x * QPB + i / top_k;
                if (q < n_q) {
                    cand[q * top_k + i % top_k] = -1;
                    if (scores) scores[q * top_k + i % top_k] = -INFINITY;
                }
            }
            return;
        }
    }

    // query vectors -> registers (warp-uniform addresses, broadcast loads)
    float qv[TQ][ED];
    float tau[TQ];
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
        const long long q = q0 + t;
        const bool live = q < n_q && (!active || active[q]);
        const float *src = Q + (q < n_q ? q : 0) * ED;
#pragma unroll
        for (int k = 0; k < ED; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + k));
            qv[t][k] = v.x; qv[t][k + 1] = v.y; qv[t][k + 2] = v.z; qv[t][k + 3] = v.w;
        }
        tau[t] = live ? -INFINITY : INFINITY;   // +inf: nothing ever passes, row stays -1
    }
    float *my_s = list_s + (warp * TQ) * L;
    int *my_i = list_i + (warp * TQ) * L;
    for (int p = lane; p < TQ * L; p += 32) { my_s[p] = -INFINITY; my_i[p] = -1; }
    __syncwarp();

    // gridDim.y > 1: the table is split between several CTAs per query block (few queries, e.g. the fallback of
    // the tensor-core search); each writes its partial top_k and merge_split_kernel merges them
    const long long n_tiles_all = (n_d + kStageRows - 1) / kStageRows;
    const long long t_lo = n_tiles_all * blockIdx.y / gridDim.y, n_tiles = n_tiles_all * (blockIdx.y + 1) / gridDim.y;
    auto issue = [&](long long tile) {
        if (tile < n_tiles) {
            float4 *dst = stage + ((tile - t_lo) % kStages) * (C4 * COLS);
            const long long row0 = tile * kStageRows;
#pragma unroll
            for (int g = threadIdx.x; g < kStageRows * C4; g += kThreads) {
                const int d = g / C4, c = g % C4;
                const long long row = row0 + d;
                const bool ok = row < n_d;
                cp_async16(dst + c * COLS + d, E + (ok ? row : 0) * ED + c * 4, ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) issue(t_lo + s);

    for (long long tile = t_lo; tile < n_tiles; ++tile) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        issue(tile + kStages - 1);
        const float4 *src = stage + ((tile - t_lo) % kStages) * (C4 * COLS);
#pragma unroll 1
        for (int j = 0; j < kStageRows / 32; ++j) {
            float e[ED];
#pragma unroll
            for (int c = 0; c < C4; ++c) {
                const float4 v = src[c * COLS + j * 32 + lane];
                e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
            }
            float acc[TQ];
            bool pass = false;
#pragma unroll
            for (int t = 0; t < TQ; ++t) {
                float a = 0.0f;
#pragma unroll
                for (int k = 0; k < ED; ++k) a = fmaf(qv[t][k], e[k], a);
                acc[t] = a;
                pass |= a > tau[t];
            }
            if (__any_sync(kFull, pass)) {
                const long long base = tile * kStageRows + j * 32;
#pragma unroll
                for (int t = 0; t < TQ; ++t)
                    if (__any_sync(kFull, acc[t] > tau[t]))
                        tau[t] = insert_passing(acc[t], tau[t], base, n_d, my_s + t * L, my_i + t * L, L, lane);
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();

    // rank sort of each list (score desc, index asc, slot asc) and write-out
    for (int t = 0; t < TQ; ++t) {
        const long long q = q0 + t;
        if (q >= n_q) break;
        const float *ls = my_s + t * L;
        const int *li = my_i + t * L;
        for (int p = lane; p < L; p += 32) {
            const float s = ls[p];
            const int i = li[p];
            int rank = 0;
            for (int o = 0; o < L; ++o) {
                const float so = ls[o];
                const int io = li[o];
                rank += (ranks_before(so, io, s, i) || (so == s && io == i && o < p)) ? 1 : 0;
            }
            if (rank < top_k) {
                if (gridDim.y > 1) {
                    const long long o = (q * gridDim.y + blockIdx.y) * top_k + rank;
                    part_s[o] = s;
                    part_i[o] = i;
                } else {
                    cand[q * top_k + rank] = i;
                    if (scores) scores[q * top_k + rank] = s;
                }
            }
        }
    }
}

// merges the partial results of a split launch: one warp per query, top_k rounds of "best entry ranked after the
// previous one" over n_split * top_k (score, index) entries; same total order as the kernel itself
__global__ void __launch_bounds__(128)
merge_split_kernel(const float *__restrict__ part_s, const int32_t *__restrict__ part_i, long long n_q, int n_split,
                   int top_k, const uint8_t *__restrict__ active, int32_t *__restrict__ cand, float *__restrict__ scores) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * 4 + warp;
    if (q >= n_q) return;
    const int n = n_split * top_k;
    const float *ps = part_s + q * n;
    const int32_t *pi = part_i + q * n;
    const bool live = !active || active[q];
    float prev_s = INFINITY;
    int prev_i = -1;
    for (int r = 0; r < top_k; ++r) {
        float bs = -INFINITY;
        int bi = -1;
        if (live) {
            for (int e = lane; e < n; e += 32) {
                const float s = ps[e];
                const int i = pi[e];
                if (i < 0) continue;
                if (ranks_before(prev_s, prev_i, s, i) && (bi < 0 || ranks_before(s, i, bs, bi))) { bs = s; bi = i; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float os = __shfl_xor_sync(kFull, bs, o);
                const int oi = __shfl_xor_sync(kFull, bi, o);
                if (oi >= 0 && (bi < 0 || ranks_before(os, oi, bs, bi))) { bs = os; bi = oi; }
            }
        }
        if (lane == 0) {
            cand[q * top_k + r] = bi;
            if (scores) scores[q * top_k + r] = bi >= 0 ? bs : -INFINITY;
        }
        if (bi >= 0) { prev_s = bs; prev_i = bi; }
    }
}

template <int ED, int TQ>
int launch(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb, int64_t n_d, int top_k,
           const uint8_t *d_active, int32_t *d_cand, float *d_scores, cudaStream_t st) {
    constexpr int C4 = ED / 4;
    constexpr int PAD = C4 >= 8 ? 1 : 8 / C4;
    constexpr int QPB = kWarps * TQ;
    const int L = ((top_k + 31) / 32) * 32;
    const size_t smem = sizeof(float4) * kStages * C4 * (kStageRows + PAD) + (size_t)QPB * L * 8;
    FWAV_CUDA(ctx, cudaFuncSetAttribute(topk_ffma_kernel<ED, TQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    const long long grid = (n_q + QPB - 1) / QPB;
    // few queries: split the table over several CTAs per query block so that the machine is full
    const long long n_tiles = (n_d + kStageRows - 1) / kStageRows;
    long long split = 1;
    if (grid < ctx->num_sms) {
        split = ctx->num_sms / grid;
        if (split > 64) split = 64;
        if (split > n_tiles / 8) split = n_tiles / 8;
        if (split < 1) split = 1;
    }
    float *d_ps = nullptr;
    int32_t *d_pi = nullptr;
    if (split > 1) {
        unsigned char *blk = nullptr;
        const size_t half = (size_t)n_q * split * top_k * 4;
        int rc = fwav_ws_reserve(ctx, WS_FFMA_PARTS, 2 * half, (void **)&blk);
        if (rc) return rc;
        d_ps = reinterpret_cast<float *>(blk);
        d_pi = reinterpret_cast<int32_t *>(blk + half);
    }
    topk_ffma_kernel<ED, TQ><<<dim3((unsigned)grid, (unsigned)split), kThreads, smem, st>>>(
        d_q, n_q, d_emb, n_d, top_k, L, d_active, d_cand, d_scores, d_ps, d_pi);
    FWAV_LAUNCH_CHECK(ctx);
    if (split > 1) {
        merge_split_kernel<<<(unsigned)((n_q + 3) / 4), 128, 0, st>>>(d_ps, d_pi, n_q, (int)split, top_k, d_active, d_cand,
                                                                    d_scores);
        FWAV_LAUNCH_CHECK(ctx);
    }
    return FWAV_OK;
}

}  // namespace

int fwav_launch_topk_ffma(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb,
                          int64_t n_d, int emb_dim, int top_k, const uint8_t *d_active,
                          int32_t *d_cand, float *d_scores, cudaStream_t st) {
    FWAV_REQUIRE(ctx, top_k >= 1 && top_k <= 256, "top_k %d outside [1, 256]", top_k);
    FWAV_REQUIRE(ctx, n_d < (1ll << 31), "n_domains %lld does not fit the int32 match index", (long long)n_d);
    FWAV_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_emb)) & 15) == 0,
                 "embedding tables must be 16-byte aligned");
    if (n_q == 0) return FWAV_OK;
    switch (emb_dim) {
        case 8: return launch<8, 8>(ctx, d_q, n_q, d_emb, n_d, top_k, d_active, d_cand, d_scores, st);
        case 16: return launch<16, 8>(ctx, d_q, n_q, d_emb, n_d, top_k, d_active, d_cand, d_scores, st);
        case 32: return launch<32, 4>(ctx, d_q, n_q, d_emb, n_d, top_k, d_active, d_cand, d_scores, st);
        case 64: return launch<64, 2>(ctx, d_q, n_q, d_emb, n_d, top_k, d_active, d_cand, d_scores, st);
        default:
            return fwav_set_error(ctx, FWAV_ERR_UNSUPPORTED,
                                  "emb_dim %d: search kernels are built for 8, 16, 32 and 64", emb_dim);
    }
}

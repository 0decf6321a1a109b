// tables.cu — A1 + A2/A3 in one pass: the domain table AND its embeddings from the signal (replaces
// build_domains_memmap, /root/reference/fractal.py:285-334, followed by build_domain_embeddings, :238-280, which
// re-reads the memmap row by row in a Python loop).
//
// For every tile_size that is a multiple of 256 (>= 1024) a domain value is the mean of run = 256 samples, which
// numpy reduces as leaf(128) + leaf(128); a leaf ("half sum") depends only on where it starts (domains.cu).  Two
// more sharing steps make the half sums themselves cheap, still bit for bit numpy's order:
//
//   * numpy's leaf keeps eight strided accumulators r_i = x[p+i] + x[p+i+8] + ... + x[p+i+120] (left to right) and
//     combines them as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).  r_i of the leaf at p is the CHAIN c(p+i), and c(q)
//     depends on q alone: the leaf at p is the tree over c(p..p+7).  15 additions per SAMPLE for the chains plus 7
//     per half sum, instead of 127 per half sum (one per 4 samples at config 2, one per sample at range_size 4).
//   * a thread computes 17 chains q, q+8, ..., q+128 from 32 staged samples (chain i and chain i+1 share 15 of
//     their 16 inputs, not their partial sums: each is still its own left-to-right chain).
//
// half_sums_chain_kernel: signal chunk -> shared memory (coalesced float4), chains -> shared memory, half sums ->
// HBM (n / domain_step floats, L2-resident for the next kernel).
// tables_from_halves_kernel<N, DS>: a block stages the window of half sums its 512 domains need, every thread builds
// a domain row in registers (two shared-memory loads, one add, one divide per value), stores it, runs the two-head
// embedding on the registers (embed_static.cuh) and stores that: the domain table is never read back.
//
// HBM traffic (algorithmic): 4 n read, 4 (range_size + 16) n_domains written — config 2: 286 MB.
#include "common.cuh"
#include "embed_static.cuh"
#include "fwav_math.cuh"
#include "tables_geom.h"

namespace {

static_assert(kChainOut % 8 == 0 && kChainX % 4 == 0, "block origins stay 16-byte aligned");

template <bool VEC>
__global__ void __launch_bounds__(kChainThreads)
half_sums_chain_kernel(const float *__restrict__ signal, long long n, long long n_half, int stride,
                       float *__restrict__ half) {
    __shared__ __align__(16) float xs[kChainX];
    __shared__ __align__(16) float cs[kChainP];
    const int tid = threadIdx.x;
    const long long Q0 = (long long)blockIdx.x * kChainOut;
    // 1. the block's samples (zero past the end: only chains no valid leaf uses read them)
    if (VEC) {
        for (int i = tid; i < kChainX / 4; i += kChainThreads) {
            const long long q = Q0 + 4ll * i;
            float4 v;
            if (q + 3 < n) {
                v = ld_stream_f4(reinterpret_cast<const float4 *>(signal + q));
            } else {
                v.x = q < n ? __ldg(signal + q) : 0.0f;
                v.y = q + 1 < n ? __ldg(signal + q + 1) : 0.0f;
                v.z = q + 2 < n ? __ldg(signal + q + 2) : 0.0f;
                v.w = 0.0f;
            }
            reinterpret_cast<float4 *>(xs)[i] = v;
        }
    } else {
        for (int i = tid; i < kChainX; i += kChainThreads) {
            const long long q = Q0 + i;
            xs[i] = q < n ? __ldg(signal + q) : 0.0f;
        }
    }
    __syncthreads();
    // 2. chains: thread = (residue r, run of 17 consecutive multiples of 8)
    {
        const int base = (tid & 7) + 8 * kChainT * (tid >> 3);
        float v[kChainT + 15];
#pragma unroll
        for (int i = 0; i < kChainT + 15; ++i) v[i] = xs[base + 8 * i];
#pragma unroll
        for (int i = 0; i < kChainT; ++i) {
            float c = v[i];
#pragma unroll
            for (int m = 1; m < 16; ++m) c = npm::add(c, v[i + m]);
            cs[base + 8 * i] = c;
        }
    }
    __syncthreads();
    // 3. leaves that start in [Q0, Q0 + kChainOut) at multiples of `stride`
    const long long u_lo = (Q0 + stride - 1) / stride;
    long long u_hi = (Q0 + kChainOut + stride - 1) / stride;
    if (u_hi > n_half) u_hi = n_half;
    const bool vec_c = (stride & 3) == 0;           // Q0 and p are multiples of 4: two 16-byte loads
    for (long long u = u_lo + tid; u < u_hi; u += kChainThreads) {
        const int p = (int)(u * stride - Q0);
        float c0, c1, c2, c3, c4, c5, c6, c7;
        if (vec_c) {
            const float4 a = *reinterpret_cast<const float4 *>(cs + p), b = *reinterpret_cast<const float4 *>(cs + p + 4);
            c0 = a.x; c1 = a.y; c2 = a.z; c3 = a.w; c4 = b.x; c5 = b.y; c6 = b.z; c7 = b.w;
        } else {
            c0 = cs[p]; c1 = cs[p + 1]; c2 = cs[p + 2]; c3 = cs[p + 3];
            c4 = cs[p + 4]; c5 = cs[p + 5]; c6 = cs[p + 6]; c7 = cs[p + 7];
        }
        half[u] = fwm::half_from_chains(c0, c1, c2, c3, c4, c5, c6, c7);
    }
}


// run == 256, leaves at every DS-th sample (DS = domain_step divides 128), emb_dim 16
template <int N, int DS>
__global__ void __launch_bounds__(kTabThreads, N <= 16 ? 7 : 6)
tables_from_halves_kernel(const float *__restrict__ half, long long n_half, long long n_dom,
                          float *__restrict__ domains, float *__restrict__ emb,
                          const __grid_constant__ TablesP<N, 8> T) {
    constexpr int KS = 256 / DS, HS = 128 / DS;      // half-sum index advance per column / to the second leaf
    constexpr int W = kTabJ + (N * 256 - 128) / DS;  // half sums a pass of kTabJ domains touches
    __shared__ float hs[W];
    const int tid = threadIdx.x;
    const bool wide = ((reinterpret_cast<uintptr_t>(domains) | reinterpret_cast<uintptr_t>(emb)) & 31) == 0;
    for (long long j0 = (long long)blockIdx.x * kTabJ; j0 < n_dom; j0 += (long long)gridDim.x * kTabJ) {
        for (int i = tid; i < W; i += kTabThreads) {
            const long long h = j0 + i;
            hs[i] = h < n_half ? __ldg(half + h) : 0.0f;
        }
        __syncthreads();
#pragma unroll 1
        for (int it = 0; it < kTabJ / kTabThreads; ++it) {
            const int jl = it * kTabThreads + tid;
            const long long j = j0 + jl;
            if (j < n_dom) {
                float x[N];
#pragma unroll
                for (int k = 0; k < N; ++k) x[k] = fwm::domain_from_halves(hs[jl + k * KS], hs[jl + k * KS + HS]);
                // every lane stores its own row: 32-byte stores where the tables allow (half the requests)
                if (N >= 8 && wide) {
#pragma unroll
                    for (int k = 0; k < N; k += 8) st_stream_f8(domains + j * N + k, x + k);
                } else {
                    float4 *drow = reinterpret_cast<float4 *>(domains + j * N);
#pragma unroll
                    for (int k = 0; k < N / 4; ++k)
                        st_stream_f4(drow + k, make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]));
                }
                float *erow = emb + j * 16;
                float out[8];
                embed_tonal_static<N, 8>(x, T, out);
                if (wide) {
                    st_stream_f8(erow, out);
                } else {
                    st_stream_f4(reinterpret_cast<float4 *>(erow), make_float4(out[0], out[1], out[2], out[3]));
                    st_stream_f4(reinterpret_cast<float4 *>(erow) + 1, make_float4(out[4], out[5], out[6], out[7]));
                }
                embed_transient_static<N, 8>(x, T, out);
                if (wide) {
                    st_stream_f8(erow + 8, out);
                } else {
                    st_stream_f4(reinterpret_cast<float4 *>(erow) + 2, make_float4(out[0], out[1], out[2], out[3]));
                    st_stream_f4(reinterpret_cast<float4 *>(erow) + 3, make_float4(out[4], out[5], out[6], out[7]));
                }
            }
        }
        __syncthreads();
    }
}

template <int N, int DS>
int launch_tables(fwav_ctx *ctx, const float *d_half, long long n_half, long long n_dom, float *d_domains,
                  float *d_emb, cudaStream_t st) {
    const FwavEmbedTables t = fwav_make_embed_tables(N, 8);
    TablesP<N, 8> P;
    for (int i = 0; i < 8 * N; ++i) { P.tonal[i] = t.tonal[i]; P.transient[i] = t.transient[i]; }
    for (int i = 0; i < N; ++i) P.w[i] = t.w[i];
    const long long need = (n_dom + kTabJ - 1) / kTabJ, cap = (long long)ctx->num_sms * 8;
    const int grid = (int)(need < cap ? need : cap);
    tables_from_halves_kernel<N, DS><<<grid, kTabThreads, 0, st>>>(d_half, n_half, n_dom, d_domains, d_emb, P);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

}  // namespace

// half[u] = numpy's 128-sample pairwise leaf starting at sample u * stride, u < n_half (needs (n_half-1)*stride + 128 <= n)
int fwav_launch_half_sums(fwav_ctx *ctx, const float *d_signal, int64_t n, int64_t n_half, int stride, float *d_half,
                          cudaStream_t st) {
    if (n_half <= 0) return FWAV_OK;
    const long long blocks = ((long long)(n_half - 1) * stride) / kChainOut + 1;
    FWAV_REQUIRE(ctx, blocks < (1ll << 31), "signal too long for one launch (%lld samples)", (long long)n);
    if ((reinterpret_cast<uintptr_t>(d_signal) & 15) == 0)
        half_sums_chain_kernel<true><<<(unsigned)blocks, kChainThreads, 0, st>>>(d_signal, n, n_half, stride, d_half);
    else
        half_sums_chain_kernel<false><<<(unsigned)blocks, kChainThreads, 0, st>>>(d_signal, n, n_half, stride, d_half);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

bool fwav_tables_fused_supported(const fwav_ctx *ctx, int tile, int N, int ds, int emb_dim, const float *d_domains,
                                 const float *d_emb) {
    if (ctx->embed_kind != FWAV_EMBED_TWO_HEAD || emb_dim != 16 || N <= 0 || tile / N != 256) return false;
    if (((reinterpret_cast<uintptr_t>(d_domains) | reinterpret_cast<uintptr_t>(d_emb)) & 15) != 0) return false;
    return (N == 4 && ds == 1) || (N == 8 && ds == 2) || (N == 16 && ds == 4) || (N == 32 && ds == 8);
}

// A1 + A3: domains and embeddings of a signal.  Fused where the geometry is one the reference derives from a
// tile_size that is a multiple of 256 (range_size 4 / 8 / 16 / 32, domain_step = range_size / 4) and the embedding is
// the two-head one at emb_dim 16; every other case runs the two stand-alone launchers (same bits).
int fwav_launch_tables(fwav_ctx *ctx, const float *d_signal, int64_t n, int tile, int N, int ds, int emb_dim,
                       float *d_domains, float *d_emb, cudaStream_t st) {
    FWAV_REQUIRE(ctx, tile > 0 && N > 0 && ds > 0 && tile / N >= 1, "bad geometry tile=%d N=%d ds=%d", tile, N, ds);
    const int64_t n_dom = fwav_count_domains(n, tile, ds);
    if (n_dom == 0) return FWAV_OK;
    int rc;
    if (!fwav_tables_fused_supported(ctx, tile, N, ds, emb_dim, d_domains, d_emb)) {
        if ((rc = fwav_launch_domains(ctx, d_signal, n, tile, N, ds, d_domains, st))) return rc;
        return fwav_launch_embed(ctx, d_domains, n_dom, N, emb_dim, d_emb, st);
    }
    const long long n_half = (n - 128) / ds + 1;
    float *d_half = nullptr;
    if ((rc = fwav_ws_reserve(ctx, WS_HALF, sizeof(float) * (size_t)n_half, (void **)&d_half))) return rc;
    if ((rc = fwav_launch_half_sums(ctx, d_signal, n, n_half, ds, d_half, st))) return rc;
    switch (N) {
    case 4: return launch_tables<4, 1>(ctx, d_half, n_half, n_dom, d_domains, d_emb, st);
    case 8: return launch_tables<8, 2>(ctx, d_half, n_half, n_dom, d_domains, d_emb, st);
    case 16: return launch_tables<16, 4>(ctx, d_half, n_half, n_dom, d_domains, d_emb, st);
    default: return launch_tables<32, 8>(ctx, d_half, n_half, n_dom, d_domains, d_emb, st);
    }
}

// affine.cu — A6: least-squares affine match R ~ s*D + o over the candidates
// and their mirrors (replaces _process_gpu_batch,
// /root/reference/fractal.py:757-850, which materialises six (B, 2K, N)
// temporaries in ~25 generic array kernels).
//
// One warp per range, one lane per candidate (32 candidates per pass).  A lane
// gathers its candidate's domain row (N floats, one or a few 16-byte loads),
// fits both orientations with fwm::affine_fit_pair — numpy's float32 operation
// order, so s, o and err are the reference's bits — and the warp takes the
// first minimum over [plain 0..K-1, mirrored 0..K-1] with a shuffle argmin
// keyed on (err, position).
//
// Bound: HBM (random row gathers): K*N*4 + N*4 + K*4 + 17 bytes per range.
// The north-star's "precomputed domain sums" are not used: caching sum(D) per
// domain would save flops this kernel has to spare, while the mirrored sums
// round differently for N < 8, so they would have to be stored twice.
#include "common.cuh"
#include "fwav_math.cuh"

#include <stdlib.h>

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Best {
    float err;
    int pos;      // orientation * K + candidate slot
    float s, o;
    int idx;
};

__device__ __forceinline__ bool better(float e1, int p1, float e2, int p2) {
    return e1 < e2 || (e1 == e2 && p1 < p2);
}

// PIPE = 0: every range loads its candidate indices, then the rows they name, then fits: two dependent memory
// latencies per range in front of ~600 instructions of arithmetic.
// PIPE = 1 (NT > 0, the first 32 candidates of a range): the warp's NEXT range is in flight while the current one
// is fitted -- its candidate indices were loaded one range earlier, its rows are requested before the fits start
// (NT more registers per lane) -- so the fits never wait for memory.
template <int NT, int PIPE>
__global__ void __launch_bounds__(256, NT == 16 ? (PIPE ? 2 : 3) : 1)
affine_kernel(const float *__restrict__ ranges, long long n_r, int N,
              const float *__restrict__ domains, long long n_d, const int32_t *__restrict__ cand, int K, float clipf,
              int32_t *__restrict__ o_idx, float *__restrict__ o_s, float *__restrict__ o_o,
              uint8_t *__restrict__ o_sym, float *__restrict__ o_err) {
    constexpr int NR = NT > 0 ? NT : 1;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    // candidate index of this lane in the first chunk of range i (-1: none / padding / past the table)
    auto load_raw = [&](long long i) -> int {
        int raw = -1;
        if (i < n_r && lane < K) raw = __ldg(cand + i * K + lane);
        return raw >= n_d ? -1 : raw;                                 // caller-supplied table: never read past the domains
    };
    auto load_row = [&](int raw, float (&t)[NR]) {
        if constexpr (NT > 0) {
            // (32-byte requests, LDG.E.256, were measured here and lost: 0.41 -> 0.46 ms)
            const float4 *tp = reinterpret_cast<const float4 *>(domains + (long long)(raw < 0 ? 0 : raw) * NT);   // :772-773
#pragma unroll
            for (int k = 0; k < NT; k += 4) {
                const float4 v = __ldg(tp + k / 4);
                t[k] = v.x; t[k + 1] = v.y; t[k + 2] = v.z; t[k + 3] = v.w;
            }
        }
    };
    int raw_cur = -1, raw_next = -1;
    float t_next[NR];
    if constexpr (PIPE) {
        raw_cur = load_raw(warp0);
        raw_next = load_raw(warp0 + n_warps);
        load_row(raw_cur, t_next);
    }
    for (long long i = warp0; i < n_r; i += n_warps) {
        const float *rp = ranges + i * N;
        float rreg[NR];
        if constexpr (NT > 0) {
#pragma unroll
            for (int k = 0; k < NT; k += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(rp + k));
                rreg[k] = v.x; rreg[k + 1] = v.y; rreg[k + 2] = v.z; rreg[k + 3] = v.w;
            }
        }
        float t_cur[NR];
        int raw0 = -1;
        if constexpr (PIPE) {
            // rotate: this range's rows arrived during the previous fits; request the next range's rows and the
            // indices of the one after it
            raw0 = raw_cur;
#pragma unroll
            for (int k = 0; k < NR; ++k) t_cur[k] = t_next[k];
            raw_cur = raw_next;
            load_row(raw_cur, t_next);
            raw_next = load_raw(i + 2 * n_warps);
        }
        auto r = [&](int k) {
            if constexpr (NT > 0) return rreg[k];
            else return __ldg(rp + k);
        };
        const float r_mean = fwm::range_mean<NT>(r, N);
        float rcreg[NR];                                               // r - mean(r), once per range (:791)
        if constexpr (NT > 0) {
#pragma unroll
            for (int k = 0; k < NT; ++k) rcreg[k] = npm::sub(rreg[k], r_mean);
        }
        auto rc = [&](int k) {
            if constexpr (NT > 0) return rcreg[k];
            else return npm::sub(__ldg(rp + k), r_mean);
        };

        Best best{INFINITY, 0x7fffffff, 0.0f, 0.0f, 0};
        for (int c0 = 0; c0 < K; c0 += 32) {
            const int c = c0 + lane;
            if (c < K) {
                int raw;
                if (PIPE && c0 == 0) {
                    raw = raw0;
                } else {
                    raw = __ldg(cand + i * K + c);
                    if (raw >= n_d) raw = -1;
                    if constexpr (NT > 0) load_row(raw, t_cur);
                }
                const int d = raw < 0 ? 0 : raw;                      // :772-773
                const float *tp = domains + (long long)d * N;
                auto plain = [&](int k) {
                    if constexpr (NT > 0) return t_cur[k];
                    else return __ldg(tp + k);
                };
                // both orientations; for N = 8 / 16 the tile's mean, centred values and sum of squares are shared
                // (mirroring cannot change numpy's pairwise sums there: fwm::affine_fit_pair)
                fwm::Fit f0, f1;
                fwm::affine_fit_pair<NT>(r, rc, r_mean, plain, N, f0, f1);
                if (raw < 0) { f0.err = INFINITY; f1.err = INFINITY; }  // :816-817
                if (better(f0.err, c, best.err, best.pos)) best = Best{f0.err, c, f0.s, f0.o, d};
                if (better(f1.err, K + c, best.err, best.pos)) best = Best{f1.err, K + c, f1.s, f1.o, d};
            }
        }
        // first argmin over [plain 0..K-1, mirrored 0..K-1]  (:820)
        float we = best.err;
        int wp = best.pos;
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const float oe = __shfl_xor_sync(kFull, we, off);
            const int op = __shfl_xor_sync(kFull, wp, off);
            if (better(oe, op, we, wp)) { we = oe; wp = op; }
        }
        if (best.pos == wp) {
            o_idx[i] = best.idx;
            o_s[i] = fwm::clip(best.s, -clipf, clipf);                // :823
            o_o[i] = best.o;                                          // :824
            o_sym[i] = wp >= K ? 1 : 0;
            o_err[i] = best.err;
        }
    }
}

}  // namespace

int fwav_launch_affine(fwav_ctx *ctx, const float *d_ranges, int64_t n_r, int N,
                       const float *d_domains, int64_t n_d, const int32_t *d_cand, int K,
                       double s_clip, int32_t *d_idx, float *d_s, float *d_o, uint8_t *d_sym,
                       float *d_err, cudaStream_t st) {
    FWAV_REQUIRE(ctx, N >= 1 && N <= fwm::kMaxRangeSize, "range_size %d out of range", N);
    FWAV_REQUIRE(ctx, K >= 1, "top_k %d must be positive", K);
    FWAV_REQUIRE(ctx, n_d >= 1, "affine match needs at least one domain");
    if (n_r == 0) return FWAV_OK;
    const float clipf = (float)fabs(s_clip);
    long long need = (n_r * 32 + 255) / 256;
    long long cap = (long long)ctx->num_sms * 8;
    const int grid = (int)(need < cap ? need : cap);
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_ranges) | reinterpret_cast<uintptr_t>(d_domains)) & 15) == 0;
    // FWAV_AFFINE_PIPE=0 / 1: measurement knob for the software-pipelined form (default below)
    const char *pipe_env = getenv("FWAV_AFFINE_PIPE");
    const bool pipe = pipe_env ? atoi(pipe_env) != 0 : true;
#define FWAV_AFFINE(NT)                                                                                        \
    do {                                                                                                       \
        if (pipe)                                                                                              \
            affine_kernel<NT, 1><<<grid, 256, 0, st>>>(d_ranges, n_r, N, d_domains, (long long)n_d, d_cand, K, clipf, d_idx, \
                                                       d_s, d_o, d_sym, d_err);                                \
        else                                                                                                   \
            affine_kernel<NT, 0><<<grid, 256, 0, st>>>(d_ranges, n_r, N, d_domains, (long long)n_d, d_cand, K, clipf, d_idx, \
                                                       d_s, d_o, d_sym, d_err);                                \
    } while (0)
    if (aligned && N == 4) FWAV_AFFINE(4);
    else if (aligned && N == 8) FWAV_AFFINE(8);
    else if (aligned && N == 16) FWAV_AFFINE(16);
    else if (aligned && N == 32) FWAV_AFFINE(32);
    else affine_kernel<0, 0><<<grid, 256, 0, st>>>(d_ranges, n_r, N, d_domains, (long long)n_d, d_cand, K, clipf, d_idx,
                                                   d_s, d_o, d_sym, d_err);
#undef FWAV_AFFINE
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

// common.cuh — context, error plumbing and small device helpers shared by the
// translation units of libfwav_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fwav_b200.h"
#include "embed_tables.h"

constexpr int kNumSMsB200 = 148;

struct fwav_ctx {
    int device = 0;
    int num_sms = kNumSMsB200;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host-buffer entry points: downloads that overlap the compute stream
    cudaEvent_t copy_event = nullptr;
    char err[512] = {0};
    int search_impl = FWAV_SEARCH_AUTO;
    int64_t launches = 0;
    int64_t umma_ffma_queries = 0;       // of those (top_k > 32), queries that also failed the second tensor-core chance
    int64_t umma_fallback_queries = 0;   // queries the fast search path handed to the exact list kernel
    // CUDA events around the kernels of the last tensor-core search call (fwav_ctx_search_timings)
    static constexpr int kSearchPhases = 5;          // pack, threshold pass, collect pass, finalize, list kernel
    static constexpr int kSearchSlots = 8;           // batches per call that are timed
    cudaEvent_t search_ev[kSearchSlots][kSearchPhases + 1] = {};
    int search_slots_used = 0;
    bool search_fast_path = false;
    bool search_hi_only = false;          // last collect pass filtered with the hi*hi term alone
    int search_route = 0;                 // last batch: 0 list kernel, 1 full split, 2 hi*hi float32 accumulators, 3 hi*hi fp16 accumulators

    // range_size the embedding tables of the coming searches were built for (0: unknown).  Set by the pipeline
    // entry points around their own search, or by fwav_ctx_set_search_range_size; tells the tensor-core search
    // which embedding dimensions can be non-zero at all (fractal.py:154-208) without a device round trip.
    int search_range_size = 0;
    int embed_kind = FWAV_EMBED_TWO_HEAD;     // what fwav_embed / the pipeline compute (fwav_ctx_set_embedding)

    // embedding matrices cached per (N, half)
    int emb_N = 0, emb_half = 0;
    double *d_tonal = nullptr, *d_transient = nullptr, *d_w = nullptr;

    // grow-only scratch arenas (device) and a pinned host staging block
    void *ws[32] = {nullptr};
    size_t ws_bytes[32] = {0};
    void *pinned = nullptr;               // staging ring of the host-buffer entry points (api.cu)
    size_t pinned_bytes = 0;
    cudaEvent_t ring_ev[16] = {};
};

// scratch slots
enum {
    WS_HALF = 0, WS_ACTIVE, WS_CAND, WS_QEMB, WS_DECODE_A, WS_DECODE_RED, WS_UMMA_E, WS_UMMA_Q, WS_UMMA_MISC,
    WS_UMMA_THETA, WS_UMMA_CBUF, WS_UMMA_CNT, WS_UMMA_FAIL, WS_UMMA_FB, WS_UMMA_PARTS, WS_FFMA_PARTS, WS_UMMA_TAIL,
    // device mirrors of the host-buffer entry points
    WS_H_SIGNAL, WS_H_RANGES, WS_H_DOMAINS, WS_H_EMB, WS_H_MATCH, WS_H_OUT,
    WS_PRESTEP, WS_DECODE_TILES, WS_UMMA_NORMS, WS_UMMA_LFAIL,
    WS_COUNT
};
static_assert(WS_COUNT <= 32, "grow fwav_ctx::ws");

int fwav_set_error(fwav_ctx *ctx, int code, const char *fmt, ...);
int fwav_ws_reserve(fwav_ctx *ctx, int slot, size_t bytes, void **out);
int fwav_embed_tables_device(fwav_ctx *ctx, int N, int half);

#define FWAV_CUDA(ctx, call)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fwav_set_error((ctx), FWAV_ERR_CUDA, "%s failed: %s (%s:%d)", #call,        \
                                  cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

#define FWAV_REQUIRE(ctx, cond, ...)                                                           \
    do {                                                                                       \
        if (!(cond)) return fwav_set_error((ctx), FWAV_ERR_INVALID, __VA_ARGS__);              \
    } while (0)

#define FWAV_LAUNCH_CHECK(ctx)                                                                 \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        FWAV_CUDA((ctx), cudaGetLastError());                                                  \
    } while (0)

static inline cudaStream_t fwav_stream(fwav_ctx *ctx, void *stream) {
    return stream ? (cudaStream_t)stream : ctx->stream;
}

// internal launchers (one per .cu file)
int fwav_launch_domains(fwav_ctx *ctx, const float *d_signal, int64_t n, int tile, int N, int ds,
                        float *d_domains, cudaStream_t st);
int fwav_launch_half_sums(fwav_ctx *ctx, const float *d_signal, int64_t n, int64_t n_half, int stride, float *d_half,
                          cudaStream_t st);
// domains + embeddings in one pass where the geometry allows (tables.cu), else the two launchers below
int fwav_launch_tables(fwav_ctx *ctx, const float *d_signal, int64_t n, int tile, int N, int ds, int emb_dim,
                       float *d_domains, float *d_emb, cudaStream_t st);
int fwav_launch_prestep(fwav_ctx *ctx, const float *d_signal, int64_t n, int N, double energy_thresh, float *d_ranges,
                        double *d_sumsq, cudaStream_t st);
int fwav_launch_embed(fwav_ctx *ctx, const float *d_rows, int64_t rows, int N, int emb_dim,
                      float *d_emb, cudaStream_t st);
int fwav_launch_activity(fwav_ctx *ctx, const float *d_ranges, int64_t n_r, int N, double thr,
                         int fast_mode, uint8_t *d_active, cudaStream_t st);
int fwav_launch_topk_ffma(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb,
                          int64_t n_d, int emb_dim, int top_k, const uint8_t *d_active,
                          int32_t *d_cand, float *d_scores, cudaStream_t st);
int fwav_launch_topk_umma(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb,
                          int64_t n_d, int emb_dim, int top_k, const uint8_t *d_active,
                          int32_t *d_cand, float *d_scores, cudaStream_t st);
bool fwav_topk_umma_supported(int emb_dim, int top_k, int64_t n_q, int64_t n_d);
int fwav_launch_affine(fwav_ctx *ctx, const float *d_ranges, int64_t n_r, int N,
                       const float *d_domains, int64_t n_d, const int32_t *d_cand, int K,
                       double s_clip, int32_t *d_idx, float *d_s, float *d_o, uint8_t *d_sym,
                       float *d_err, cudaStream_t st);
int fwav_launch_decode(fwav_ctx *ctx, const float *d_domains, int64_t n_d, const int32_t *d_idx,
                       const float *d_s, const float *d_o, const uint8_t *d_sym, int64_t n_r, int N,
                       int iterations, double eps, double s_clip, double s_damping, float *d_out,
                       int *iters_run, float *last_delta, cudaStream_t st);

int fwav_launch_decode_iter(fwav_ctx *ctx, const float *d_domains, int64_t n_d, const int32_t *d_idx,
                            const float *d_s, const float *d_o, const uint8_t *d_sym, int64_t n_r, int N,
                            double s_clip, double s_damping, int first, const float *d_cur, float *d_next,
                            double *d_sums, void *d_user_state, cudaStream_t st, void *const *targets = nullptr,
                            int n_targets = 0, int multimem = 0, int64_t target_offset = 0);
int fwav_launch_decode_converge(fwav_ctx *ctx, const double *d_sums_all, int n_parts, double eps, void *d_state,
                                cudaStream_t st);

#if defined(__CUDACC__)
// streaming loads/stores that do not pollute L1 (data touched once)
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// 32 bytes per lane in one instruction (sm_100: LDG.E.256): random row gathers of 64-byte rows cost two requests per
// row instead of four -- half the L1 wavefronts where every lane reads a different row (p must be 32-byte aligned)
__device__ __forceinline__ void ldg_f8(const float *p, float *v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void st_stream_f8(float *p, const float *v) {      // p 32-byte aligned
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
#endif

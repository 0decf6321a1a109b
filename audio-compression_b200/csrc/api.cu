// api.cu — the C ABI of libfwav_b200.so (include/fwav_b200.h): context and
// workspace management, argument checking, the device pipeline that replaces
// the reference's producer/consumer process pipeline
// (/root/reference/fractal.py:1174-1245), and the host-buffer entry points.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "fwav_math.cuh"

int fwav_set_error(fwav_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int fwav_ws_reserve(fwav_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->ws_bytes[slot] < bytes) {
        if (ctx->ws[slot]) {
            // the old block may still be in use by queued work
            FWAV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            FWAV_CUDA(ctx, cudaFree(ctx->ws[slot]));
            ctx->ws[slot] = nullptr;
            ctx->ws_bytes[slot] = 0;
        }
        const size_t want = bytes + bytes / 8;   // headroom so slowly growing inputs do not realloc
        cudaError_t e = cudaMalloc(&ctx->ws[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&ctx->ws[slot], bytes);
            if (e != cudaSuccess)
                return fwav_set_error(ctx, FWAV_ERR_NOMEM, "cudaMalloc(%zu) for workspace slot %d: %s",
                                      bytes, slot, cudaGetErrorString(e));
            ctx->ws_bytes[slot] = bytes;
        } else {
            ctx->ws_bytes[slot] = want;
        }
    }
    *out = ctx->ws[slot];
    return FWAV_OK;
}

extern "C" {

const char *fwav_version(void) { return "fwav_b200 0.1 (sm_100a)"; }

int fwav_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int fwav_ctx_create(int device, fwav_ctx **out) {
    if (!out) return FWAV_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        return FWAV_ERR_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess) return FWAV_ERR_CUDA;
    fwav_ctx *ctx = new fwav_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return FWAV_ERR_CUDA;
    }
    *out = ctx;
    return FWAV_OK;
}

int fwav_ctx_destroy(fwav_ctx *ctx) {
    if (!ctx) return FWAV_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < WS_COUNT; ++i)
        if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    for (auto &slot : ctx->search_ev)
        for (cudaEvent_t e : slot)
            if (e) cudaEventDestroy(e);
    if (ctx->d_tonal) cudaFree(ctx->d_tonal);
    if (ctx->d_transient) cudaFree(ctx->d_transient);
    if (ctx->d_w) cudaFree(ctx->d_w);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FWAV_OK;
}

const char *fwav_last_error(const fwav_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int fwav_ctx_sync(fwav_ctx *ctx) {
    if (!ctx) return FWAV_ERR_INVALID;
    FWAV_CUDA(ctx, cudaSetDevice(ctx->device));
    FWAV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FWAV_OK;
}

int fwav_ctx_set_search_impl(fwav_ctx *ctx, int impl) {
    if (!ctx) return FWAV_ERR_INVALID;
    FWAV_REQUIRE(ctx, impl >= FWAV_SEARCH_AUTO && impl <= FWAV_SEARCH_UMMA, "unknown search impl %d", impl);
    ctx->search_impl = impl;
    return FWAV_OK;
}

int64_t fwav_ctx_launch_count(const fwav_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t fwav_ctx_search_fallbacks(const fwav_ctx *ctx) { return ctx ? ctx->umma_fallback_queries : 0; }
int fwav_ctx_search_timings(fwav_ctx *ctx, float ms[5]) {
    if (!ctx || !ms) return FWAV_ERR_INVALID;
    FWAV_REQUIRE(ctx, ctx->search_slots_used > 0, "no tensor-core search has run on this context yet");
    for (int k = 0; k < fwav_ctx::kSearchPhases; ++k) ms[k] = 0.0f;
    for (int b = 0; b < ctx->search_slots_used; ++b)
        for (int k = 0; k < fwav_ctx::kSearchPhases; ++k) {
            float t = 0.0f;
            FWAV_CUDA(ctx, cudaEventElapsedTime(&t, ctx->search_ev[b][k], ctx->search_ev[b][k + 1]));
            ms[k] += t;
        }
    return FWAV_OK;
}

int fwav_geometry(int tile_size, int *range_size, int *domain_step) {
    const int rs = tile_size / 256 > 4 ? tile_size / 256 : 4;      // fractal.py:1070
    const int ds = rs / 4 > 1 ? rs / 4 : 1;                        // fractal.py:1071
    if (range_size) *range_size = rs;
    if (domain_step) *domain_step = ds;
    return FWAV_OK;
}

int64_t fwav_count_domains(int64_t n_samples, int tile_size, int domain_step) {
    if (tile_size <= 0 || domain_step <= 0 || n_samples < tile_size) return 0;   // fractal.py:297
    return (n_samples - tile_size) / domain_step + 1;                            // :301-302
}

#define FWAV_ENTER(ctx)                                          \
    do {                                                         \
        if (!(ctx)) return FWAV_ERR_INVALID;                     \
        (ctx)->err[0] = 0;                                       \
        FWAV_CUDA((ctx), cudaSetDevice((ctx)->device));          \
    } while (0)

int fwav_build_domains(fwav_ctx *ctx, const float *d_signal, int64_t n_samples, int tile_size,
                       int range_size, int domain_step, float *d_domains, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_signal && d_domains, "null buffer");
    return fwav_launch_domains(ctx, d_signal, n_samples, tile_size, range_size, domain_step, d_domains,
                               fwav_stream(ctx, stream));
}

int fwav_embed(fwav_ctx *ctx, const float *d_rows, int64_t rows, int range_size, int emb_dim,
               float *d_emb, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, rows == 0 || (d_rows && d_emb), "null buffer");
    return fwav_launch_embed(ctx, d_rows, rows, range_size, emb_dim, d_emb, fwav_stream(ctx, stream));
}

int fwav_range_activity(fwav_ctx *ctx, const float *d_ranges, int64_t n_ranges, int range_size,
                        double energy_thresh, int fast_mode, uint8_t *d_active, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, n_ranges == 0 || (d_ranges && d_active), "null buffer");
    return fwav_launch_activity(ctx, d_ranges, n_ranges, range_size, energy_thresh, fast_mode, d_active,
                                fwav_stream(ctx, stream));
}

static int topk_dispatch(fwav_ctx *ctx, const float *d_q, int64_t n_q, const float *d_emb, int64_t n_d,
                         int emb_dim, int top_k, const uint8_t *d_active, int32_t *d_cand, float *d_scores,
                         cudaStream_t st) {
    int impl = ctx->search_impl;
    if (impl == FWAV_SEARCH_AUTO)
        impl = fwav_topk_umma_supported(emb_dim, top_k, n_q, n_d) ? FWAV_SEARCH_UMMA : FWAV_SEARCH_FFMA;
    if (impl == FWAV_SEARCH_UMMA) {
        if (!fwav_topk_umma_supported(emb_dim, top_k, n_q, n_d))
            return fwav_set_error(ctx, FWAV_ERR_UNSUPPORTED,
                                  "tensor-core search does not cover emb_dim=%d top_k=%d", emb_dim, top_k);
        return fwav_launch_topk_umma(ctx, d_q, n_q, d_emb, n_d, emb_dim, top_k, d_active, d_cand, d_scores, st);
    }
    return fwav_launch_topk_ffma(ctx, d_q, n_q, d_emb, n_d, emb_dim, top_k, d_active, d_cand, d_scores, st);
}

int fwav_topk(fwav_ctx *ctx, const float *d_queries, int64_t n_queries, const float *d_emb,
              int64_t n_domains, int emb_dim, int top_k, const uint8_t *d_active, int32_t *d_cand,
              float *d_scores, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, n_queries == 0 || (d_queries && d_emb && d_cand), "null buffer");
    FWAV_REQUIRE(ctx, n_domains >= 0 && n_queries >= 0, "negative size");
    return topk_dispatch(ctx, d_queries, n_queries, d_emb, n_domains, emb_dim, top_k, d_active, d_cand,
                         d_scores, fwav_stream(ctx, stream));
}

int fwav_affine_match(fwav_ctx *ctx, const float *d_ranges, int64_t n_ranges, int range_size,
                      const float *d_domains, int64_t n_domains, const int32_t *d_cand, int top_k,
                      double s_clip, int32_t *d_idx, float *d_s, float *d_o, uint8_t *d_sym, float *d_err,
                      void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, n_ranges == 0 || (d_ranges && d_domains && d_cand && d_idx && d_s && d_o && d_sym && d_err),
                 "null buffer");
    return fwav_launch_affine(ctx, d_ranges, n_ranges, range_size, d_domains, n_domains, d_cand, top_k, s_clip,
                              d_idx, d_s, d_o, d_sym, d_err, fwav_stream(ctx, stream));
}

int fwav_decode(fwav_ctx *ctx, const float *d_domains, int64_t n_domains, const int32_t *d_idx,
                const float *d_s, const float *d_o, const uint8_t *d_sym, int64_t n_ranges, int range_size,
                int iterations, double convergence_eps, double s_clip, double s_damping, float *d_out,
                int *iters_run, float *last_delta, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, n_ranges == 0 || (d_domains && d_idx && d_s && d_o && d_sym && d_out), "null buffer");
    return fwav_launch_decode(ctx, d_domains, n_domains, d_idx, d_s, d_o, d_sym, n_ranges, range_size,
                              iterations, convergence_eps, s_clip, s_damping, d_out, iters_run, last_delta,
                              fwav_stream(ctx, stream));
}

int fwav_decode_iter(fwav_ctx *ctx, const float *d_domains, int64_t n_domains, const int32_t *d_idx,
                     const float *d_s, const float *d_o, const uint8_t *d_sym, int64_t n_ranges, int range_size,
                     double s_clip, double s_damping, int first, const float *d_cur, float *d_next,
                     double *d_sums, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_sums && (n_ranges == 0 || (d_domains && d_idx && d_s && d_o && d_sym && d_next)),
                 "null buffer");
    FWAV_REQUIRE(ctx, first || n_ranges == 0 || d_cur, "d_cur is required after the first iteration");
    return fwav_launch_decode_iter(ctx, d_domains, n_domains, d_idx, d_s, d_o, d_sym, n_ranges, range_size,
                                   s_clip, s_damping, first, d_cur, d_next, d_sums, fwav_stream(ctx, stream));
}

int fwav_compress_device(fwav_ctx *ctx, const float *d_signal, int64_t n_samples, const float *d_ranges,
                         int64_t n_ranges, int64_t query_offset, int tile_size, int emb_dim, int top_k,
                         double energy_thresh, int fast_mode, int query_mode, int build, float *d_domains,
                         float *d_emb, int32_t *d_idx, float *d_s, float *d_o, uint8_t *d_sym, float *d_err,
                         void *stream) {
    FWAV_ENTER(ctx);
    cudaStream_t st = fwav_stream(ctx, stream);
    int N, ds;
    fwav_geometry(tile_size, &N, &ds);
    const int64_t n_dom = fwav_count_domains(n_samples, tile_size, ds);
    FWAV_REQUIRE(ctx, n_dom > 0, "signal of %lld samples is shorter than one tile (%d): no domains",
                 (long long)n_samples, tile_size);
    FWAV_REQUIRE(ctx, n_ranges >= 0 && query_offset >= 0, "negative size");
    FWAV_REQUIRE(ctx, d_signal && d_domains && d_emb, "null table buffer");
    if (n_ranges == 0 && !build) return FWAV_OK;
    FWAV_REQUIRE(ctx, n_ranges == 0 || (d_ranges && d_idx && d_s && d_o && d_sym && d_err), "null match buffer");
    if (query_mode == 0)
        // the reference views the domain-embedding file as (n_ranges, emb_dim)
        // (fractal.py:1190-1195); np.memmap raises this when it is too short
        FWAV_REQUIRE(ctx, query_offset + n_ranges <= n_dom,
                     "mmap length is greater than file size: %lld ranges need rows [%lld, %lld) of a "
                     "%lld-row domain embedding table",
                     (long long)n_ranges, (long long)query_offset, (long long)(query_offset + n_ranges),
                     (long long)n_dom);
    int rc;
    if (build) {
        if ((rc = fwav_launch_domains(ctx, d_signal, n_samples, tile_size, N, ds, d_domains, st))) return rc;
        if ((rc = fwav_launch_embed(ctx, d_domains, n_dom, N, emb_dim, d_emb, st))) return rc;
    }
    if (n_ranges == 0) return FWAV_OK;
    uint8_t *d_active = nullptr;
    int32_t *d_cand = nullptr;
    if ((rc = fwav_ws_reserve(ctx, WS_ACTIVE, (size_t)n_ranges, (void **)&d_active))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_CAND, sizeof(int32_t) * (size_t)n_ranges * top_k, (void **)&d_cand))) return rc;
    if ((rc = fwav_launch_activity(ctx, d_ranges, n_ranges, N, energy_thresh, fast_mode, d_active, st))) return rc;
    const float *d_q = d_emb + query_offset * emb_dim;
    if (query_mode != 0) {
        float *d_qe = nullptr;
        if ((rc = fwav_ws_reserve(ctx, WS_QEMB, sizeof(float) * (size_t)n_ranges * emb_dim, (void **)&d_qe))) return rc;
        if ((rc = fwav_launch_embed(ctx, d_ranges, n_ranges, N, emb_dim, d_qe, st))) return rc;
        d_q = d_qe;
    }
    if ((rc = topk_dispatch(ctx, d_q, n_ranges, d_emb, n_dom, emb_dim, top_k, d_active, d_cand, nullptr, st)))
        return rc;
    return fwav_launch_affine(ctx, d_ranges, n_ranges, N, d_domains, n_dom, d_cand, top_k, 16.0, d_idx, d_s,
                              d_o, d_sym, d_err, st);
}

int fwav_compress_host(fwav_ctx *ctx, const float *h_signal, int64_t n_samples, const float *h_ranges,
                       int64_t n_ranges, int tile_size, int emb_dim, int top_k, double energy_thresh,
                       int fast_mode, int query_mode, float *h_domains, int32_t *h_idx, float *h_s,
                       float *h_o, uint8_t *h_sym, float *h_err) {
    FWAV_ENTER(ctx);
    cudaStream_t st = ctx->stream;
    int N, ds;
    fwav_geometry(tile_size, &N, &ds);
    const int64_t n_dom = fwav_count_domains(n_samples, tile_size, ds);
    FWAV_REQUIRE(ctx, h_signal && n_dom > 0, "signal of %lld samples is shorter than one tile (%d)",
                 (long long)n_samples, tile_size);
    FWAV_REQUIRE(ctx, n_ranges == 0 || (h_ranges && h_idx && h_s && h_o && h_sym && h_err), "null buffer");
    float *d_signal, *d_ranges, *d_domains, *d_emb;
    unsigned char *d_match;
    int rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_SIGNAL, sizeof(float) * (size_t)n_samples, (void **)&d_signal))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_RANGES, sizeof(float) * (size_t)n_ranges * N, (void **)&d_ranges))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_DOMAINS, sizeof(float) * (size_t)n_dom * N, (void **)&d_domains))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_EMB, sizeof(float) * (size_t)n_dom * emb_dim, (void **)&d_emb))) return rc;
    // idx | s | o | err | sym, each 16-byte aligned
    const size_t nr = (size_t)n_ranges, stride = (nr * 4 + 15) / 16 * 16;
    if ((rc = fwav_ws_reserve(ctx, WS_H_MATCH, stride * 4 + nr + 16, (void **)&d_match))) return rc;
    int32_t *d_idx = (int32_t *)d_match;
    float *d_s = (float *)(d_match + stride), *d_o = (float *)(d_match + 2 * stride);
    float *d_err = (float *)(d_match + 3 * stride);
    uint8_t *d_sym = d_match + 4 * stride;
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_signal, h_signal, sizeof(float) * (size_t)n_samples, cudaMemcpyHostToDevice, st));
    if (n_ranges)
        FWAV_CUDA(ctx, cudaMemcpyAsync(d_ranges, h_ranges, sizeof(float) * nr * N, cudaMemcpyHostToDevice, st));
    // Build the tables first; the domain table (the .fwav payload, the largest transfer of the call) then travels
    // to the host on the copy stream while the search runs on the compute stream.
    if ((rc = fwav_launch_domains(ctx, d_signal, n_samples, tile_size, N, ds, d_domains, st))) return rc;
    if (h_domains) {
        if (!ctx->copy_stream) FWAV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        if (!ctx->copy_event) FWAV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
        FWAV_CUDA(ctx, cudaEventRecord(ctx->copy_event, st));
        FWAV_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
        FWAV_CUDA(ctx, cudaMemcpyAsync(h_domains, d_domains, sizeof(float) * (size_t)n_dom * N, cudaMemcpyDeviceToHost,
                                       ctx->copy_stream));
    }
    // Once the download has been queued the caller's h_domains is in flight: every exit below goes through
    // `finish`, which waits for the copy stream before the buffer can be freed or reused.
    auto finish = [&](int code) -> int {
        if (h_domains) {
            const cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
            if (e != cudaSuccess && code == FWAV_OK)
                code = fwav_set_error(ctx, FWAV_ERR_CUDA, "download of the domain table failed: %s", cudaGetErrorString(e));
        }
        return code;
    };
    if ((rc = fwav_launch_embed(ctx, d_domains, n_dom, N, emb_dim, d_emb, st))) return finish(rc);
    rc = fwav_compress_device(ctx, d_signal, n_samples, d_ranges, n_ranges, 0, tile_size, emb_dim, top_k,
                              energy_thresh, fast_mode, query_mode, 0, d_domains, d_emb, d_idx, d_s, d_o,
                              d_sym, d_err, st);
    if (rc) return finish(rc);
    cudaError_t ce = cudaSuccess;
    if (n_ranges) {
        const struct { void *h; const void *d; size_t n; } out[5] = {
            {h_idx, d_idx, nr * 4}, {h_s, d_s, nr * 4}, {h_o, d_o, nr * 4}, {h_err, d_err, nr * 4}, {h_sym, d_sym, nr}};
        for (const auto &o : out)
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(o.h, o.d, o.n, cudaMemcpyDeviceToHost, st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess)
        return finish(fwav_set_error(ctx, FWAV_ERR_CUDA, "download of the matches failed: %s", cudaGetErrorString(ce)));
    return finish(FWAV_OK);
}

int fwav_decode_host(fwav_ctx *ctx, const float *h_domains, int64_t n_domains, const int32_t *h_idx,
                     const float *h_s, const float *h_o, const uint8_t *h_sym, int64_t n_ranges,
                     int range_size, int iterations, double convergence_eps, double s_clip,
                     double s_damping, float *h_out, int *iters_run, float *last_delta) {
    FWAV_ENTER(ctx);
    cudaStream_t st = ctx->stream;
    if (iters_run) *iters_run = 0;
    if (last_delta) *last_delta = 0.0f;
    if (n_ranges == 0) return FWAV_OK;
    FWAV_REQUIRE(ctx, h_domains && h_idx && h_s && h_o && h_sym && h_out, "null buffer");
    FWAV_REQUIRE(ctx, n_domains >= 1, "decoder needs at least one domain row");
    // a corrupt or forged container must not index past the table (the reference raises IndexError, fractal.py:1414)
    for (int64_t i = 0; i < n_ranges; ++i)
        FWAV_REQUIRE(ctx, h_idx[i] < n_domains, "index out of bounds: match %lld points at domain %d of %lld",
                     (long long)i, h_idx[i], (long long)n_domains);
    const int N = range_size;
    float *d_domains, *d_out;
    unsigned char *d_match;
    int rc;
    const size_t nr = (size_t)n_ranges, stride = (nr * 4 + 15) / 16 * 16;
    if ((rc = fwav_ws_reserve(ctx, WS_H_DOMAINS, sizeof(float) * (size_t)n_domains * N, (void **)&d_domains))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_MATCH, stride * 4 + nr + 16, (void **)&d_match))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_OUT, sizeof(float) * nr * N, (void **)&d_out))) return rc;
    int32_t *d_idx = (int32_t *)d_match;
    float *d_s = (float *)(d_match + stride), *d_o = (float *)(d_match + 2 * stride);
    uint8_t *d_sym = d_match + 4 * stride;
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_domains, h_domains, sizeof(float) * (size_t)n_domains * N, cudaMemcpyHostToDevice, st));
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_idx, h_idx, nr * 4, cudaMemcpyHostToDevice, st));
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_s, h_s, nr * 4, cudaMemcpyHostToDevice, st));
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_o, h_o, nr * 4, cudaMemcpyHostToDevice, st));
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_sym, h_sym, nr, cudaMemcpyHostToDevice, st));
    rc = fwav_launch_decode(ctx, d_domains, n_domains, d_idx, d_s, d_o, d_sym, n_ranges, N, iterations,
                            convergence_eps, s_clip, s_damping, d_out, iters_run, last_delta, st);
    if (rc) return rc;
    FWAV_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(float) * nr * N, cudaMemcpyDeviceToHost, st));
    FWAV_CUDA(ctx, cudaStreamSynchronize(st));
    return FWAV_OK;
}

int fwav_malloc(fwav_ctx *ctx, int64_t bytes, void **d_ptr) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_ptr && bytes >= 0, "bad argument");
    cudaError_t e = cudaMalloc(d_ptr, bytes > 0 ? (size_t)bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fwav_set_error(ctx, FWAV_ERR_NOMEM, "cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    }
    return FWAV_OK;
}

int fwav_free(fwav_ctx *ctx, void *d_ptr) {
    FWAV_ENTER(ctx);
    if (d_ptr) FWAV_CUDA(ctx, cudaFree(d_ptr));
    return FWAV_OK;
}

int fwav_host_alloc(fwav_ctx *ctx, int64_t bytes, void **h_ptr) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, h_ptr && bytes >= 0, "bad argument");
    cudaError_t e = cudaHostAlloc(h_ptr, bytes > 0 ? (size_t)bytes : 16, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *h_ptr = nullptr;
        return fwav_set_error(ctx, FWAV_ERR_NOMEM, "cudaHostAlloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    }
    return FWAV_OK;
}

int fwav_host_free(fwav_ctx *ctx, void *h_ptr) {
    // ctx may be NULL: a buffer can outlive the context it was allocated through (portable allocation)
    (void)ctx;
    if (h_ptr && cudaFreeHost(h_ptr) != cudaSuccess) {
        cudaGetLastError();
        return FWAV_ERR_CUDA;
    }
    return FWAV_OK;
}

int fwav_memcpy_h2d(fwav_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes, void *stream) {
    FWAV_ENTER(ctx);
    cudaStream_t st = fwav_stream(ctx, stream);
    FWAV_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, st));
    FWAV_CUDA(ctx, cudaStreamSynchronize(st));
    return FWAV_OK;
}

int fwav_memcpy_d2h(fwav_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes, void *stream) {
    FWAV_ENTER(ctx);
    cudaStream_t st = fwav_stream(ctx, stream);
    FWAV_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, st));
    FWAV_CUDA(ctx, cudaStreamSynchronize(st));
    return FWAV_OK;
}

}  // extern "C"

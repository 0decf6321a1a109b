// api.cu — the C ABI of libfwav_b200.so (include/fwav_b200.h): context and
// workspace management, argument checking, the device pipeline that replaces
// the reference's producer/consumer process pipeline
// (/root/reference/fractal.py:1174-1245), and the host-buffer entry points.
#include <stdarg.h>
#include <string.h>

#include <thread>

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

// ---------------------------------------------------------------------------
// Page-locked staging ring of the host-buffer entry points (row X1).  A caller's pageable buffer cannot be the
// end point of an asynchronous copy (the driver stages it and blocks), so such buffers go through a context-owned
// ring of kRingSlots x kRingSlotBytes of pinned memory, one cudaMemcpyAsync per chunk:
//   upload   : memcpy(user -> slot) by kUploadThreads threads while the DMA of the other chunks runs;
//   download : the DMA of chunk k+1.. runs while a helper thread memcpy's chunk k into the user's buffer, and all
//              of it runs beside the search on the compute stream.
// Buffers that already are page-locked (fwav_host_alloc, cudaHostRegister, torch pin_memory) take the direct
// asynchronous copy.
// ---------------------------------------------------------------------------
namespace {

constexpr size_t kRingSlotBytes = 2u << 20;      // 2 MB chunks: the first DMA of an upload starts 0.2 ms after the call (8 MB: 0.8 ms)
constexpr int kRingSlots = 16;
constexpr int kUploadThreads = 4;                 // a single thread copies ~10 GB/s into the ring, a third of what the DMA engine takes out
static_assert(kRingSlots % kUploadThreads == 0, "every slot belongs to one upload thread");

bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int ring_reserve(fwav_ctx *ctx) {
    if (!ctx->pinned) {
        cudaError_t e = cudaHostAlloc(&ctx->pinned, kRingSlotBytes * kRingSlots, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->pinned = nullptr;
            return fwav_set_error(ctx, FWAV_ERR_NOMEM, "cudaHostAlloc of the staging ring: %s", cudaGetErrorString(e));
        }
        ctx->pinned_bytes = kRingSlotBytes * kRingSlots;
        for (int i = 0; i < kRingSlots; ++i) FWAV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
    }
    return FWAV_OK;
}

// host -> device on `st`; returns once every byte has left the caller's buffer (the DMA may still be running)
int upload(fwav_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return FWAV_OK;
    if (is_pinned(h_src)) {
        FWAV_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return FWAV_OK;
    }
    int rc = ring_reserve(ctx);
    if (rc) return rc;
    const size_t n_chunks = (bytes + kRingSlotBytes - 1) / kRingSlotBytes;
    // one thread's share: chunks c = t, t + kUploadThreads, ... (memcpy into slot c % kRingSlots, DMA out of it)
    auto run_share = [&](int t, bool set_device) -> cudaError_t {
        cudaError_t e = set_device ? cudaSetDevice(ctx->device) : cudaSuccess;
        for (size_t c = (size_t)t; c < n_chunks && e == cudaSuccess; c += kUploadThreads) {
            const int slot = (int)(c % kRingSlots);
            unsigned char *stage = static_cast<unsigned char *>(ctx->pinned) + slot * kRingSlotBytes;
            const size_t off = c * kRingSlotBytes, len = bytes - off < kRingSlotBytes ? bytes - off : kRingSlotBytes;
            if (c >= (size_t)kRingSlots) e = cudaEventSynchronize(ctx->ring_ev[slot]);      // the slot's previous DMA has left it
            if (e != cudaSuccess) break;
            memcpy(stage, static_cast<const unsigned char *>(h_src) + off, len);
            e = cudaMemcpyAsync(static_cast<unsigned char *>(d_dst) + off, stage, len, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ring_ev[slot], st);
        }
        return e;
    };
    cudaError_t err = cudaSuccess;
    if (n_chunks < 4) {
        err = run_share(0, false);
        for (int t = 1; t < kUploadThreads && err == cudaSuccess; ++t) err = run_share(t, false);
    } else {
        // nothing on the device can start before the whole signal is there: several threads fill the ring (the
        // copies into device memory are independent, their order on the stream does not matter; a slot is only
        // ever touched by the thread that owns its chunks)
        cudaError_t errs[kUploadThreads];
        std::thread thr[kUploadThreads];
        for (int t = 1; t < kUploadThreads; ++t) thr[t] = std::thread([&, t]() { errs[t] = run_share(t, true); });
        errs[0] = run_share(0, false);
        for (int t = 1; t < kUploadThreads; ++t) thr[t].join();
        for (int t = 0; t < kUploadThreads; ++t)
            if (errs[t] != cudaSuccess && err == cudaSuccess) err = errs[t];
    }
    if (err != cudaSuccess)
        return fwav_set_error(ctx, FWAV_ERR_CUDA, "upload through the staging ring failed: %s", cudaGetErrorString(err));
    return FWAV_OK;
}

// device -> pageable host through the ring, on the context's copy stream, from a helper thread.  The copy stream
// already waits for the kernel that produced d_src.
struct Download {
    std::thread worker;
    cudaError_t err = cudaSuccess;
    bool active = false;

    void start(fwav_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
        active = true;
        worker = std::thread([=]() {
            cudaError_t e = cudaSetDevice(ctx->device);
            cudaStream_t cs = ctx->copy_stream;
            const size_t n_chunks = (bytes + kRingSlotBytes - 1) / kRingSlotBytes;
            auto issue = [&](size_t c) {
                const int slot = (int)(c % kRingSlots);
                const size_t off = c * kRingSlotBytes, len = bytes - off < kRingSlotBytes ? bytes - off : kRingSlotBytes;
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(static_cast<unsigned char *>(ctx->pinned) + slot * kRingSlotBytes,
                                        static_cast<const unsigned char *>(d_src) + off, len, cudaMemcpyDeviceToHost, cs);
                if (e == cudaSuccess) e = cudaEventRecord(ctx->ring_ev[slot], cs);
            };
            for (size_t c = 0; c < n_chunks && c < (size_t)kRingSlots; ++c) issue(c);
            for (size_t c = 0; c < n_chunks && e == cudaSuccess; ++c) {
                const int slot = (int)(c % kRingSlots);
                const size_t off = c * kRingSlotBytes, len = bytes - off < kRingSlotBytes ? bytes - off : kRingSlotBytes;
                e = cudaEventSynchronize(ctx->ring_ev[slot]);
                if (e != cudaSuccess) break;
                memcpy(static_cast<unsigned char *>(h_dst) + off, static_cast<unsigned char *>(ctx->pinned) + slot * kRingSlotBytes, len);
                if (c + kRingSlots < n_chunks) issue(c + kRingSlots);
            }
            err = e;
        });
    }
    cudaError_t join() {
        if (active) {
            worker.join();
            active = false;
        }
        return err;
    }
};

}  // namespace

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
    for (cudaEvent_t e : ctx->ring_ev)
        if (e) cudaEventDestroy(e);
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

int fwav_ctx_set_embedding(fwav_ctx *ctx, int kind) {
    if (!ctx) return FWAV_ERR_INVALID;
    FWAV_REQUIRE(ctx, kind == FWAV_EMBED_TWO_HEAD || kind == FWAV_EMBED_TONAL, "unknown embedding kind %d", kind);
    ctx->embed_kind = kind;
    return FWAV_OK;
}

int fwav_ctx_set_search_range_size(fwav_ctx *ctx, int range_size) {
    if (!ctx) return FWAV_ERR_INVALID;
    FWAV_REQUIRE(ctx, range_size >= 0 && range_size <= fwm::kMaxRangeSize, "range_size %d out of range", range_size);
    ctx->search_range_size = range_size;
    return FWAV_OK;
}

int64_t fwav_ctx_launch_count(const fwav_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t fwav_ctx_search_fallbacks(const fwav_ctx *ctx) { return ctx ? ctx->umma_fallback_queries : 0; }
int fwav_ctx_search_route(const fwav_ctx *ctx) { return ctx ? ctx->search_route : -1; }
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

int fwav_build_tables(fwav_ctx *ctx, const float *d_signal, int64_t n_samples, int tile_size, int range_size,
                      int domain_step, int emb_dim, float *d_domains, float *d_emb, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_signal && d_domains && d_emb, "null buffer");
    return fwav_launch_tables(ctx, d_signal, n_samples, tile_size, range_size, domain_step, emb_dim, d_domains, d_emb,
                              fwav_stream(ctx, stream));
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
                                   s_clip, s_damping, first, d_cur, d_next, d_sums, nullptr, fwav_stream(ctx, stream));
}

int fwav_decode_iter_gated(fwav_ctx *ctx, const float *d_domains, int64_t n_domains, const int32_t *d_idx,
                           const float *d_s, const float *d_o, const uint8_t *d_sym, int64_t n_ranges, int range_size,
                           double s_clip, double s_damping, int first, const float *d_cur, float *d_next,
                           double *d_sums, fwav_decode_state *d_state, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_sums && d_state && (n_ranges == 0 || (d_domains && d_idx && d_s && d_o && d_sym && d_next)),
                 "null buffer");
    FWAV_REQUIRE(ctx, first || n_ranges == 0 || d_cur, "d_cur is required after the first iteration");
    return fwav_launch_decode_iter(ctx, d_domains, n_domains, d_idx, d_s, d_o, d_sym, n_ranges, range_size,
                                   s_clip, s_damping, first, d_cur, d_next, d_sums, d_state, fwav_stream(ctx, stream));
}

int fwav_decode_iter_bcast(fwav_ctx *ctx, const float *d_domains, int64_t n_domains, const int32_t *d_idx,
                           const float *d_s, const float *d_o, const uint8_t *d_sym, int64_t n_ranges, int range_size,
                           double s_clip, double s_damping, int first, const float *d_cur, float *d_next,
                           double *d_sums, fwav_decode_state *d_state, void *const *targets, int n_targets,
                           int multimem, int64_t target_offset, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_sums && d_state && (n_ranges == 0 || (d_domains && d_idx && d_s && d_o && d_sym && d_next)),
                 "null buffer");
    FWAV_REQUIRE(ctx, first || n_ranges == 0 || d_cur, "d_cur is required after the first iteration");
    FWAV_REQUIRE(ctx, n_targets >= 1 && n_targets <= 8 && targets && (!multimem || n_targets == 1),
                 "1..8 broadcast targets (exactly one multicast address)");
    return fwav_launch_decode_iter(ctx, d_domains, n_domains, d_idx, d_s, d_o, d_sym, n_ranges, range_size,
                                   s_clip, s_damping, first, d_cur, d_next, d_sums, d_state, fwav_stream(ctx, stream),
                                   targets, n_targets, multimem, target_offset);
}

int fwav_decode_converge(fwav_ctx *ctx, const double *d_sums_all, int n_parts, double convergence_eps,
                         fwav_decode_state *d_state, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_sums_all && d_state && n_parts >= 1, "bad argument");
    return fwav_launch_decode_converge(ctx, d_sums_all, n_parts, convergence_eps, d_state, fwav_stream(ctx, stream));
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
        if ((rc = fwav_launch_tables(ctx, d_signal, n_samples, tile_size, N, ds, emb_dim, d_domains, d_emb, st))) return rc;
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
    const int hint_before = ctx->search_range_size;
    ctx->search_range_size = N;          // both tables were embedded for this geometry (queries: rows of d_emb, or the ranges)
    rc = topk_dispatch(ctx, d_q, n_ranges, d_emb, n_dom, emb_dim, top_k, d_active, d_cand, nullptr, st);
    ctx->search_range_size = hint_before;
    if (rc) return rc;
    return fwav_launch_affine(ctx, d_ranges, n_ranges, N, d_domains, n_dom, d_cand, top_k, 16.0, d_idx, d_s,
                              d_o, d_sym, d_err, st);
}

// Shared body of the two host-buffer compress entry points.  h_ranges_in != NULL: the caller framed the ranges
// (fwav_compress_host); NULL: the pre-step runs on the device from the raw signal (fwav_compress_signal_host) and
// h_ranges_out, if given, receives the framed ranges.  *silent (may be NULL) is set when the reference takes its
// "silent input" early-out (fractal.py:1083): no outputs are written then.
static int compress_host_impl(fwav_ctx *ctx, const float *h_signal, int64_t n_samples, const float *h_ranges_in,
                              int64_t n_ranges, int tile_size, int emb_dim, int top_k, double energy_thresh,
                              int fast_mode, int query_mode, float *h_ranges_out, float *h_domains, int32_t *h_idx,
                              float *h_s, float *h_o, uint8_t *h_sym, float *h_err, int *silent) {
    cudaStream_t st = ctx->stream;
    int N, ds;
    fwav_geometry(tile_size, &N, &ds);
    const int64_t n_dom = fwav_count_domains(n_samples, tile_size, ds);
    FWAV_REQUIRE(ctx, h_signal && n_dom > 0, "signal of %lld samples is shorter than one tile (%d)",
                 (long long)n_samples, tile_size);
    FWAV_REQUIRE(ctx, n_ranges == 0 || (h_idx && h_s && h_o && h_sym && h_err), "null buffer");
    if (silent) *silent = 0;
    float *d_signal, *d_ranges, *d_domains, *d_emb;
    unsigned char *d_match;
    int rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_SIGNAL, sizeof(float) * (size_t)n_samples, (void **)&d_signal))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_RANGES, sizeof(float) * (size_t)n_ranges * N + 8, (void **)&d_ranges))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_DOMAINS, sizeof(float) * (size_t)n_dom * N, (void **)&d_domains))) return rc;
    if ((rc = fwav_ws_reserve(ctx, WS_H_EMB, sizeof(float) * (size_t)n_dom * emb_dim, (void **)&d_emb))) return rc;
    // idx | s | o | err | sym, each 16-byte aligned
    const size_t nr = (size_t)n_ranges, stride = (nr * 4 + 15) / 16 * 16;
    if ((rc = fwav_ws_reserve(ctx, WS_H_MATCH, stride * 4 + nr + 16, (void **)&d_match))) return rc;
    int32_t *d_idx = (int32_t *)d_match;
    float *d_s = (float *)(d_match + stride), *d_o = (float *)(d_match + 2 * stride);
    float *d_err = (float *)(d_match + 3 * stride);
    uint8_t *d_sym = d_match + 4 * stride;
    if (!ctx->copy_stream) FWAV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_event) FWAV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
    if ((rc = upload(ctx, d_signal, h_signal, sizeof(float) * (size_t)n_samples, st))) return rc;
    // The tables (domains and embeddings, one fused pass) are built first; the domain table (the .fwav payload, the
    // largest transfer of the call) then travels to the host on the copy stream while the pre-step and the search run on the compute stream.
    if ((rc = fwav_launch_tables(ctx, d_signal, n_samples, tile_size, N, ds, emb_dim, d_domains, d_emb, st))) return rc;
    FWAV_CUDA(ctx, cudaEventRecord(ctx->copy_event, st));
    if (h_ranges_in) {
        if ((rc = upload(ctx, d_ranges, h_ranges_in, sizeof(float) * nr * N, st))) return rc;
    } else {
        // A0 on the device; the reference's early-out needs the sum of squares on the host: 8 bytes and one sync,
        // while the domain kernels are already queued behind it
        double *d_sumsq = reinterpret_cast<double *>(d_ranges + nr * N), h_sumsq = 0.0;
        if ((rc = fwav_launch_prestep(ctx, d_signal, n_samples, N, energy_thresh, d_ranges, d_sumsq, st))) return rc;
        FWAV_CUDA(ctx, cudaMemcpyAsync(&h_sumsq, d_sumsq, sizeof(double), cudaMemcpyDeviceToHost, st));
        FWAV_CUDA(ctx, cudaStreamSynchronize(st));
        if ((float)h_sumsq < 1e-8f) {            // np.sum(weighted_signal ** 2) < 1e-8 (:1083); float64 sum here
            if (silent) *silent = 1;
            return FWAV_OK;
        }
    }
    // Once a download has been queued the caller's buffers are in flight: every exit below goes through `finish`,
    // which waits for the copy stream (and the helper thread) before they can be freed or reused.
    Download dl;
    bool direct_dl = false;
    if (h_domains) {
        FWAV_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
        const size_t bytes = sizeof(float) * (size_t)n_dom * N;
        if (is_pinned(h_domains)) {
            direct_dl = true;
            FWAV_CUDA(ctx, cudaMemcpyAsync(h_domains, d_domains, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        } else {
            if ((rc = ring_reserve(ctx))) return rc;
            FWAV_CUDA(ctx, cudaStreamSynchronize(st));      // the upload may still own ring slots
            dl.start(ctx, h_domains, d_domains, bytes);
        }
    }
    auto finish = [&](int code) -> int {
        cudaError_t e = dl.join();
        if (e == cudaSuccess && direct_dl) e = cudaStreamSynchronize(ctx->copy_stream);
        if (e != cudaSuccess && code == FWAV_OK)
            code = fwav_set_error(ctx, FWAV_ERR_CUDA, "download of the domain table failed: %s", cudaGetErrorString(e));
        return code;
    };
    rc = fwav_compress_device(ctx, d_signal, n_samples, d_ranges, n_ranges, 0, tile_size, emb_dim, top_k,
                              energy_thresh, fast_mode, query_mode, 0, d_domains, d_emb, d_idx, d_s, d_o,
                              d_sym, d_err, st);
    if (rc) return finish(rc);
    cudaError_t ce = cudaSuccess;
    if (n_ranges) {
        const struct { void *h; const void *d; size_t n; } out[6] = {
            {h_idx, d_idx, nr * 4}, {h_s, d_s, nr * 4}, {h_o, d_o, nr * 4}, {h_err, d_err, nr * 4}, {h_sym, d_sym, nr},
            {h_ranges_out, d_ranges, h_ranges_out ? nr * N * 4 : 0}};
        for (const auto &o : out)
            if (ce == cudaSuccess && o.n) ce = cudaMemcpyAsync(o.h, o.d, o.n, cudaMemcpyDeviceToHost, st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess)
        return finish(fwav_set_error(ctx, FWAV_ERR_CUDA, "download of the matches failed: %s", cudaGetErrorString(ce)));
    return finish(FWAV_OK);
}

int fwav_compress_host(fwav_ctx *ctx, const float *h_signal, int64_t n_samples, const float *h_ranges,
                       int64_t n_ranges, int tile_size, int emb_dim, int top_k, double energy_thresh,
                       int fast_mode, int query_mode, float *h_domains, int32_t *h_idx, float *h_s,
                       float *h_o, uint8_t *h_sym, float *h_err) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, n_ranges == 0 || h_ranges, "null buffer");
    return compress_host_impl(ctx, h_signal, n_samples, h_ranges, n_ranges, tile_size, emb_dim, top_k, energy_thresh,
                              fast_mode, query_mode, nullptr, h_domains, h_idx, h_s, h_o, h_sym, h_err, nullptr);
}

int fwav_compress_signal_host(fwav_ctx *ctx, const float *h_signal, int64_t n_samples, int tile_size, int emb_dim,
                              int top_k, double energy_thresh, int fast_mode, int query_mode, float *h_ranges,
                              float *h_domains, int32_t *h_idx, float *h_s, float *h_o, uint8_t *h_sym,
                              float *h_err, int *silent) {
    FWAV_ENTER(ctx);
    int N, ds;
    fwav_geometry(tile_size, &N, &ds);
    const int64_t n_ranges = (n_samples + N - 1) / N;
    return compress_host_impl(ctx, h_signal, n_samples, nullptr, n_ranges, tile_size, emb_dim, top_k, energy_thresh,
                              fast_mode, query_mode, h_ranges, h_domains, h_idx, h_s, h_o, h_sym, h_err, silent);
}

int fwav_prepare_ranges(fwav_ctx *ctx, const float *d_signal, int64_t n_samples, int range_size,
                        double energy_thresh, float *d_ranges, double *d_sumsq, void *stream) {
    FWAV_ENTER(ctx);
    FWAV_REQUIRE(ctx, d_signal && d_ranges && d_sumsq, "null buffer");
    return fwav_launch_prestep(ctx, d_signal, n_samples, range_size, energy_thresh, d_ranges, d_sumsq,
                               fwav_stream(ctx, stream));
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

/*
 * fwav_b200.h — C ABI of the B200-native FWAV hot path (libfwav_b200.so).
 *
 * The reference (xavenordu/Audio-Compression, fractal.py) has no FFI layer: its
 * boundary is the Python function surface.  These entry points are what a
 * `ctypes` binding inside fractal.py would call in place of the numpy/CuPy
 * bodies cited beside each declaration (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every call returns 0 on success or a negative fwav_status; the message is
 *     retrievable with fwav_last_error(ctx).  Nothing throws across the ABI.
 *   - `d_` parameters are DEVICE pointers on the context's device, `h_`
 *     parameters are HOST pointers.  The caller owns every data buffer; the
 *     context owns its stream and scratch workspace.
 *   - `stream` is a cudaStream_t passed as void*; NULL means the context's own
 *     stream.  Device-pointer calls are asynchronous on that stream unless
 *     stated otherwise; host-pointer calls return after the results are in the
 *     host buffers.
 *   - one context per host thread; contexts are created lazily by callers so a
 *     forked worker can initialise CUDA itself (fractal.py:1605 runs files in a
 *     multiprocessing.Pool).
 *   - all floating point data is IEEE binary32, row-major.
 */
#ifndef FWAV_B200_H
#define FWAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fwav_ctx fwav_ctx;

typedef enum fwav_status {
    FWAV_OK = 0,
    FWAV_ERR_INVALID = -1,     /* bad argument (message says which) */
    FWAV_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    FWAV_ERR_UNSUPPORTED = -3, /* geometry outside the compiled kernels */
    FWAV_ERR_NOMEM = -4
} fwav_status;

/* Which kernel computes the similarity + top-K contraction. */
typedef enum fwav_search_impl {
    FWAV_SEARCH_AUTO = 0,  /* tensor-core path when the shape allows it */
    FWAV_SEARCH_FFMA = 1,  /* FP32 FFMA warp-select kernel */
    FWAV_SEARCH_UMMA = 2   /* tcgen05 split-fp16 kernels (sampled threshold, collect, exact FP32 finalize; list kernel) */
} fwav_search_impl;

/* Which embedding fwav_embed and the pipeline entry points compute. */
typedef enum fwav_embedding {
    FWAV_EMBED_TWO_HEAD = 0, /* multi_head_embedding(tile, emb_dim/2, emb_dim/2) (fractal.py:166-175): the reference's LIVE path */
    FWAV_EMBED_TONAL = 1     /* tile_embedding(tile, k = emb_dim) (fractal.py:178-208): the DCT-II / DC-removed / high-frequency-
                                weighted / L2-normalised embedding the README and the north star describe (EMBED_K = 32) */
} fwav_embedding;

const char *fwav_version(void);
int fwav_device_count(void);

int fwav_ctx_create(int device, fwav_ctx **out);
int fwav_ctx_destroy(fwav_ctx *ctx);
const char *fwav_last_error(const fwav_ctx *ctx);
/* Block until everything queued on the context's stream has finished. */
int fwav_ctx_sync(fwav_ctx *ctx);
/* Select the search kernel for subsequent calls (default FWAV_SEARCH_AUTO). */
int fwav_ctx_set_search_impl(fwav_ctx *ctx, int impl);
/* Select the embedding for subsequent calls (default FWAV_EMBED_TWO_HEAD).  Together with query_mode = 1 of the
 * pipeline entry points this is the north star's "fixed" mode: true range embeddings, tile_embedding with EMBED_K. */
int fwav_ctx_set_embedding(fwav_ctx *ctx, int kind);
/* Tell subsequent fwav_topk calls that BOTH of their tables were produced by fwav_embed for this range_size (0 =
 * unknown, the default).  For range_size < 8 at most seven embedding dimensions can be non-zero (fractal.py:154-208)
 * and the tensor-core search then uses its two-MMA compact split; results are identical either way.  The pipeline
 * entry points (fwav_compress_*) know their geometry and do not need this. */
int fwav_ctx_set_search_range_size(fwav_ctx *ctx, int range_size);
/* Number of kernels this library has launched through `ctx` since creation. */
int64_t fwav_ctx_launch_count(const fwav_ctx *ctx);
/* Queries the tensor-core search handed from its sampled-threshold fast path to the exact list kernel since
   creation (verification failed or candidate buffer overflowed).  Diagnostics only; results are exact either way. */
int64_t fwav_ctx_search_fallbacks(const fwav_ctx *ctx);
/* Device time, in milliseconds, of the phases of the LAST tensor-core fwav_topk / fwav_compress_* search on `ctx`,
   from CUDA events recorded on its stream: ms[0] operand packing, ms[1] threshold pass (sample table), ms[2] collect
   pass (the dominant kernel: every query against every domain), ms[3] finalize (exact re-score + verification),
   ms[4] exact list kernel (small tables, or the queries the fast path handed over).  Call after the stream has been
   synchronised; returns FWAV_ERR_INVALID if no tensor-core search has run yet. */
int fwav_ctx_search_timings(fwav_ctx *ctx, float ms[5]);
/* Which collect pass the LAST batch of the last tensor-core search took: 0 exact list kernel (small tables), 1 full
   fp16 hi/lo split (three instructions per tile), 2 hi*hi term alone with float32 accumulators, 3 hi*hi term alone with
   half-precision accumulators (collect_hi_kernel).  Decided per batch from the room the queries leave for each
   filter's error bound; results are identical either way. */
int fwav_ctx_search_route(const fwav_ctx *ctx);

/* Derived geometry of compress_audio (fractal.py:1070-1071) and the domain
 * count of build_domains_memmap (fractal.py:297-304). */
int fwav_geometry(int tile_size, int *range_size, int *domain_step);
int64_t fwav_count_domains(int64_t n_samples, int tile_size, int domain_step);

/* A1 — replaces build_domains_memmap (fractal.py:285-334).
 * d_domains[j*range_size + k] = numpy-order float32 mean of the k-th run of
 * tile_size/range_size samples of the window starting at j*domain_step.
 * Bit-identical to the reference's memmap contents. */
int fwav_build_domains(fwav_ctx *ctx, const float *d_signal, int64_t n_samples,
                       int tile_size, int range_size, int domain_step,
                       float *d_domains, void *stream);

/* A2/A3 — replaces multi_head_embedding over every row
 * (fractal.py:154-208 called from build_domain_embeddings :271-277).
 * d_emb is (rows, emb_dim): [tonal emb_dim/2 | transient | zero pad]. */
int fwav_embed(fwav_ctx *ctx, const float *d_rows, int64_t rows, int range_size,
               int emb_dim, float *d_emb, void *stream);

/* A1 + A3 in one pass — replaces build_domains_memmap (fractal.py:285-334) followed by
 * build_domain_embeddings (:238-280, the pass that re-reads the memmap row by row).  Same outputs, bit for bit, as
 * fwav_build_domains + fwav_embed; for the geometries compress_audio derives from a tile_size that is a multiple
 * of 256 (range_size 4 / 8 / 16 / 32, domain_step = range_size / 4) with the two-head embedding at emb_dim 16 the
 * domain rows never leave the registers between the two steps (the table is written once and not read back). */
int fwav_build_tables(fwav_ctx *ctx, const float *d_signal, int64_t n_samples, int tile_size, int range_size,
                      int domain_step, int emb_dim, float *d_domains, float *d_emb, void *stream);

/* A4/A5 — replaces cpu_worker's linear search + pad_candidates
 * (fractal.py:535-552, 598-623) for all queries at once.
 * d_cand is (n_queries, top_k) int32: indices of the top_k largest
 * d_emb·q, best first, -1 padded.  d_active (may be NULL) marks queries to
 * search; rows of inactive queries are filled with -1 (energy prune, :602).
 * d_scores (may be NULL) receives the matching scores. */
int fwav_topk(fwav_ctx *ctx, const float *d_queries, int64_t n_queries,
              const float *d_emb, int64_t n_domains, int emb_dim, int top_k,
              const uint8_t *d_active, int32_t *d_cand, float *d_scores, void *stream);

/* Energy prune flags of cpu_worker (fractal.py:602):
 * d_active[i] = !(fast_mode && mean(range_i^2) < 0.75*energy_thresh). */
int fwav_range_activity(fwav_ctx *ctx, const float *d_ranges, int64_t n_ranges,
                        int range_size, double energy_thresh, int fast_mode,
                        uint8_t *d_active, void *stream);

/* A6 — replaces _process_gpu_batch (fractal.py:757-850): least-squares
 * R ~ s*D + o over the candidates and their mirrors, first argmin of the L2
 * residual, s clipped to +-s_clip afterwards.  Arithmetic follows numpy's
 * float32 operation order, so results are bit-identical given the same
 * candidate table. */
int fwav_affine_match(fwav_ctx *ctx, const float *d_ranges, int64_t n_ranges, int range_size,
                      const float *d_domains, int64_t n_domains,
                      const int32_t *d_cand, int top_k, double s_clip,
                      int32_t *d_idx, float *d_s, float *d_o, uint8_t *d_sym, float *d_err,
                      void *stream);

/* A9 — replaces the iteration loop of decompress_audio (fractal.py:1411-1467).
 * d_out must hold n_ranges*range_size floats.  The convergence test runs on
 * the device; *iters_run and *last_delta are written after an internal sync. */
int fwav_decode(fwav_ctx *ctx, const float *d_domains, int64_t n_domains,
                const int32_t *d_idx, const float *d_s, const float *d_o, const uint8_t *d_sym,
                int64_t n_ranges, int range_size, int iterations, double convergence_eps,
                double s_clip, double s_damping, float *d_out,
                int *iters_run, float *last_delta, void *stream);

/* One iteration of the same loop over a slice of ranges, for range-sharded
 * multi-GPU decoding: reads d_cur (ignored when first != 0: the reconstruction
 * starts at zero, fractal.py:1389), writes d_next, and leaves
 * d_sums[0] = sum (next-cur)^2, d_sums[1] = sum cur^2 (float64, fixed order) on
 * the device.  Asynchronous; the caller combines the sums across ranks and
 * applies the convergence test of fractal.py:1460-1467. */
int fwav_decode_iter(fwav_ctx *ctx, const float *d_domains, int64_t n_domains,
                     const int32_t *d_idx, const float *d_s, const float *d_o, const uint8_t *d_sym,
                     int64_t n_ranges, int range_size, double s_clip, double s_damping, int first,
                     const float *d_cur, float *d_next, double *d_sums, void *stream);

/* The same with the convergence decision on the device (no host read-back per iteration).  d_state points at a
 * ZEROED fwav_decode_state in device memory that the caller keeps for the whole decode:
 *   fwav_decode_iter_gated   returns at once (d_next untouched) when d_state->done is set;
 *   fwav_decode_converge     adds the n_parts pairs of sums in d_sums_all ([part][2] float64, e.g. the all-gathered
 *                            d_sums of every rank) in index order, computes delta as fractal.py:1460-1461 does,
 *                            counts the iteration and sets `done` when delta < convergence_eps (:1465).
 * After the last launch the caller reads the state once: iters_run iterations ran, and iteration `it` wrote the
 * buffer that was its d_next. */
typedef struct fwav_decode_state {
    int32_t iters_run;
    int32_t done;
    float delta;
    int32_t bad_index;   /* a match pointed past the domain table (it was decoded as a sentinel) */
} fwav_decode_state;
int fwav_decode_iter_gated(fwav_ctx *ctx, const float *d_domains, int64_t n_domains,
                           const int32_t *d_idx, const float *d_s, const float *d_o, const uint8_t *d_sym,
                           int64_t n_ranges, int range_size, double s_clip, double s_damping, int first,
                           const float *d_cur, float *d_next, double *d_sums,
                           fwav_decode_state *d_state, void *stream);
int fwav_decode_converge(fwav_ctx *ctx, const double *d_sums_all, int n_parts, double convergence_eps,
                         fwav_decode_state *d_state, void *stream);
/* fwav_decode_iter_gated fused with the exchange the north star asks for ("the reconstruction buffer is all-gathered
 * over NVLink each iteration"): besides d_next, every output sample is stored at element target_offset + i of the FULL
 * reconstruction buffer of every GPU, from inside the kernel.  `targets` is a HOST array of n_targets device addresses
 * of that buffer: with multimem != 0 exactly one, the NVSwitch multicast address of a symmetric allocation (one
 * multimem.st per 16 bytes, replicated by the switch); otherwise one peer-mapped pointer per GPU (plain stores over
 * NVLink).  No collective follows the kernel; the caller synchronises the ranks once before reading the full
 * buffers.  range_size 4, 8, 16 or 32. */
int fwav_decode_iter_bcast(fwav_ctx *ctx, const float *d_domains, int64_t n_domains,
                           const int32_t *d_idx, const float *d_s, const float *d_o, const uint8_t *d_sym,
                           int64_t n_ranges, int range_size, double s_clip, double s_damping, int first,
                           const float *d_cur, float *d_next, double *d_sums, fwav_decode_state *d_state,
                           void *const *targets, int n_targets, int multimem, int64_t target_offset,
                           void *stream);

/* Device pipeline A1..A7 on resident inputs: replaces the process/queue
 * pipeline of compress_audio (fractal.py:1114-1245) between "ranges framed"
 * and "matches collected".  d_signal is the raw signal (domains are built
 * from it, :1120), d_ranges the masked, reflect-padded, framed ranges (:1112).
 * Queries are rows [query_offset, query_offset+n_ranges) of the domain
 * embedding table when query_mode==0 (the reference's live aliasing,
 * :1190-1195) or embeddings of the ranges themselves when query_mode==1.
 * d_domains (n_domains*range_size) and d_emb (n_domains*emb_dim) are outputs
 * the caller keeps; pass build=0 to reuse tables already built there. */
int fwav_compress_device(fwav_ctx *ctx, const float *d_signal, int64_t n_samples,
                         const float *d_ranges, int64_t n_ranges, int64_t query_offset,
                         int tile_size, int emb_dim, int top_k, double energy_thresh,
                         int fast_mode, int query_mode, int build,
                         float *d_domains, float *d_emb,
                         int32_t *d_idx, float *d_s, float *d_o, uint8_t *d_sym, float *d_err,
                         void *stream);

/* Host-buffer form of the same pipeline (what compress_audio calls):
 * copies the signal and ranges in, runs the pipeline, copies domains and
 * matches out.  h_domains may be NULL when the caller does not want them. */
int fwav_compress_host(fwav_ctx *ctx, const float *h_signal, int64_t n_samples,
                       const float *h_ranges, int64_t n_ranges,
                       int tile_size, int emb_dim, int top_k, double energy_thresh,
                       int fast_mode, int query_mode,
                       float *h_domains,
                       int32_t *h_idx, float *h_s, float *h_o, uint8_t *h_sym, float *h_err);

/* A0 / row N2 -- replaces voiced_detection + masking + reflect padding + framing (fractal.py:880-909, :1074-1112)
 * on the device: d_ranges receives ceil(n_samples / range_size) * range_size floats (the gated signal with its
 * reflected tail, i.e. the (n_ranges, range_size) array of :1112, bit-identical to the reference's), d_sumsq[0]
 * (device, float64) the sum of squares of the gated signal that decides the "silent input" early-out of :1083.
 * Needs at least five frames of 2 * range_size samples. */
int fwav_prepare_ranges(fwav_ctx *ctx, const float *d_signal, int64_t n_samples, int range_size,
                        double energy_thresh, float *d_ranges, double *d_sumsq, void *stream);

/* compress_audio from the RAW signal (what fractal.compress_audio calls): fwav_prepare_ranges on the device, then
 * the pipeline of fwav_compress_host; n_ranges = ceil(n_samples / range_size).  h_ranges (may be NULL) receives
 * the framed ranges.  *silent is set to 1, and no output is written, when the reference returns its empty result
 * for a silent input (:1083-1093).  Pageable buffers are staged through a context-owned page-locked ring in 2 MB
 * chunks (upload on the calling thread, download of the domain table on a helper thread beside the search);
 * page-locked buffers (fwav_host_alloc) are the end points of the asynchronous copies themselves. */
int fwav_compress_signal_host(fwav_ctx *ctx, const float *h_signal, int64_t n_samples,
                              int tile_size, int emb_dim, int top_k, double energy_thresh,
                              int fast_mode, int query_mode,
                              float *h_ranges, float *h_domains,
                              int32_t *h_idx, float *h_s, float *h_o, uint8_t *h_sym, float *h_err,
                              int *silent);

/* Host-buffer decoder (what decompress_audio calls). h_out: n_ranges*range_size. */
int fwav_decode_host(fwav_ctx *ctx, const float *h_domains, int64_t n_domains,
                     const int32_t *h_idx, const float *h_s, const float *h_o, const uint8_t *h_sym,
                     int64_t n_ranges, int range_size, int iterations, double convergence_eps,
                     double s_clip, double s_damping, float *h_out,
                     int *iters_run, float *last_delta);

/* Plain device memory helpers so a host language without a CUDA binding can
 * drive the device-pointer entry points. */
int fwav_malloc(fwav_ctx *ctx, int64_t bytes, void **d_ptr);
int fwav_free(fwav_ctx *ctx, void *d_ptr);
/* Page-locked host memory for the buffers handed to fwav_compress_host / fwav_decode_host: with pageable
 * buffers the result copies are staged by the driver and cannot overlap the search (config 2: the 127 MB domain
 * table).  Portable across contexts; free with fwav_host_free (its ctx argument may be NULL: a buffer may
 * outlive the context it was allocated through). */
int fwav_host_alloc(fwav_ctx *ctx, int64_t bytes, void **h_ptr);
int fwav_host_free(fwav_ctx *ctx, void *h_ptr);
int fwav_memcpy_h2d(fwav_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes, void *stream);
int fwav_memcpy_d2h(fwav_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FWAV_B200_H */

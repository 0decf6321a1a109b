// host_harness.cpp — TEST INFRASTRUCTURE.  Compiles the host/device math of
// audio-compression_b200/csrc (np_math.cuh, fwav_math.cuh, embed_tables.h) for
// the CPU so tests/test_host_math.py can pin it to the reference's golden
// vectors without a GPU.  It is never linked into libfwav_b200.so and the
// product never loads it.
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "fwav_math.cuh"
#include "embed_tables.h"
#include "tables_geom.h"
#include "embed_static.cuh"

template <int NS, class R, class T>
static fwm::Fit fit_dispatch(R r, float r_mean, T t, int N) {
    return fwm::affine_fit<NS>(r, r_mean, t, N);
}
#define HH_BY_N(N, CALL)                        \
    ((N) == 4 ? CALL(4) : (N) == 8 ? CALL(8) : (N) == 16 ? CALL(16) : (N) == 32 ? CALL(32) : CALL(0))

// the kernel's own form: both orientations of a candidate through fwm::affine_fit_pair (shared tile statistics for
// N = 8 / 16), first minimum over [plain 0..K-1, mirrored 0..K-1]
template <int NS>
static void affine_pair_rows(const float *ranges, long long n_r, int N, const float *domains, const int32_t *cand, int K,
                             float clipf, int32_t *idx, float *s, float *o, uint8_t *sym, float *err) {
    for (long long i = 0; i < n_r; ++i) {
        const float *r = ranges + i * N;
        auto rr = [&](int k) { return r[k]; };
        const float r_mean = fwm::range_mean<NS>(rr, N);
        auto rc = [&](int k) { return npm::sub(r[k], r_mean); };
        float best = std::numeric_limits<float>::infinity();
        int best_pos = -1;
        fwm::Fit best_fit{0, 0, 0};
        std::vector<fwm::Fit> f0(K), f1(K);
        for (int c = 0; c < K; ++c) {
            const int raw = cand[i * K + c];
            const float *t = domains + (long long)(raw < 0 ? 0 : raw) * N;
            auto tt = [&](int k) { return t[k]; };
            fwm::affine_fit_pair<NS>(rr, rc, r_mean, tt, N, f0[c], f1[c]);
            if (raw < 0) f0[c].err = f1[c].err = std::numeric_limits<float>::infinity();
        }
        for (int pos = 0; pos < 2 * K; ++pos) {
            const fwm::Fit &f = pos < K ? f0[pos] : f1[pos - K];
            if (best_pos < 0 || f.err < best) { best = f.err; best_pos = pos; best_fit = f; }
        }
        const int raw = cand[i * K + best_pos % K];
        idx[i] = raw < 0 ? 0 : raw;
        s[i] = fwm::clip(best_fit.s, -clipf, clipf);
        o[i] = best_fit.o;
        sym[i] = best_pos >= K;
        err[i] = best_fit.err;
    }
}

// the kernels' compile-time form of the two-head embedding (embed_static.cuh: half-length chains on the sums and
// differences of mirrored samples)
template <int N>
static void embed_static_rows(const float *rows, long long n_rows, float *out) {
    const FwavEmbedTables t = fwav_make_embed_tables(N, 8);
    static TablesP<N, 8> P;
    for (int i = 0; i < 8 * N; ++i) { P.tonal[i] = t.tonal[i]; P.transient[i] = t.transient[i]; }
    for (int i = 0; i < N; ++i) P.w[i] = t.w[i];
    for (long long r = 0; r < n_rows; ++r) {
        float x[N], o[16];
        for (int i = 0; i < N; ++i) x[i] = rows[r * N + i];
        embed_row_static<N, 8>(x, P, o);
        for (int i = 0; i < 16; ++i) out[r * 16 + i] = o[i];
    }
}

extern "C" {

float hh_np_mean(const float *a, int n) {
    auto at = [&](int i) { return a[i]; };
    return npm::np_mean<3>(at, n);
}

// same decomposition as domains.cu: half sums when run == 256, generic otherwise
void hh_build_domains(const float *sig, long long n, int tile, int N, int ds, float *out, int force_generic) {
    long long nd = n < tile ? 0 : (n - tile) / ds + 1;
    int run = tile / N;
    auto s = [&](long long i) { return sig[i]; };
    for (long long j = 0; j < nd; ++j)
        for (int k = 0; k < N; ++k) {
            long long start = j * ds + (long long)k * run;
            if (run == 256 && !force_generic)
                out[j * N + k] = fwm::domain_from_halves(fwm::half_sum128(s, start), fwm::half_sum128(s, start + 128));
            else
                out[j * N + k] = fwm::domain_value(s, start, run);
        }
}

// CPU emulation of tables.cu, block by block and thread by thread with the kernels' own index arithmetic:
// half_sums_chain_kernel (staged samples, 17 chains per thread, leaves from eight chains) ...
void hh_half_sums_chain(const float *sig, long long n, long long n_half, int stride, float *half) {
    if (n_half <= 0) return;
    const long long blocks = ((n_half - 1) * stride) / kChainOut + 1;
    std::vector<float> xs(kChainX), cs(kChainP);
    for (long long b = 0; b < blocks; ++b) {
        const long long Q0 = b * kChainOut;
        for (int i = 0; i < kChainX; ++i) xs[i] = Q0 + i < n ? sig[Q0 + i] : 0.0f;
        for (int tid = 0; tid < kChainThreads; ++tid) {
            const int base = (tid & 7) + 8 * kChainT * (tid >> 3);
            float v[kChainT + 15];
            for (int i = 0; i < kChainT + 15; ++i) v[i] = xs[base + 8 * i];
            for (int i = 0; i < kChainT; ++i) {
                float c = v[i];
                for (int m = 1; m < 16; ++m) c = npm::add(c, v[i + m]);
                cs[base + 8 * i] = c;
            }
        }
        const long long u_lo = (Q0 + stride - 1) / stride;
        long long u_hi = (Q0 + kChainOut + stride - 1) / stride;
        if (u_hi > n_half) u_hi = n_half;
        for (long long u = u_lo; u < u_hi; ++u) {
            const int p = (int)(u * stride - Q0);
            half[u] = fwm::half_from_chains(cs[p], cs[p + 1], cs[p + 2], cs[p + 3], cs[p + 4], cs[p + 5], cs[p + 6], cs[p + 7]);
        }
    }
}

// ... and the domain rows of tables_from_halves_kernel<N, DS> (window of half sums per pass of kTabJ domains)
void hh_domains_from_halves(const float *half, long long n_half, long long n_dom, int N, int DS, float *out) {
    const int KS = 256 / DS, HS = 128 / DS, W = kTabJ + (N * 256 - 128) / DS;
    std::vector<float> hs(W);
    for (long long j0 = 0; j0 < n_dom; j0 += kTabJ) {
        for (int i = 0; i < W; ++i) hs[i] = j0 + i < n_half ? half[j0 + i] : 0.0f;
        for (int jl = 0; jl < kTabJ && j0 + jl < n_dom; ++jl)
            for (int k = 0; k < N; ++k)
                out[(j0 + jl) * N + k] = fwm::domain_from_halves(hs[jl + k * KS], hs[jl + k * KS + HS]);
    }
}

void hh_embed(const float *rows, long long n_rows, int N, int emb_dim, float *out) {
    int half = emb_dim / 2;
    FwavEmbedTables t = fwav_make_embed_tables(N, half);
    std::vector<float> tmp(2 * half);
    for (long long r = 0; r < n_rows; ++r) {
        const float *x = rows + r * N;
        auto row = [&](int i) { return x[i]; };
        fwm::embed_row(row, N, half, t.tonal.data(), t.transient.data(), t.w.data(), tmp.data());
        float *o = out + r * emb_dim;
        for (int i = 0; i < emb_dim; ++i) o[i] = i < 2 * half ? tmp[i] : 0.0f;
    }
}

int hh_embed_static(const float *rows, long long n_rows, int N, float *out) {
    if (N == 4) embed_static_rows<4>(rows, n_rows, out);
    else if (N == 8) embed_static_rows<8>(rows, n_rows, out);
    else if (N == 16) embed_static_rows<16>(rows, n_rows, out);
    else if (N == 32) embed_static_rows<32>(rows, n_rows, out);
    else return 1;
    return 0;
}

void hh_embed_tonal(const float *rows, long long n_rows, int N, int k, float *out) {
    FwavEmbedTables t = fwav_make_embed_tables(N, k);
    for (long long r = 0; r < n_rows; ++r) {
        const float *x = rows + r * N;
        auto row = [&](int i) { return x[i]; };
        fwm::embed_tonal_row(row, N, k, t.tonal.data(), out + r * k);
    }
}

void hh_activity(const float *ranges, long long n_r, int N, double thr, int fast, uint8_t *act) {
    for (long long i = 0; i < n_r; ++i) {
        const float *r = ranges + i * N;
        auto row = [&](int k) { return r[k]; };
        bool p;
        if (N == 4) p = fwm::range_is_pruned<4>(row, N, thr, fast);          // same static paths as the kernels
        else if (N == 16) p = fwm::range_is_pruned<16>(row, N, thr, fast);
        else p = fwm::range_is_pruned(row, N, thr, fast);
        act[i] = p ? 0 : 1;
    }
}

void hh_affine(const float *ranges, long long n_r, int N, const float *domains, const int32_t *cand, int K,
               double s_clip, int32_t *idx, float *s, float *o, uint8_t *sym, float *err) {
    const float clipf = (float)std::fabs(s_clip);
    for (long long i = 0; i < n_r; ++i) {
        const float *r = ranges + i * N;
        auto rr = [&](int k) { return r[k]; };
#define HH_MEAN(NS) fwm::range_mean<NS>(rr, N)
        const float r_mean = HH_BY_N(N, HH_MEAN);
        float best = std::numeric_limits<float>::infinity();
        int best_pos = -1;
        fwm::Fit best_fit{0, 0, 0};
        for (int orient = 0; orient < 2; ++orient)
            for (int c = 0; c < K; ++c) {
                int raw = cand[i * K + c];
                int d = raw < 0 ? 0 : raw;
                const float *t = domains + (long long)d * N;
                fwm::Fit f;
#define HH_FIT(NS) fit_dispatch<NS>(rr, r_mean, tt, N)
                if (orient == 0) { auto tt = [&](int k) { return t[k]; }; f = HH_BY_N(N, HH_FIT); }
                else { auto tt = [&](int k) { return t[N - 1 - k]; }; f = HH_BY_N(N, HH_FIT); }
                if (raw < 0) f.err = std::numeric_limits<float>::infinity();
                int pos = orient * K + c;
                if (best_pos < 0 || f.err < best) { best = f.err; best_pos = pos; best_fit = f; }
            }
        int c = best_pos % K;
        int raw = cand[i * K + c];
        idx[i] = raw < 0 ? 0 : raw;
        s[i] = fwm::clip(best_fit.s, -clipf, clipf);
        o[i] = best_fit.o;
        sym[i] = best_pos >= K;
        err[i] = best_fit.err;
    }
}

void hh_affine_pair(const float *ranges, long long n_r, int N, const float *domains, const int32_t *cand, int K,
                    double s_clip, int32_t *idx, float *s, float *o, uint8_t *sym, float *err) {
    const float clipf = (float)std::fabs(s_clip);
#define HH_PAIR(NS) affine_pair_rows<NS>(ranges, n_r, N, domains, cand, K, clipf, idx, s, o, sym, err)
    if (N == 4) HH_PAIR(4); else if (N == 8) HH_PAIR(8); else if (N == 16) HH_PAIR(16); else if (N == 32) HH_PAIR(32); else HH_PAIR(0);
}

// full decoder loop on the CPU with the device's per-range function
int hh_decode(const float *domains, const int32_t *idx, const float *s, const float *o, const uint8_t *sym,
              long long n_r, int N, int iterations, double eps, double s_clip, double s_damping,
              float *out, float *last_delta) {
    std::vector<float> a((size_t)n_r * N, 0.0f), b((size_t)n_r * N, 0.0f);
    float *cur = a.data(), *nxt = b.data();
    const float clipf = (float)std::fabs(s_clip);
    const bool damped = s_damping > 0;
    const float omd = (float)(1.0 - s_damping), dmp = (float)s_damping;
    int it = 0;
    float delta = 0;
    for (; it < iterations;) {
        double dsq = 0, csq = 0;
        for (long long i = 0; i < n_r; ++i) {
            const bool dead = idx[i] < 0;
            const float *t = domains + (long long)(dead ? 0 : idx[i]) * N;
            const bool flip = !dead && sym[i];
            const float *c = cur + i * N;
            float *w = nxt + i * N;
            auto cc = [&](int k) { return c[k]; };
            auto tt = [&](int k) { return dead ? 0.0f : (flip ? t[N - 1 - k] : t[k]); };
            auto put = [&](int k, float v) { w[k] = v; };
#define HH_DEC(NS) (fwm::decode_range<NS>(cc, tt, dead ? 0.0f : s[i], dead ? 0.0f : o[i], N, clipf, damped, omd, dmp, put, &dsq, &csq), 0)
            (void)HH_BY_N(N, HH_DEC);
        }
        delta = fwm::decode_delta(dsq, csq);
        float *tmp = cur; cur = nxt; nxt = tmp;
        ++it;
        if ((double)delta < eps) break;
    }
    std::memcpy(out, cur, sizeof(float) * (size_t)n_r * N);
    *last_delta = delta;
    return it;
}

float hh_score(const float *q, const float *e, int dim) { return fwm::score_chain(q, e, dim); }

// pre-step (A0 / N2) with the device's per-frame functions and the reference's sequential gate:
// mask (n bytes), ranges (ceil(n / N) * N floats), returns the float64 sum of squares of the gated signal
double hh_prestep(const float *sig, long long n, int N, double thr, uint8_t *mask, float *ranges) {
    const int fs = 2 * N;
    const long long n_frames = (n + fs - 1) / fs;
    auto s = [&](long long i) { return sig[i]; };
    std::vector<float> energy(n_frames);
    for (long long f = 0; f < n_frames; ++f) {
        if (fs == 8) energy[f] = fwm::frame_energy<8>(s, f, fs, n);
        else if (fs == 32) energy[f] = fwm::frame_energy<32>(s, f, fs, n);
        else energy[f] = fwm::frame_energy(s, f, fs, n);
    }
    auto e = [&](long long j) { return energy[j]; };
    std::vector<uint8_t> gate(n_frames);
    long long run = -1;                       // running maximum of the keys, as the device scan computes it
    for (long long f = 0; f < n_frames; ++f) {
        const long long k = fwm::gate_key(fwm::smooth5(e, f, n_frames), f, thr);
        if (k > run) run = k;
        gate[f] = run >= 0 && (run & 1);
    }
    for (long long i = 0; i < n; ++i) mask[i] = gate[i / fs];
    const long long n_out = (n + N - 1) / N * N;
    double ssq = 0.0;
    for (long long i = 0; i < n_out; ++i) {
        const long long src = fwm::reflect_index(i, n);
        const float v = npm::mul(sig[src], gate[src / fs] ? 1.0f : 0.0f);
        ranges[i] = v;
        if (i < n) ssq += (double)v * (double)v;
    }
    return ssq;
}

}  // extern "C"

// fwav_math.cuh — per-element math of the FWAV hot path, host + device.
//
// The CUDA kernels are index/launch wrappers around these functions; the same
// functions are compiled for the CPU by tests/csrc/host_harness.cpp so their
// numerics can be pinned to the reference's golden vectors without a GPU.
// Citations are into /root/reference/fractal.py.
#pragma once

#include <math.h>
#include <stdint.h>

#include "np_math.cuh"

namespace fwm {

// Largest range_size any header can describe: tile_size is a u16 (:1296) and
// range_size = max(4, tile_size // 256) (:1070).
constexpr int kMaxRangeSize = 255;

// ---------------------------------------------------------------------------
// A0  pre-step of compress_audio: voiced gate, masking, reflect padding, framing (fractal.py:880-909, 1074-1112).
// ---------------------------------------------------------------------------
// np.pad(x, (0, pad), mode='reflect') with pad < n: position i of the padded array reads x[reflect_index(i, n)]
FWAV_HD long long reflect_index(long long i, long long n) { return i < n ? i : 2 * (n - 1) - i; }

// energy of frame f: np.mean(frames * frames, axis=1) in float32, numpy's pairwise order over the frame (:890-892)
template <int NS = 0, class Sig>
FWAV_HD float frame_energy(Sig sig, long long f, int frame_size, long long n) {
    const long long start = f * frame_size;
    auto sq = [&](int k) { const float v = sig(reflect_index(start + k, n)); return npm::mul(v, v); };
    return npm::mean_n<NS>(sq, frame_size);
}

// np.convolve(energies, ones(5, float32) / 5, mode='same') at frame f, n_frames >= 5 (:894-896).  Interior outputs
// are numpy's small-kernel loop: a float32 multiply-add chain over ascending taps, each step rounded.  The two
// outputs at either end have fewer taps and go through the dot-product routine (OpenBLAS sdot): float32 products
// accumulated in float64, rounded once.  Both forms verified against numpy 2.3 / OpenBLAS 0.3.30.
template <class E>
FWAV_HD float smooth5(E e, long long f, long long n_frames) {
    const float w = npm::div(1.0f, 5.0f);
    const long long lo = f - 2 < 0 ? 0 : f - 2, hi = f + 3 > n_frames ? n_frames : f + 3;
    if (hi - lo == 5) {
        float acc = 0.0f;
        for (long long j = lo; j < hi; ++j) acc = npm::add(acc, npm::mul(e(j), w));
        return acc;
    }
    double d = 0.0;
    for (long long j = lo; j < hi; ++j) d += (double)npm::mul(e(j), w);
    return (float)d;
}

// hysteresis of :901-907 as a scan: a frame whose smoothed energy exceeds the threshold switches the gate on, one
// below the low threshold switches it off, any other keeps the previous state.  Encoded so that the running
// MAXIMUM of the keys carries the latest switch: key = 2 * f + on for a switching frame, -1 otherwise; the gate at
// frame f is (max key up to f) & 1, off while that maximum is still -1.  Thresholds are Python floats cast to
// float32 at the comparison (NEP 50).
FWAV_HD long long gate_key(float smoothed, long long f, double energy_threshold) {
    const float hi = (float)energy_threshold, lo = (float)(energy_threshold * 0.5);
    if (smoothed > hi) return 2 * f + 1;
    if (smoothed < lo) return 2 * f;
    return -1;
}

// ---------------------------------------------------------------------------
// A1  domain value (fractal.py:314-327): mean of `run` consecutive samples.
// run = tile_size // range_size is < 512 for every legal tile_size.
// ---------------------------------------------------------------------------
template <class Sig>
FWAV_HD float domain_value(Sig sig, long long start, int run) {
    auto at = [&](int i) { return sig(start + i); };
    return npm::np_mean<3>(at, run);
}

// Fast path for run == 256: numpy splits 256 = 128 + 128, and each 128-sample
// leaf depends only on its start position, so leaves are computed once per
// position ("half sums") and shared by every output that needs them.
template <class Sig>
FWAV_HD float half_sum128(Sig sig, long long start) {
    auto at = [&](int i) { return sig(start + i); };
    return npm::pairwise_leaf(at, 0, 128);
}
// The leaf's eight strided accumulators are chains that depend on their start alone: r_i of the leaf at p is
// c(p + i), c(q) = x[q] + x[q+8] + ... + x[q+120] added left to right, and the leaf is the tree over c(p..p+7)
// (npm::pairwise_leaf with n = 128).  tables.cu computes every chain once and shares it between the leaves.
template <class Sig>
FWAV_HD float leaf_chain(Sig sig, long long q) {
    float c = sig(q);
    FWAV_UNROLL
    for (int m = 1; m < 16; ++m) c = npm::add(c, sig(q + 8 * m));
    return c;
}
FWAV_HD float half_from_chains(float c0, float c1, float c2, float c3, float c4, float c5, float c6, float c7) {
    return npm::add(npm::add(npm::add(c0, c1), npm::add(c2, c3)), npm::add(npm::add(c4, c5), npm::add(c6, c7)));
}
FWAV_HD float domain_from_halves(float h0, float h1) {
    return npm::mul(npm::add(0.0f, npm::add(h0, h1)), 1.0f / 256.0f);      // == / 256 exactly (a power of two)
}

// ---------------------------------------------------------------------------
// A2/A3  two-head embedding of one row (fractal.py:154-208).
//   tonal_m    : half x N doubles, row k = w[k+1]*c*cos(pi (k+1)(2n+1)/2N); rows >= N-1 are zero
//   transient_m: half x N doubles, row k = c_k*cos(pi k (2n+1)/2N);          rows >= N   are zero
//   w          : N doubles, linspace(1, 2, N)
// out has emb_dim = 2*half floats: [tonal | transient | zeros].
// ---------------------------------------------------------------------------
template <class Row>
FWAV_HD void embed_row(Row x, int N, int half, const double *tonal_m, const double *transient_m,
                       const double *w, float *out) {
    // tonal head: DCT of the raw row, float32 cast, float32 norm (:186-207)
    double ssq = 0.0;
    for (int k = 0; k < half; ++k) {
        double acc = 0.0;
        const double *m = tonal_m + (long long)k * N;
        for (int n = 0; n < N; ++n) acc = fma((double)x(n), m[n], acc);
        float v = (float)acc;
        out[k] = v;
        ssq += (double)v * (double)v;
    }
    float nrm = npm::sqrt((float)ssq);
    if (nrm > 1e-8f)
        for (int k = 0; k < half; ++k) out[k] = npm::div(out[k], nrm);
    // transient head: first difference in float32, weights and DCT in float64 (:156-164)
    const int live = half < N ? half : N;
    double tv[128];
    double tsq = 0.0;
    for (int k = 0; k < live; ++k) {
        double acc = 0.0;
        const double *m = transient_m + (long long)k * N;
        float prev = x(0);
        for (int n = 0; n < N; ++n) {
            float cur = x(n);
            double u = (double)npm::sub(cur, prev) * w[n];
            acc = fma(u, m[n], acc);
            prev = cur;
        }
        tv[k] = acc;
        tsq += acc * acc;
    }
    double tn = ::sqrt(tsq);
    for (int k = 0; k < half; ++k) {
        float v = 0.0f;
        if (k < live) v = (float)(tn > 1e-8 ? tv[k] / tn : tv[k]);
        out[half + k] = v;
    }
}

// tile_embedding(x, k) on its own (fractal.py:178-208, the EMBED_K = 32 form the README describes): the tonal head
// with k coefficients.  tonal_m: k x N doubles as above.
template <class Row>
FWAV_HD void embed_tonal_row(Row x, int N, int k_dim, const double *tonal_m, float *out) {
    double ssq = 0.0;
    for (int k = 0; k < k_dim; ++k) {
        double acc = 0.0;
        const double *m = tonal_m + (long long)k * N;
        for (int n = 0; n < N; ++n) acc = fma((double)x(n), m[n], acc);
        float v = (float)acc;
        out[k] = v;
        ssq += (double)v * (double)v;
    }
    float nrm = npm::sqrt((float)ssq);
    if (nrm > 1e-8f)
        for (int k = 0; k < k_dim; ++k) out[k] = npm::div(out[k], nrm);
}

// ---------------------------------------------------------------------------
// A4  canonical float32 score: one FMA chain in ascending k.
// ---------------------------------------------------------------------------
FWAV_HD float score_chain(const float *q, const float *e, int dim) {
    float acc = 0.0f;
    for (int k = 0; k < dim; ++k) acc = fmaf(q[k], e[k], acc);
    return acc;
}

// ---------------------------------------------------------------------------
// A5  energy prune (fractal.py:602): mean(r*r) < 0.75 * energy_thresh.
// The threshold is a Python float product; numpy (NEP 50) casts it to float32
// before comparing it with the float32 mean.
// ---------------------------------------------------------------------------
template <int NS = 0, class Row>
FWAV_HD bool range_is_pruned(Row r, int N, double energy_thresh, int fast_mode) {
    auto sq = [&](int i) { float v = r(i); return npm::mul(v, v); };
    float m = npm::mean_n<NS>(sq, N);
    return fast_mode && (m < (float)(energy_thresh * 0.75));
}

// ---------------------------------------------------------------------------
// A6  least-squares fit of one (possibly mirrored) candidate tile to one range
// (fractal.py:790-813).  r_mean / r_c are computed once per range by the caller.
// ---------------------------------------------------------------------------
struct Fit {
    float s, o, err;
};

template <int NS = 0, class Row>
FWAV_HD float range_mean(Row r, int N) {
    return npm::mean_n<NS>(r, N);
}

template <int NS = 0, class RowR, class RowT>
FWAV_HD Fit affine_fit(RowR r, float r_mean, RowT t, int N) {
    const float d_mean = npm::mean_n<NS>(t, N);                                   // :796
    auto d_c = [&](int i) { return npm::sub(t(i), d_mean); };                      // :797
    auto r_c = [&](int i) { return npm::sub(r(i), r_mean); };                      // :791
    auto cross = [&](int i) { return npm::mul(d_c(i), r_c(i)); };
    auto self = [&](int i) { float v = d_c(i); return npm::mul(v, v); };
    const float num = npm::sum_n<NS>(cross, N);                                   // :802
    const float den = npm::add(npm::sum_n<NS>(self, N), 1e-12f);                  // :803
    Fit f;
    f.s = npm::div(num, den);                                                     // :804
    f.o = npm::sub(r_mean, npm::mul(f.s, d_mean));                                // :805
    const float s = f.s, o = f.o;
    auto resid = [&](int i) {                                                     // :811-812
        float v = npm::sub(npm::add(npm::mul(s, t(i)), o), r(i));
        return npm::mul(v, v);
    };
    f.err = npm::sqrt(npm::sum_n<NS>(resid, N));                                  // :813
    return f;
}

// Both orientations of one candidate tile (plain, then mirrored: t(N-1-i)) in one go.  rc(i) = r(i) - r_mean is the
// caller's (once per range).  For N = 8 and N = 16 numpy's pairwise sum of the mirrored row adds the same pairs as
// that of the plain row -- the eight accumulators swap places (r'_j = r_(7-j)) and every level of the combining tree
// only sees its two operands exchanged, which IEEE addition does not notice -- so the tile's mean, its centred
// values and its sum of squares are computed once and serve both fits: the same bits as two affine_fit calls with a
// quarter of the work removed.  Other sizes (N = 4: a left-to-right sum; N = 32: four-term chains per accumulator)
// round differently when mirrored and take the two calls.
template <int NS = 0, class RowR, class RowC, class RowT>
FWAV_HD void affine_fit_pair(RowR r, RowC rc, float r_mean, RowT t, int N, Fit &plain, Fit &mirr) {
    if constexpr (NS == 8 || NS == 16) {
        const float d_mean = npm::mean_n<NS>(t, N);                                // :796 (both orientations)
        float dc[NS];
        FWAV_UNROLL
        for (int i = 0; i < NS; ++i) dc[i] = npm::sub(t(i), d_mean);               // :797
        auto self = [&](int i) { return npm::mul(dc[i], dc[i]); };
        const float den = npm::add(npm::sum_n<NS>(self, N), 1e-12f);              // :803
        auto finish = [&](float num, auto tt, Fit &f) {
            f.s = npm::div(num, den);                                              // :804
            f.o = npm::sub(r_mean, npm::mul(f.s, d_mean));                         // :805
            const float s = f.s, o = f.o;
            auto resid = [&](int i) {                                              // :811-812
                float v = npm::sub(npm::add(npm::mul(s, tt(i)), o), r(i));
                return npm::mul(v, v);
            };
            f.err = npm::sqrt(npm::sum_n<NS>(resid, N));                          // :813
        };
        auto cross0 = [&](int i) { return npm::mul(dc[i], rc(i)); };
        auto cross1 = [&](int i) { return npm::mul(dc[NS - 1 - i], rc(i)); };
        auto t1 = [&](int i) { return t(NS - 1 - i); };
        finish(npm::sum_n<NS>(cross0, N), t, plain);                               // :802
        finish(npm::sum_n<NS>(cross1, N), t1, mirr);
    } else {
        const int n_run = NS > 0 ? NS : N;
        auto t1 = [&](int i) { return t(n_run - 1 - i); };
        plain = affine_fit<NS>(r, r_mean, t, N);
        mirr = affine_fit<NS>(r, r_mean, t1, N);
    }
}

FWAV_HD float clip(float v, float lo, float hi) {   // np.clip
    if (v != v) return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

// ---------------------------------------------------------------------------
// A9  one decoder iteration for one range (fractal.py:1414-1461).
//   cur(i)  current reconstruction of the range
//   t(i)    its tile, already mirrored / zeroed for sentinel entries
//   put(i,v) stores the next reconstruction
// Adds the float64 squares of (next - cur) and cur to *dsq and *csq.
// one_minus_damp / damp are float32(1.0 - s_damping) and float32(s_damping).
// ---------------------------------------------------------------------------
template <int NS = 0, class RowC, class RowT, class Put>
FWAV_HD void decode_range(RowC cur, RowT t, float s_st, float o_st, int N, float s_clip,
                          bool damped, float one_minus_damp, float damp, Put put,
                          double *dsq, double *csq) {
    const float mean_d = npm::mean_n<NS>(t, N);                                    // :1431
    const float mean_r = npm::mean_n<NS>(cur, N);                                  // :1434
    auto t_c = [&](int i) { return npm::sub(t(i), mean_d); };
    auto r_c = [&](int i) { return npm::sub(cur(i), mean_r); };
    auto cross = [&](int i) { return npm::mul(r_c(i), t_c(i)); };
    auto self = [&](int i) { float v = t_c(i); return npm::mul(v, v); };
    const float num = npm::sum_n<NS>(cross, N);                                   // :1437
    const float den = npm::sum_n<NS>(self, N);                                    // :1438
    const bool ok = den > 1e-12f;                                                  // :1440
    const float s_opt = ok ? npm::div(num, den) : 0.0f;                            // :1441-1443
    float s_use = damped ? npm::add(npm::mul(one_minus_damp, s_st), npm::mul(damp, s_opt))
                         : (ok ? s_opt : s_st);                                    // :1445
    s_use = clip(s_use, -s_clip, s_clip);                                          // :1446
    double a = 0.0, b = 0.0;
    const int n_run = NS > 0 ? NS : N;
    FWAV_UNROLL
    for (int i = 0; i < n_run; ++i) {
        const float c = cur(i);
        const float v = npm::add(npm::mul(s_use, t(i)), o_st);                     // :1449
        put(i, v);
        const float d = npm::sub(v, c);                                            // :1461
        a += (double)d * (double)d;
        b += (double)c * (double)c;
    }
    *dsq += a;
    *csq += b;
}

// delta of fractal.py:1460-1461 from the two float64 sums of squares.
FWAV_HD float decode_delta(double dsq, double csq) {
    float nd = (float)::sqrt(dsq);
    float nc = (float)::sqrt(csq);
    return npm::div(nd, nc > 0.0f ? nc : 1.0f);
}

}  // namespace fwm

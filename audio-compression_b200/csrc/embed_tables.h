// embed_tables.h — host-side construction of the constant matrices behind the
// two-head embedding (fractal.py:154-208).  Both heads are linear maps of the
// row followed by a norm, so each is one (emb_dim/2 x N) float64 matrix.
#pragma once

#include <cmath>
#include <vector>

struct FwavEmbedTables {
    int N = 0, half = 0;
    std::vector<double> tonal;      // half x N: weighted DCT-II rows 1..half (zero rows past N-1)
    std::vector<double> transient;  // half x N: DCT-II rows 0..half-1     (zero rows past N)
    std::vector<double> w;          // N: numpy.linspace(1, 2, N)
};

inline FwavEmbedTables fwav_make_embed_tables(int N, int half) {
    FwavEmbedTables t;
    t.N = N;
    t.half = half;
    t.w.resize(N);
    const double step = N > 1 ? 1.0 / (double)(N - 1) : 0.0;
    for (int i = 0; i < N; ++i) t.w[i] = (double)i * step + 1.0;
    if (N > 1) t.w[N - 1] = 2.0;
    const double pi = 3.14159265358979323846264338327950288;
    const double c0 = std::sqrt(1.0 / N), c = std::sqrt(2.0 / N);
    t.tonal.assign((size_t)half * N, 0.0);
    t.transient.assign((size_t)half * N, 0.0);
    for (int k = 0; k < half; ++k) {
        const int coef = k + 1;                       // DC dropped (:192-195)
        if (coef <= N - 1)
            for (int n = 0; n < N; ++n)
                t.tonal[(size_t)k * N + n] = t.w[coef] * c * std::cos(pi * coef * (2 * n + 1) / (2.0 * N));
        if (k < N)                                    // DC kept (:160)
            for (int n = 0; n < N; ++n)
                t.transient[(size_t)k * N + n] = (k == 0 ? c0 : c) * std::cos(pi * k * (2 * n + 1) / (2.0 * N));
    }
    return t;
}

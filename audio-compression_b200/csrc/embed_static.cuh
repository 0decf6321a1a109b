// embed_static.cuh — the two-head embedding of one row held in registers (fractal.py:154-208), shared by
// embed.cu (rows from HBM) and tables.cu (rows built from the half sums in the same kernel).
//
// Every coefficient is one float64 FMA chain; the loops run n-outer / k-inner so that the eight chains of a head
// advance side by side (eight independent DFMAs in flight per thread, nothing but the accumulators live): the
// k-outer form kept the whole weighted difference vector and every partial result in registers (128 registers and
// a spilled frame).  Host + device, so tests/test_host_math.py pins it to the reference's embeddings on the CPU.
#pragma once

#include "fwav_math.cuh"

// a float64 product the compiler may not fuse into the addition that follows (host build: -ffp-contract=off)
FWAV_HD double dmul_sep(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

template <int N, int HALF>
struct TablesP {
    double tonal[HALF * N];
    double transient[HALF * N];
    double w[N];
};

// Both heads are DCT-II rows, and cos(pi m (2n+1) / 2N) at n and at N-1-n differ by the factor (-1)^m: an even
// coefficient sees only the sums x[n] + x[N-1-n], an odd one only the differences (both exact in float64), so a
// coefficient is N/2 DFMAs instead of N.  The coefficients differ from the full-length chain of fwm::embed_row by
// float64 rounding noise (1e-16 relative, invisible after the cast to float32 except at a rounding boundary); both
// sit within 6e-7 of the reference's float32 pocketfft result, the gate is 2e-6 (tests/test_host_math.py).
template <int N, int HALF>
FWAV_HD void embed_tonal_static(const float (&x)[N], const TablesP<N, HALF> &T, float (&out)[HALF]) {
    constexpr int H = N / 2;
    double xs[H], xd[H];
    FWAV_UNROLL
    for (int n = 0; n < H; ++n) {
        xs[n] = (double)x[n] + (double)x[N - 1 - n];
        xd[n] = (double)x[n] - (double)x[N - 1 - n];
    }
    double acc[HALF];
    FWAV_UNROLL
    for (int k = 0; k < HALF; ++k) acc[k] = 0.0;
    FWAV_UNROLL
    for (int n = 0; n < H; ++n) {
        FWAV_UNROLL
        for (int k = 0; k < HALF; ++k)             // row k is DCT coefficient k + 1 (DC dropped, :192-195)
            acc[k] = fma(((k + 1) & 1) ? xd[n] : xs[n], T.tonal[k * N + n], acc[k]);
    }
    double ssq = 0.0;
    FWAV_UNROLL
    for (int k = 0; k < HALF; ++k) {
        const float v = (float)acc[k];
        out[k] = v;
        ssq += (double)v * (double)v;
    }
    const float nrm = npm::sqrt((float)ssq);
    if (nrm > 1e-8f) {
        FWAV_UNROLL
        for (int k = 0; k < HALF; ++k) out[k] = npm::div(out[k], nrm);
    }
}

// transient head (:156-164): first difference in float32, weights and DCT rows 0..HALF-1 in float64, float64 norm
template <int N, int HALF>
FWAV_HD void embed_transient_static(const float (&x)[N], const TablesP<N, HALF> &T, float (&out)[HALF]) {
    constexpr int LIVE = HALF < N ? HALF : N;
    constexpr int H = N / 2;
    double us[H], ud[H];
    FWAV_UNROLL
    for (int n = 0; n < H; ++n) {
        const int m = N - 1 - n;
        const double a = n == 0 ? 0.0 : dmul_sep((double)npm::sub(x[n], x[n - 1]), T.w[n]);
        const double b = dmul_sep((double)npm::sub(x[m], x[m - 1]), T.w[m]);
        us[n] = a + b;
        ud[n] = a - b;
    }
    double tv[LIVE];
    FWAV_UNROLL
    for (int k = 0; k < LIVE; ++k) tv[k] = 0.0;
    FWAV_UNROLL
    for (int n = 0; n < H; ++n) {
        FWAV_UNROLL
        for (int k = 0; k < LIVE; ++k)             // row k is DCT coefficient k (DC kept, :160)
            tv[k] = fma((k & 1) ? ud[n] : us[n], T.transient[k * N + n], tv[k]);
    }
    double tsq = 0.0;
    FWAV_UNROLL
    for (int k = 0; k < LIVE; ++k) tsq += tv[k] * tv[k];
    // one float64 reciprocal instead of LIVE divisions (each an inlined Newton chain with its own slow path): the
    // products sit within one float64 ulp of the quotients, far below the float32 cast that follows
    const double tn = sqrt(tsq);
    const double inv = tn > 1e-8 ? 1.0 / tn : 1.0;
    FWAV_UNROLL
    for (int k = 0; k < HALF; ++k) out[k] = k < LIVE ? (float)dmul_sep(tv[k], inv) : 0.0f;
}

template <int N, int HALF>
FWAV_HD void embed_row_static(const float (&x)[N], const TablesP<N, HALF> &T, float (&out)[2 * HALF]) {
    float a[HALF], b[HALF];
    embed_tonal_static<N, HALF>(x, T, a);
    embed_transient_static<N, HALF>(x, T, b);
    FWAV_UNROLL
    for (int k = 0; k < HALF; ++k) { out[k] = a[k]; out[HALF + k] = b[k]; }
}

// embed_static.cuh — the two-head embedding of one row held in registers (fractal.py:154-208), shared by
// embed.cu (rows from HBM) and tables.cu (rows built from the half sums in the same kernel).
//
// Same arithmetic, in the same order per output, as fwm::embed_row: every coefficient k is one float64 FMA chain
// over ascending n.  The loops run n-outer / k-inner so that the eight chains of a head advance side by side
// (eight independent DFMAs in flight per thread, nothing but the accumulators live): the k-outer form kept the
// whole weighted difference vector and every partial result in registers (128 registers and a spilled frame).
#pragma once

#include "fwav_math.cuh"

template <int N, int HALF>
struct TablesP {
    double tonal[HALF * N];
    double transient[HALF * N];
    double w[N];
};

// tonal head (:186-207): DCT rows 1..HALF of the raw row, float32 cast, float32 norm
template <int N, int HALF>
__device__ __forceinline__ void embed_tonal_static(const float (&x)[N], const TablesP<N, HALF> &T, float (&out)[HALF]) {
    double acc[HALF];
#pragma unroll
    for (int k = 0; k < HALF; ++k) acc[k] = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double xn = (double)x[n];
#pragma unroll
        for (int k = 0; k < HALF; ++k) acc[k] = fma(xn, T.tonal[k * N + n], acc[k]);
    }
    double ssq = 0.0;
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
        const float v = (float)acc[k];
        out[k] = v;
        ssq += (double)v * (double)v;
    }
    const float nrm = npm::sqrt((float)ssq);
    if (nrm > 1e-8f) {
#pragma unroll
        for (int k = 0; k < HALF; ++k) out[k] = npm::div(out[k], nrm);
    }
}

// transient head (:156-164): first difference in float32, weights and DCT rows 0..HALF-1 in float64, float64 norm
template <int N, int HALF>
__device__ __forceinline__ void embed_transient_static(const float (&x)[N], const TablesP<N, HALF> &T, float (&out)[HALF]) {
    constexpr int LIVE = HALF < N ? HALF : N;
    double tv[LIVE];
#pragma unroll
    for (int k = 0; k < LIVE; ++k) tv[k] = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double u = n == 0 ? 0.0 * T.w[0] : (double)npm::sub(x[n], x[n - 1]) * T.w[n];
#pragma unroll
        for (int k = 0; k < LIVE; ++k) tv[k] = fma(u, T.transient[k * N + n], tv[k]);
    }
    double tsq = 0.0;
#pragma unroll
    for (int k = 0; k < LIVE; ++k) tsq += tv[k] * tv[k];
    const double tn = sqrt(tsq);
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
        float v = 0.0f;
        if (k < LIVE) v = (float)(tn > 1e-8 ? tv[k] / tn : tv[k]);
        out[k] = v;
    }
}

template <int N, int HALF>
__device__ __forceinline__ void embed_row_static(const float (&x)[N], const TablesP<N, HALF> &T,
                                                 float (&out)[2 * HALF]) {
    float a[HALF], b[HALF];
    embed_tonal_static<N, HALF>(x, T, a);
    embed_transient_static<N, HALF>(x, T, b);
#pragma unroll
    for (int k = 0; k < HALF; ++k) { out[k] = a[k]; out[HALF + k] = b[k]; }
}

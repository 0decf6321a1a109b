// np_math.cuh — float32 arithmetic in numpy's operation order, host + device.
//
// The reference computes every reduction on the path with numpy's float32
// pairwise summation (8 strided accumulators per <=128-element leaf, halves
// rounded down to a multiple of 8 above that) and every element-wise step as a
// separately rounded float32 operation.  Matching that order on the GPU makes
// the domain payload, the affine parameters and the decoded samples
// bit-identical to the reference.
//
// Everything here is __host__ __device__ so tests/test_np_math.py can run the
// very same code on the CPU (compiled by g++ with -ffp-contract=off) against
// numpy before any GPU time is spent.
#pragma once

#if defined(__CUDACC__)
#define FWAV_HD __host__ __device__ __forceinline__
#define FWAV_UNROLL _Pragma("unroll")
#else
#define FWAV_HD inline
#define FWAV_UNROLL
#endif

namespace npm {

// Separately rounded IEEE operations: never contracted into an FMA.
FWAV_HD float add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
FWAV_HD float sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
FWAV_HD float mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
FWAV_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
FWAV_HD float sqrt(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return __builtin_sqrtf(a);
#endif
}

// Leaf of numpy's pairwise sum: n <= 128 (numpy PW_BLOCKSIZE).  `at(i)` yields
// element i already converted to float32.
template <class F>
FWAV_HD float pairwise_leaf(F at, int lo, int n) {
    if (n < 8) {
        float r = -0.0f;  // numpy starts from -0 to keep the sign of an all -0 input
        FWAV_UNROLL
        for (int i = 0; i < n; ++i) r = add(r, at(lo + i));
        return r;
    }
    float r0 = at(lo + 0), r1 = at(lo + 1), r2 = at(lo + 2), r3 = at(lo + 3);
    float r4 = at(lo + 4), r5 = at(lo + 5), r6 = at(lo + 6), r7 = at(lo + 7);
    int i = 8;
    const int body = n - (n % 8);
    FWAV_UNROLL
    for (; i < body; i += 8) {
        r0 = add(r0, at(lo + i + 0));
        r1 = add(r1, at(lo + i + 1));
        r2 = add(r2, at(lo + i + 2));
        r3 = add(r3, at(lo + i + 3));
        r4 = add(r4, at(lo + i + 4));
        r5 = add(r5, at(lo + i + 5));
        r6 = add(r6, at(lo + i + 6));
        r7 = add(r7, at(lo + i + 7));
    }
    float res = add(add(add(r0, r1), add(r2, r3)), add(add(r4, r5), add(r6, r7)));
    FWAV_UNROLL
    for (; i < n; ++i) res = add(res, at(lo + i));
    return res;
}

// numpy pairwise sum for n up to 128 << DEPTH.  The recursion is unrolled at
// compile time so the device code has no call stack.
template <int DEPTH, class F>
FWAV_HD float pairwise(F at, int lo, int n) {
    if constexpr (DEPTH == 0) {
        return pairwise_leaf(at, lo, n);
    } else {
        if (n <= 128) return pairwise_leaf(at, lo, n);
        int half = n / 2;
        half -= half % 8;
        return add(pairwise<DEPTH - 1>(at, lo, half), pairwise<DEPTH - 1>(at, lo + half, n - half));
    }
}

// np.add.reduce of n float32 values: the output starts at the additive identity
// and receives the pairwise sum.
template <int DEPTH, class F>
FWAV_HD float np_sum(F at, int n) {
    return add(0.0f, pairwise<DEPTH>(at, 0, n));
}

// ndarray.mean(dtype=float32): the sum divided by the count in float32.
template <int DEPTH, class F>
FWAV_HD float np_mean(F at, int n) {
    return div(np_sum<DEPTH>(at, n), (float)n);
}

// Compile-time-sized forms (same order): every index is a constant, so callers
// can keep their rows in registers.
template <int LO, int N, class F>
FWAV_HD float pairwise_static(F at) {
    if constexpr (N < 8) {
        float r = -0.0f;
        FWAV_UNROLL
        for (int i = 0; i < N; ++i) r = add(r, at(LO + i));
        return r;
    } else if constexpr (N <= 128) {
        float r[8];
        FWAV_UNROLL
        for (int j = 0; j < 8; ++j) r[j] = at(LO + j);
        constexpr int BODY = N - (N % 8);
        FWAV_UNROLL
        for (int i = 8; i < BODY; i += 8) {
            FWAV_UNROLL
            for (int j = 0; j < 8; ++j) r[j] = add(r[j], at(LO + i + j));
        }
        float res = add(add(add(r[0], r[1]), add(r[2], r[3])), add(add(r[4], r[5]), add(r[6], r[7])));
        FWAV_UNROLL
        for (int i = BODY; i < N; ++i) res = add(res, at(LO + i));
        return res;
    } else {
        constexpr int HALF = (N / 2) - ((N / 2) % 8);
        return add(pairwise_static<LO, HALF>(at), pairwise_static<LO + HALF, N - HALF>(at));
    }
}

// NS > 0: compile-time size NS (n is ignored); NS == 0: run-time size n <= 512.
template <int NS, class F>
FWAV_HD float sum_n(F at, int n) {
    if constexpr (NS > 0) return add(0.0f, pairwise_static<0, NS>(at));
    else return np_sum<2>(at, n);
}
// (a count that is a power of two divides exactly like the multiplication by its reciprocal: both round the same
// real number once, subnormal results included -- and the multiplication is one instruction, not a division sequence)
template <int NS, class F>
FWAV_HD float mean_n(F at, int n) {
    if constexpr (NS > 0 && (NS & (NS - 1)) == 0) return mul(sum_n<NS>(at, n), 1.0f / (float)NS);
    else if constexpr (NS > 0) return div(sum_n<NS>(at, n), (float)NS);
    else return div(sum_n<0>(at, n), (float)n);
}

}  // namespace npm

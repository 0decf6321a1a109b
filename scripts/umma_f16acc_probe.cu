// umma_f16acc_probe.cu — exploratory: tcgen05.mma kind::f16 with HALF-PRECISION accumulators.
// Questions (B200, sm_100a): does D=F16 run at N=256, how do the 256 results of a row sit in TMEM (one per
// 32-bit column, or two packed), what does tcgen05.ld ...pack::16b return, how accurate is the result, and
// how fast is the instruction.  One CTA, one MMA on recognisable data, TMEM dumped raw to the host.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_f16acc_probe umma_f16acc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes contiguous; LBO between the two K chunks, SBO between 8-row groups
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
#define R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
               "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
               "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
               "=r"(v[29]), "=r"(v[30]), "=r"(v[31])

// mode bit 0: D format (0 = f16 accumulators, 1 = f32); bit 1: use tcgen05.ld ... pack::16b for the dump
// A2 / B2 (may be null): a second operand pair issued FIRST (D = A2*B2, then D += A*B) -- the "bias" command
__global__ void __launch_bounds__(160) probe_kernel(const __half *A, const __half *B, uint32_t *dump, long long *cycles,
                                                    int mode, int n_rep, const __half *A2 = nullptr, const __half *B2 = nullptr) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __half *sa = reinterpret_cast<__half *>(smem);             // 128 x 16
    __half *sb = reinterpret_cast<__half *>(smem + 4096);      // 256 x 16
    const uint32_t bar = smem_u32(smem + 4096 + 8192);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 4096 + 8192 + 16);
    __half *sa2 = reinterpret_cast<__half *>(smem + 16384);    // only with A2 (the launch then asks for 28 KB + 64)
    __half *sb2 = reinterpret_cast<__half *>(smem + 16384 + 4096);
    if (A2) {
        for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
            const int r = i / 16, k = i % 16;
            sa2[(k / 8) * (128 * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = A2[i];
        }
        for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
            const int r = i / 16, k = i % 16;
            sb2[(k / 8) * (256 * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = B2[i];
        }
    }
    // stage operands in the canonical no-swizzle K-major layout
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
        const int r = i / 16, k = i % 16;
        sa[(k / 8) * (128 * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = A[i];
    }
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
        const int r = i / 16, k = i % 16;
        sb[(k / 8) * (256 * 8) + (r / 8) * 64 + (r % 8) * 8 + (k % 8)] = B[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (warp == 4 && lane == 0) {
        const uint32_t c_fmt = (mode & 1) ? 1u : 0u;
        const uint32_t idesc = (c_fmt << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = desc_none(smem_u32(sa), 128 * 16, 128), db = desc_none(smem_u32(sb), 256 * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < n_rep; ++i) {
            if (A2) {
                const uint64_t da2 = desc_none(smem_u32(sa2), 128 * 16, 128), db2 = desc_none(smem_u32(sb2), 256 * 16, 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da2), "l"(db2), "r"(idesc), "r"(0)
                             : "memory");
            }
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(A2 ? 1 : 0)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        mbar_wait(bar, 0);
        cycles[0] = clock64() - t0;
    }
    __syncthreads();
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        // dump the first 256 columns of this warp's 32 lanes: dump[row][col]
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            if (mode & 2)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : R32(v) : "r"(ta) : "memory");
            else
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : R32(v) : "r"(ta) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) dump[(warp * 32 + lane) * 256 + c0 + j] = v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

__global__ void max3_h2_probe(const uint32_t *in, uint32_t *out) {
    // does a 3-input packed-half max exist?
    uint32_t a = in[0], b = in[1], c = in[2], d;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(d), "r"(c));
    out[0] = d;
}

// "acc" command (round 2): how far is a half-precision accumulator from the exact dot product?  Random two-unit-head
// rows (what the embeddings are), rounded to fp16 first (so only the tensor core's own arithmetic is measured), many
// launches; also checks what tcgen05.ld ...pack::16b returns against the plain dump.
static int accuracy(int n_launch) {
    static __half hA[128 * 16], hB[256 * 16];
    static uint32_t hD[128 * 256], hP[128 * 256];
    __half *dA, *dB;
    uint32_t *dD;
    long long *dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dC, 8);
    double max_err = 0, max_ulp = 0, sum_err = 0;
    long long n = 0, n_pack_bad = 0, n_near = 0;
    double max_err_near = 0;
    unsigned long long st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    auto gauss = [&]() { double u = rnd() + 1e-12, v = rnd(); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); };
    for (int l = 0; l < n_launch; ++l) {
        auto fill = [&](__half *dst, int rows) {
            for (int r = 0; r < rows; ++r) {
                double x[16], n0 = 0, n1 = 0;
                for (int k = 0; k < 16; ++k) { x[k] = gauss(); (k < 8 ? n0 : n1) += x[k] * x[k]; }
                // every fourth launch: rows that point the same way (scores near 2, where an fp16 ulp is largest)
                for (int k = 0; k < 16; ++k) {
                    double v = x[k] / sqrt(k < 8 ? n0 : n1);
                    if (l % 4 == 3) v = 0.97 * (k % 8 == 0 ? 1.0 : 0.05) + 0.03 * v;
                    dst[r * 16 + k] = __float2half((float)v);
                }
            }
        };
        fill(hA, 128); fill(hB, 256);
        cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
        for (int mode : {0, 2}) {
            probe_kernel<<<1, 160, 4096 + 8192 + 64>>>(dA, dB, dD, dC, mode, 1);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
            cudaMemcpy(mode ? hP : hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        }
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 256; ++c) {
                double ex = 0;
                for (int k = 0; k < 16; ++k) ex += (double)__half2float(hA[r * 16 + k]) * (double)__half2float(hB[c * 16 + k]);
                __half_raw h; h.x = hD[r * 256 + c] & 0xffff;
                const double got = __half2float(__half(h));
                const double err = fabs(got - ex), ulp = ldexp(1.0, (int)floor(log2(fmax(fabs(ex), 6.1e-5))) - 10);
                if (err > max_err) max_err = err;
                if (err / ulp > max_ulp) max_ulp = err / ulp;
                if (fabs(ex) > 1.0) { ++n_near; if (err > max_err_near) max_err_near = err; }
                sum_err += err; ++n;
                // pack::16b dump: the probe loads at column bases 0, 32, 64 ...; the load at base c0 returns, in register j,
                // columns c0 + 2j (low half) and c0 + 2j + 1 (high half)
                const int c0 = (c / 32) * 32, j = c % 32;
                if (c0 + 2 * j + 1 < 256) {
                    const uint32_t pk = hP[r * 256 + c0 + j];
                    const uint32_t lo = hD[r * 256 + c0 + 2 * j] & 0xffff, hi = hD[r * 256 + c0 + 2 * j + 1] & 0xffff;
                    if (pk != (lo | hi << 16)) ++n_pack_bad;
                }
            }
    }
    printf("{\"probe\": \"f16 accumulator accuracy\", \"scores\": %lld, \"max_abs_err\": %.3e, \"max_err_in_fp16_ulps_of_the_result\": %.3f, "
           "\"mean_abs_err\": %.3e, \"scores_above_1\": %lld, \"max_abs_err_above_1\": %.3e, \"pack16b_mismatches\": %lld}\n",
           n, max_err, max_ulp, sum_err / n, n_near, max_err_near, n_pack_bad);
    return 0;
}

// "bias" command (round 2): the threshold inside the accumulator.  D = A2*B2 (row r: -t_r, t_r a half-precision
// number) followed by D += A*B leaves score - t in the accumulator, so "reaches the threshold" is a sign bit.  How far
// is the delivered value from the exact (score - t), and how close to zero must the exact value be for the SIGN to be
// wrong?  Thresholds are taken from the row's own scores (rounded down to fp16), so near-ties are plentiful.
static int bias(int n_launch) {
    static __half hA[128 * 16], hB[256 * 16], hA2[128 * 16], hB2[256 * 16];
    static uint32_t hD[128 * 256];
    __half *dA, *dB, *dA2, *dB2;
    uint32_t *dD;
    long long *dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dA2, sizeof hA); cudaMalloc(&dB2, sizeof hB);
    cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dC, 8);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 28672 + 64);
    unsigned long long st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    auto gauss = [&]() { double u = rnd() + 1e-12, v = rnd(); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); };
    for (int fmt = 0; fmt < 2; ++fmt) {
        double max_err = 0, worst_wrong_sign = 0, max_err_small = 0;
        long long n = 0, n_wrong = 0, n_small = 0;
        for (int l = 0; l < n_launch; ++l) {
            auto fill = [&](__half *dst, int rows) {
                for (int r = 0; r < rows; ++r) {
                    double x[16], n0 = 0, n1 = 0;
                    for (int k = 0; k < 16; ++k) { x[k] = gauss(); (k < 8 ? n0 : n1) += x[k] * x[k]; }
                    for (int k = 0; k < 16; ++k) {
                        double v = x[k] / sqrt(k < 8 ? n0 : n1);
                        if (l % 4 == 3) v = 0.97 * (k % 8 == 0 ? 1.0 : 0.05) + 0.03 * v;
                        dst[r * 16 + k] = __float2half((float)v);
                    }
                }
            };
            fill(hA, 128); fill(hB, 256);
            memset(hA2, 0, sizeof hA2); memset(hB2, 0, sizeof hB2);
            static double thr[128];
            for (int r = 0; r < 128; ++r) {
                // the threshold of row r: the exact score of column (r * 7) % 256, moved by 0 .. 3 fp16 steps, rounded down
                const int c = (r * 7) % 256;
                double ex = 0;
                for (int k = 0; k < 16; ++k) ex += (double)__half2float(hA[r * 16 + k]) * (double)__half2float(hB[c * 16 + k]);
                __half t = __float2half_rd((float)fabs(ex));
                __half_raw tr = t; tr.x = (unsigned short)(tr.x + (r % 4)); t = __half(tr);
                thr[r] = __half2float(t);
                hA2[r * 16] = __float2half(-(float)thr[r]);
            }
            for (int c = 0; c < 256; ++c) hB2[c * 16] = __float2half(1.0f);
            cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
            cudaMemcpy(dA2, hA2, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB2, hB2, sizeof hB, cudaMemcpyHostToDevice);
            probe_kernel<<<1, 160, 28672 + 64>>>(dA, dB, dD, dC, fmt, 1, dA2, dB2);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
            cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < 256; ++c) {
                    double ex = -thr[r];
                    for (int k = 0; k < 16; ++k) ex += (double)__half2float(hA[r * 16 + k]) * (double)__half2float(hB[c * 16 + k]);
                    double got;
                    unsigned sign;
                    if (fmt) { float f; memcpy(&f, &hD[r * 256 + c], 4); got = f; sign = hD[r * 256 + c] >> 31; }
                    else { __half_raw h; h.x = hD[r * 256 + c] & 0xffff; got = __half2float(__half(h)); sign = (h.x >> 15) & 1; }
                    const double err = fabs(got - ex);
                    if (err > max_err) max_err = err;
                    if (fabs(ex) < 4e-3) { ++n_small; if (err > max_err_small) max_err_small = err; }
                    const bool wrong = (sign != 0) != (ex < 0);
                    if (wrong) { ++n_wrong; if (fabs(ex) > worst_wrong_sign) worst_wrong_sign = fabs(ex); }
                    ++n;
                }
        }
        printf("{\"probe\": \"threshold inside the accumulator\", \"accumulators\": \"%s\", \"scores\": %lld, \"max_abs_err\": %.3e, "
               "\"scores_within_4e-3_of_threshold\": %lld, \"max_abs_err_there\": %.3e, \"wrong_signs\": %lld, "
               "\"largest_exact_distance_with_a_wrong_sign\": %.3e}\n",
               fmt ? "f32" : "f16", n, max_err, n_small, max_err_small, n_wrong, worst_wrong_sign);
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "acc")) return accuracy(argc > 2 ? atoi(argv[2]) : 100);
    if (argc > 1 && !strcmp(argv[1], "bias")) return bias(argc > 2 ? atoi(argv[2]) : 100);
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int n_rep = argc > 2 ? atoi(argv[2]) : 1;
    __half hA[128 * 16], hB[256 * 16];
    // A[r][0] = 1, A[r][1] = r/128 : D[r][n] = B[n][0] + (r/128) * B[n][1]
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 16; ++k) hA[r * 16 + k] = __float2half(k == 0 ? 1.0f : k == 1 ? r / 128.0f : 0.0f);
    // B[n][0] = n (exact in fp16 up to 2048), B[n][1] = 1/64
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 16; ++k) hB[n * 16 + k] = __float2half(k == 0 ? (float)n : k == 1 ? 1.0f / 64 : 0.0f);
    __half *dA, *dB;
    uint32_t *dD;
    long long *dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xEE, 128 * 256 * 4);
    probe_kernel<<<1, 160, 4096 + 8192 + 64>>>(dA, dB, dD, dC, mode, n_rep);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t hD[128 * 256];
    long long cyc;
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    printf("mode %d (D=%s, ld %s), %d MMA: %lld cycles\n", mode, (mode & 1) ? "f32" : "f16", (mode & 2) ? "pack::16b" : "plain", n_rep, cyc);
    for (int r : {0, 1, 64, 127}) {
        printf("row %3d raw cols 0..7:", r);
        for (int c = 0; c < 8; ++c) printf(" %08x", hD[r * 256 + c]);
        printf("  | cols 126..131:");
        for (int c = 126; c < 132; ++c) printf(" %08x", hD[r * 256 + c]);
        printf("\n");
        printf("        as halves (lo,hi) cols 0..5:");
        for (int c = 0; c < 6; ++c) {
            __half_raw lo, hi; lo.x = hD[r * 256 + c] & 0xffff; hi.x = hD[r * 256 + c] >> 16;
            printf(" (%g,%g)", __half2float(__half(lo)), __half2float(__half(hi)));
        }
        printf("   as f32 cols 0..3:");
        for (int c = 0; c < 4; ++c) { float f; memcpy(&f, &hD[r * 256 + c], 4); printf(" %g", f); }
        printf("\n");
    }
    return 0;
}

// pipe_microbench.cu — exploratory: which SM pipe takes the packed / scalar maximum instructions the collect pass's
// epilogue is made of, and at what rate (B200, sm_100a).  16 warps per CTA (4 per scheduler), one CTA per SM, eight
// independent accumulators per thread, long unrolled loops; reports warp-instructions per cycle and scheduler.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench pipe_microbench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

enum Op { VIMAX3 = 0, HMAX2, VMAX2, HFMA2RELU, HADD2, FMAX3, FMAX2, FADDABS, MIX_VIMAX3_HMAX2, MIX_VIMAX3_HFMA2, MIX_FMAX3_FADD,
          HSET2, MIX_VIMAX3_HSET2, IADD3, MIX_VIMAX3_IADD3, N_OPS };
static const char *kNames[N_OPS] = {"vimax3_s16x2", "hmax2", "vmaxs2", "hfma2_relu", "hadd2", "fmax3_f32", "fmax2_f32", "fadd_abs_f32",
                                    "mix vimax3+hmax2 1:1", "mix vimax3+hfma2 1:1", "mix fmax3+fadd 1:1", "hset2_gt", "mix vimax3+hset2 1:1",
                                    "iadd3", "mix vimax3+iadd3 1:1"};

__device__ __forceinline__ unsigned op_vimax3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned op_hmax2(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned op_vmax2(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned op_hfma2relu(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned op_hadd2(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned op_hset2(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ float op_fmax3(float a, float b, float c) {
    float d;
    asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float op_fmax2(float a, float b) {
    float d;
    asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float op_faddabs(float a, float b) {
    float d;
    asm volatile("{.reg .f32 t; abs.f32 t, %2; add.f32 %0, %1, t;}" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ unsigned op_iadd3(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int OP>
__global__ void __launch_bounds__(512, 1) bench_kernel(const unsigned *in, unsigned *out, long long *cycles, int iters) {
    unsigned acc[8], x[4];
    for (int i = 0; i < 8; ++i) acc[i] = in[(threadIdx.x + i * 37) & 1023];
    for (int i = 0; i < 4; ++i) x[i] = in[(threadIdx.x * 3 + i) & 1023];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned a = x[(u + i) & 3], b = x[(u + i + 1) & 3];
                if (OP == VIMAX3) acc[i] = op_vimax3(acc[i], a, b);
                else if (OP == HMAX2) acc[i] = op_hmax2(acc[i], a);
                else if (OP == VMAX2) acc[i] = op_vmax2(acc[i], a);
                else if (OP == HFMA2RELU) acc[i] = op_hfma2relu(acc[i], a, b);
                else if (OP == HADD2) acc[i] = op_hadd2(acc[i], a);
                else if (OP == HSET2) acc[i] = op_hset2(acc[i], a);
                else if (OP == IADD3) acc[i] = op_iadd3(acc[i], a, b);
                else if (OP == FMAX3) acc[i] = __float_as_uint(op_fmax3(__uint_as_float(acc[i]), __uint_as_float(a), __uint_as_float(b)));
                else if (OP == FMAX2) acc[i] = __float_as_uint(op_fmax2(__uint_as_float(acc[i]), __uint_as_float(a)));
                else if (OP == FADDABS) acc[i] = __float_as_uint(op_faddabs(__uint_as_float(acc[i]), __uint_as_float(a)));
                else if (OP == MIX_VIMAX3_HMAX2) acc[i] = (i & 1) ? op_hmax2(acc[i], a) : op_vimax3(acc[i], a, b);
                else if (OP == MIX_VIMAX3_HFMA2) acc[i] = (i & 1) ? op_hfma2relu(acc[i], a, b) : op_vimax3(acc[i], a, b);
                else if (OP == MIX_VIMAX3_HSET2) acc[i] = (i & 1) ? op_hset2(acc[i], a) : op_vimax3(acc[i], a, b);
                else if (OP == MIX_VIMAX3_IADD3) acc[i] = (i & 1) ? op_iadd3(acc[i], a, b) : op_vimax3(acc[i], a, b);
                else if (OP == MIX_FMAX3_FADD)
                    acc[i] = (i & 1) ? __float_as_uint(op_faddabs(__uint_as_float(acc[i]), __uint_as_float(a)))
                                     : __float_as_uint(op_fmax3(__uint_as_float(acc[i]), __uint_as_float(a), __uint_as_float(b)));
            }
        }
    }
    const long long t1 = clock64();
    unsigned r = 0;
    for (int i = 0; i < 8; ++i) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const unsigned *d_in, unsigned *d_out, long long *d_cyc, int n_sm, int iters) {
    bench_kernel<OP><<<n_sm, 512>>>(d_in, d_out, d_cyc, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench_kernel<OP><<<n_sm, 512>>>(d_in, d_out, d_cyc, iters);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long *h = (long long *)malloc(sizeof(long long) * n_sm);
    cudaMemcpy(h, d_cyc, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (int i = 0; i < n_sm; ++i) cyc += (double)h[i];
    cyc /= n_sm;
    free(h);
    // warp instructions per scheduler: 4 warps x iters x 32
    const double instr = 4.0 * iters * 32.0;
    printf("{\"op\": \"%s\", \"warp_instr_per_cycle_per_scheduler\": %.4f, \"cycles_per_warp_instr\": %.3f, \"ms\": %.3f, \"err\": \"%s\"}\n",
           kNames[OP], instr / cyc, cyc / instr, ms, cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int n_sm = p.multiProcessorCount, iters = 4096;
    unsigned *h_in = (unsigned *)malloc(4096), *d_in, *d_out;
    long long *d_cyc;
    srand(7);
    for (int i = 0; i < 1024; ++i) {
        const __half a = __float2half((float)(rand() % 2000) / 1000.0f), b = __float2half((float)(rand() % 2000) / 1000.0f);
        h_in[i] = (unsigned)__half_as_ushort(a) | ((unsigned)__half_as_ushort(b) << 16);
    }
    cudaMalloc(&d_in, 4096);
    cudaMalloc(&d_out, sizeof(unsigned) * n_sm * 512);
    cudaMalloc(&d_cyc, sizeof(long long) * n_sm);
    cudaMemcpy(d_in, h_in, 4096, cudaMemcpyHostToDevice);
    run<VIMAX3>(d_in, d_out, d_cyc, n_sm, iters);
    run<HMAX2>(d_in, d_out, d_cyc, n_sm, iters);
    run<VMAX2>(d_in, d_out, d_cyc, n_sm, iters);
    run<HFMA2RELU>(d_in, d_out, d_cyc, n_sm, iters);
    run<HADD2>(d_in, d_out, d_cyc, n_sm, iters);
    run<HSET2>(d_in, d_out, d_cyc, n_sm, iters);
    run<IADD3>(d_in, d_out, d_cyc, n_sm, iters);
    run<FMAX3>(d_in, d_out, d_cyc, n_sm, iters);
    run<FMAX2>(d_in, d_out, d_cyc, n_sm, iters);
    run<FADDABS>(d_in, d_out, d_cyc, n_sm, iters);
    run<MIX_VIMAX3_HMAX2>(d_in, d_out, d_cyc, n_sm, iters);
    run<MIX_VIMAX3_HFMA2>(d_in, d_out, d_cyc, n_sm, iters);
    run<MIX_VIMAX3_HSET2>(d_in, d_out, d_cyc, n_sm, iters);
    run<MIX_VIMAX3_IADD3>(d_in, d_out, d_cyc, n_sm, iters);
    run<MIX_FMAX3_FADD>(d_in, d_out, d_cyc, n_sm, iters);
    return 0;
}

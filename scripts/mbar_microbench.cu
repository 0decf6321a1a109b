// mbar_microbench.cu — exploratory: how long do the mbarrier operations of the hand-over chain take for the thread
// that executes them (B200, sm_100a)?  One warp, lane 0 (or the whole warp), clock64 around N repetitions.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mbar_microbench mbar_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// mode 0: try_wait on a completed phase; 1: test_wait on a completed phase; 2: arrive (count 1: every arrive completes
// a phase) followed by try_wait on that phase (a full round trip inside one thread); 3: arrive only;
// 4: ping-pong between two warps (warp 0 arrives on A, warp 1 waits on A then arrives on B, warp 0 waits on B):
//    two hand-overs per iteration; 5: volatile shared-memory flag ping-pong (st.volatile / ld.volatile spin) for comparison
__global__ void bench(int mode, int iters, long long *out, int whole_warp) {
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ volatile unsigned flag[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t A = smem_u32(&bars[0]), B = smem_u32(&bars[1]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(A), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(B), "r"(1) : "memory");
        flag[0] = flag[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool act = whole_warp || lane == 0;
    long long t0 = 0, t1 = 0;
    unsigned sink = 0;
    if (mode <= 3 && warp == 0 && act) {
        if (mode <= 1) { if (lane == 0) arrive(A); __syncwarp(whole_warp ? 0xffffffffu : 1u); }     // phase 0 of A complete
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 0) sink += try_wait(A, 0);
            else if (mode == 1) sink += test_wait(A, 0);
            else if (mode == 2) { if (lane == 0) arrive(A); while (!try_wait(A, (uint32_t)(i & 1))) { } }
            else if (lane == 0) arrive(A);
        }
        t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = sink; }
    } else if (mode == 4 && warp < 2 && act) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ph = (uint32_t)(i & 1);
            if (warp == 0) { if (lane == 0) arrive(A); while (!try_wait(B, ph)) { } }
            else { while (!try_wait(A, ph)) { } if (lane == 0) arrive(B); }
        }
        t1 = clock64();
        if (lane == 0 && warp == 0) { out[0] = t1 - t0; out[1] = 0; }
    } else if (mode == 5 && warp < 2 && lane == 0) {
        t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            if (warp == 0) { flag[0] = i; while (flag[1] != (unsigned)i) { } }
            else { while (flag[0] != (unsigned)i) { } flag[1] = i; }
        }
        t1 = clock64();
        if (warp == 0) { out[0] = t1 - t0; out[1] = 0; }
    }
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const char *names[] = {"try_wait on a completed phase", "test_wait on a completed phase", "arrive + try_wait round trip in one thread",
                           "arrive", "two-warp ping-pong (two hand-overs per iteration)", "two-warp ping-pong through volatile shared-memory flags"};
    const int iters = 4096;
    for (int ww = 0; ww < 2; ++ww)
        for (int mode = 0; mode < 6; ++mode) {
            if (ww && mode == 5) continue;
            bench<<<1, 64>>>(mode, 16, d, ww);
            bench<<<1, 64>>>(mode, iters, d, ww);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("{\"op\": \"%s\", \"threads\": \"%s\", \"cycles_per_iteration\": %.1f, \"err\": \"%s\"}\n", names[mode],
                   ww ? "whole warp" : "lane 0", (double)h[0] / iters, cudaGetErrorString(e));
        }
    return 0;
}

// domains.cu — A1: domain construction (replaces build_domains_memmap,
// /root/reference/fractal.py:285-334) and the A5 energy-prune flags (:602).
//
// domain[j][k] = float32 mean, in numpy's pairwise order, of the `run` samples
// starting at j*domain_step + k*run, run = tile_size // range_size.
//
// HBM layout: signal (n) f32; domains (n_domains, range_size) f32 row-major —
// exactly the bytes the reference streams into the .fwav payload.
//
// Fast path (run == 256, i.e. every tile_size that is a multiple of 256 and
// >= 1024): numpy reduces 256 = leaf(128) + leaf(128) and a leaf depends only on
// where it starts, so the 128-sample "half sums" are computed once per start
// position (n/gcd(ds,128) of them; tables.cu builds them from shared chains, 15 adds
// per sample) and every output is one add of two half sums and a divide.  That
// turns the kernel pair into a streaming, HBM/L2-bound pass:
//   algorithmic bytes = 4*n read + 4*range_size*n_domains written.
#include "common.cuh"
#include "fwav_math.cuh"

namespace {

// N == 16, run == 256: a block builds 256 consecutive domains per pass.  Column k of those rows needs the half
// sums at (j*ds + k*256) / stride and + 128 / stride for 256 consecutive j: contiguous reads per k; the 256 x 16
// tile is transposed through shared memory so that the rows leave as coalesced float4 stores (16 KB per block).
constexpr int kTileJ = 256;
__global__ void __launch_bounds__(256)
domains_from_halves_t16_kernel(const float *__restrict__ half, long long n_dom, int ds, int stride,
                               float *__restrict__ domains) {
    __shared__ float tile[kTileJ][17];
    const int step = ds / stride;                 // half-sum index advance per domain (stride divides ds)
    const int kstep = 256 / stride, hstep = 128 / stride;
    for (long long j0 = (long long)blockIdx.x * kTileJ; j0 < n_dom; j0 += (long long)gridDim.x * kTileJ) {
        const long long j = j0 + threadIdx.x;
        if (j < n_dom) {
            float h0[16], h1[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {          // all 32 loads in flight before the first add
                const long long h = j * step + (long long)k * kstep;
                h0[k] = __ldg(half + h);
                h1[k] = __ldg(half + h + hstep);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) tile[threadIdx.x][k] = fwm::domain_from_halves(h0[k], h1[k]);
        }
        __syncthreads();
        // 256 rows x 16 floats = 1024 float4: thread t writes float4 t + 256 i (i = 0..3), 4 KB contiguous per i
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = threadIdx.x + 256 * i, jj = f >> 2, q = f & 3;
            if (j0 + jj < n_dom)
                st_stream_f4(reinterpret_cast<float4 *>(domains + (j0 + jj) * 16) + q,
                             make_float4(tile[jj][4 * q], tile[jj][4 * q + 1], tile[jj][4 * q + 2], tile[jj][4 * q + 3]));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
domains_from_halves_kernel(const float *__restrict__ half, long long n_out, int N, int ds,
                           int stride, float *__restrict__ domains) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_out;
         e += (long long)gridDim.x * blockDim.x) {
        const long long j = e / N;
        const int k = (int)(e - j * N);
        const long long start = j * ds + (long long)k * 256;
        const float h0 = __ldg(half + start / stride);
        const float h1 = __ldg(half + (start + 128) / stride);
        domains[e] = fwm::domain_from_halves(h0, h1);
    }
}

__global__ void __launch_bounds__(256)
domains_generic_kernel(const float *__restrict__ signal, long long n_out, int N, int ds, int run,
                       float *__restrict__ domains) {
    auto sig = [&](long long i) { return __ldg(signal + i); };
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_out;
         e += (long long)gridDim.x * blockDim.x) {
        const long long j = e / N;
        const int k = (int)(e - j * N);
        domains[e] = fwm::domain_value(sig, j * ds + (long long)k * run, run);
    }
}

template <int NT>
__global__ void __launch_bounds__(256)
activity_kernel(const float *__restrict__ ranges, long long n_r, int N, double thr, int fast_mode,
                uint8_t *__restrict__ active) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_r;
         i += (long long)gridDim.x * blockDim.x) {
        const float *r = ranges + i * N;
        if constexpr (NT > 0) {
            float v[NT];
#pragma unroll
            for (int k = 0; k < NT; k += 4) {
                float4 q = __ldg(reinterpret_cast<const float4 *>(r + k));
                v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
            }
            auto row = [&](int k) { return v[k]; };
            active[i] = fwm::range_is_pruned<NT>(row, NT, thr, fast_mode) ? 0 : 1;
        } else {
            auto row = [&](int k) { return __ldg(r + k); };
            active[i] = fwm::range_is_pruned(row, N, thr, fast_mode) ? 0 : 1;
        }
    }
}

int gcd_int(int a, int b) {
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}

int grid_for(long long work, int block, int num_sms, int ctas_per_sm) {
    long long need = (work + block - 1) / block;
    long long cap = (long long)num_sms * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace

int fwav_launch_domains(fwav_ctx *ctx, const float *d_signal, int64_t n, int tile, int N, int ds,
                        float *d_domains, cudaStream_t st) {
    FWAV_REQUIRE(ctx, tile > 0 && N > 0 && ds > 0 && tile / N >= 1, "bad geometry tile=%d N=%d ds=%d",
                 tile, N, ds);
    FWAV_REQUIRE(ctx, tile / N < 512, "run length %d outside the numpy-order kernels (<512)", tile / N);
    const int64_t n_dom = fwav_count_domains(n, tile, ds);
    if (n_dom == 0) return FWAV_OK;
    const long long n_out = (long long)n_dom * N;
    const int run = tile / N;
    if (run == 256) {
        const int stride = gcd_int(ds, 128);
        const long long n_half = (n - 128) / stride + 1;
        float *d_half = nullptr;
        int rc = fwav_ws_reserve(ctx, WS_HALF, sizeof(float) * (size_t)n_half, (void **)&d_half);
        if (rc) return rc;
        if ((rc = fwav_launch_half_sums(ctx, d_signal, n, n_half, stride, d_half, st))) return rc;
        if (N == 16 && (reinterpret_cast<uintptr_t>(d_domains) & 15) == 0)
            domains_from_halves_t16_kernel<<<grid_for(n_dom, 256, ctx->num_sms, 8), 256, 0, st>>>(
                d_half, n_dom, ds, stride, d_domains);
        else
            domains_from_halves_kernel<<<grid_for(n_out, 256, ctx->num_sms, 8), 256, 0, st>>>(
                d_half, n_out, N, ds, stride, d_domains);
        FWAV_LAUNCH_CHECK(ctx);
    } else {
        domains_generic_kernel<<<grid_for(n_out, 256, ctx->num_sms, 8), 256, 0, st>>>(
            d_signal, n_out, N, ds, run, d_domains);
        FWAV_LAUNCH_CHECK(ctx);
    }
    return FWAV_OK;
}

int fwav_launch_activity(fwav_ctx *ctx, const float *d_ranges, int64_t n_r, int N, double thr,
                         int fast_mode, uint8_t *d_active, cudaStream_t st) {
    if (n_r == 0) return FWAV_OK;
    FWAV_REQUIRE(ctx, N >= 1 && N <= fwm::kMaxRangeSize, "range_size %d out of range", N);
    const int grid = grid_for(n_r, 256, ctx->num_sms, 8);
    const bool aligned = (reinterpret_cast<uintptr_t>(d_ranges) & 15) == 0;
    if (N == 4 && aligned)
        activity_kernel<4><<<grid, 256, 0, st>>>(d_ranges, n_r, N, thr, fast_mode, d_active);
    else if (N == 16 && aligned)
        activity_kernel<16><<<grid, 256, 0, st>>>(d_ranges, n_r, N, thr, fast_mode, d_active);
    else
        activity_kernel<0><<<grid, 256, 0, st>>>(d_ranges, n_r, N, thr, fast_mode, d_active);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

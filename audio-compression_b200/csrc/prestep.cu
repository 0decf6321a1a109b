// prestep.cu — A0 / row N2: the pre-step of compress_audio on the device (replaces voiced_detection and the
// masking / reflect padding / framing of /root/reference/fractal.py:880-909 and :1074-1112; the reference walks
// the frames in a Python loop).  With it the signal goes to HBM once and the host-buffer entry point takes only
// the raw samples.
//
//   frame_energy_kernel   one thread per frame of 2 * range_size samples: float32 mean of squares, numpy order
//   gate_keys_kernel      5-tap smoothing + thresholds -> one switch key per frame (fwm::gate_key)
//   scan_*                inclusive running maximum of the keys (three launches: block maxima, their scan, carry-in)
//   apply_gate_kernel     ranges[i] = signal[reflect(i)] * gate[frame(reflect(i))]; float64 sum of squares for the
//                         "silent input" early-out (:1083)
//
// Bound: HBM, 4 B/sample read twice (energies, apply) + 4 B/sample written; everything else is per frame.
#include "common.cuh"
#include "fwav_math.cuh"

namespace {

constexpr int kScanBlock = 1024;

template <int FS>
__global__ void __launch_bounds__(256)
frame_energy_kernel(const float *__restrict__ signal, long long n, long long n_frames, int frame_size,
                    float *__restrict__ energy) {
    auto sig = [&](long long i) { return __ldg(signal + i); };
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n_frames; f += (long long)gridDim.x * blockDim.x)
        energy[f] = fwm::frame_energy<FS>(sig, f, frame_size, n);
}

__global__ void __launch_bounds__(256)
gate_keys_kernel(const float *__restrict__ energy, long long n_frames, double thr, long long *__restrict__ keys) {
    auto e = [&](long long j) { return __ldg(energy + j); };
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n_frames; f += (long long)gridDim.x * blockDim.x)
        keys[f] = fwm::gate_key(fwm::smooth5(e, f, n_frames), f, thr);
}

__device__ __forceinline__ long long block_max_scan(long long v, long long *sh, long long *total) {
    // inclusive max-scan over the 1024 threads of a block; returns this thread's prefix maximum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o && u > v) v = u;
    }
    if (lane == 31) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        long long w = sh[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o && u > w) w = u;
        }
        sh[lane] = w;
    }
    __syncthreads();
    if (warp > 0 && sh[warp - 1] > v) v = sh[warp - 1];
    if (total) *total = sh[31];
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(kScanBlock)
scan_block_max_kernel(const long long *__restrict__ keys, long long n_frames, long long *__restrict__ block_max) {
    __shared__ long long sh[32];
    const long long f = (long long)blockIdx.x * kScanBlock + threadIdx.x;
    long long tot;
    block_max_scan(f < n_frames ? keys[f] : -1, sh, &tot);
    if (threadIdx.x == 0) block_max[blockIdx.x] = tot;
}

// exclusive running maximum over the block maxima (what came before each block), one block
__global__ void __launch_bounds__(kScanBlock)
scan_carry_kernel(const long long *__restrict__ block_max, long long n_blocks, long long *__restrict__ carry_out) {
    __shared__ long long sh[32];
    __shared__ long long incl_sh[kScanBlock];
    long long carry = -1;
    for (long long b0 = 0; b0 < n_blocks; b0 += kScanBlock) {
        const long long b = b0 + threadIdx.x;
        long long tot;
        const long long incl = block_max_scan(b < n_blocks ? block_max[b] : -1, sh, &tot);
        incl_sh[threadIdx.x] = incl;
        __syncthreads();
        long long excl = threadIdx.x == 0 ? -1 : incl_sh[threadIdx.x - 1];
        if (carry > excl) excl = carry;
        if (b < n_blocks) carry_out[b] = excl;
        if (tot > carry) carry = tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanBlock)
scan_apply_kernel(const long long *__restrict__ keys, const long long *__restrict__ carry, long long n_frames,
                  uint8_t *__restrict__ gate) {
    __shared__ long long sh[32];
    const long long f = (long long)blockIdx.x * kScanBlock + threadIdx.x;
    long long v = block_max_scan(f < n_frames ? keys[f] : -1, sh, nullptr);
    const long long c = carry[blockIdx.x];
    if (c > v) v = c;
    if (f < n_frames) gate[f] = (v >= 0 && (v & 1)) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
apply_gate_kernel(const float *__restrict__ signal, long long n, const uint8_t *__restrict__ gate, int frame_size,
                  long long n_out, float *__restrict__ ranges, double *__restrict__ partial) {
    double ssq = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
        const long long src = fwm::reflect_index(i, n);
        const float v = npm::mul(__ldg(signal + src), gate[src / frame_size] ? 1.0f : 0.0f);     // signal * mask (:1079)
        ranges[i] = v;
        if (i < n) ssq += (double)v * (double)v;                                                  // :1083, before the padding
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ssq += __shfl_down_sync(0xffffffffu, ssq, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ssq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32) sum_partials_kernel(const double *__restrict__ partial, int n, double *__restrict__ out) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *out = t;
}

}  // namespace

// d_ranges receives ceil(n / N) * N floats; *d_sumsq (device, float64) the sum of squares of the gated signal.
int fwav_launch_prestep(fwav_ctx *ctx, const float *d_signal, int64_t n, int N, double energy_thresh, float *d_ranges,
                        double *d_sumsq, cudaStream_t st) {
    FWAV_REQUIRE(ctx, N >= 1 && N <= fwm::kMaxRangeSize, "range_size %d out of range", N);
    const int fs = 2 * N;
    const long long n_frames = (n + fs - 1) / fs;
    FWAV_REQUIRE(ctx, n_frames >= 5, "the device pre-step needs at least five frames of %d samples (got %lld samples)", fs,
                 (long long)n);
    FWAV_REQUIRE(ctx, n >= fs, "signal shorter than one frame");          // reflect padding stays single-fold
    const long long n_out = (n + N - 1) / N * N;
    const long long n_blocks = (n_frames + kScanBlock - 1) / kScanBlock;
    const int grid_cap = ctx->num_sms * 8;
    unsigned char *ws = nullptr;
    const size_t sz_e = ((size_t)n_frames * 4 + 255) & ~(size_t)255, sz_k = ((size_t)n_frames * 8 + 255) & ~(size_t)255,
                 sz_b = ((size_t)n_blocks * 8 + 255) & ~(size_t)255, sz_g = ((size_t)n_frames + 255) & ~(size_t)255,
                 sz_p = (size_t)grid_cap * 8;
    int rc = fwav_ws_reserve(ctx, WS_PRESTEP, sz_e + sz_k + 2 * sz_b + sz_g + sz_p, (void **)&ws);
    if (rc) return rc;
    long long *d_keys = reinterpret_cast<long long *>(ws);
    long long *d_bmax = reinterpret_cast<long long *>(ws + sz_k);
    long long *d_carry = reinterpret_cast<long long *>(ws + sz_k + sz_b);
    double *d_part = reinterpret_cast<double *>(ws + sz_k + 2 * sz_b);
    float *d_energy = reinterpret_cast<float *>(ws + sz_k + 2 * sz_b + sz_p);
    uint8_t *d_gate = ws + sz_k + 2 * sz_b + sz_p + sz_e;
    auto grid_for = [&](long long work) { long long g = (work + 255) / 256; return (int)(g < grid_cap ? g : grid_cap); };
#define FWAV_FE(FS) frame_energy_kernel<FS><<<grid_for(n_frames), 256, 0, st>>>(d_signal, n, n_frames, fs, d_energy)
    if (fs == 8) FWAV_FE(8);
    else if (fs == 16) FWAV_FE(16);
    else if (fs == 32) FWAV_FE(32);
    else if (fs == 64) FWAV_FE(64);
    else FWAV_FE(0);
#undef FWAV_FE
    FWAV_LAUNCH_CHECK(ctx);
    gate_keys_kernel<<<grid_for(n_frames), 256, 0, st>>>(d_energy, n_frames, energy_thresh, d_keys);
    FWAV_LAUNCH_CHECK(ctx);
    scan_block_max_kernel<<<(unsigned)n_blocks, kScanBlock, 0, st>>>(d_keys, n_frames, d_bmax);
    FWAV_LAUNCH_CHECK(ctx);
    scan_carry_kernel<<<1, kScanBlock, 0, st>>>(d_bmax, n_blocks, d_carry);
    FWAV_LAUNCH_CHECK(ctx);
    scan_apply_kernel<<<(unsigned)n_blocks, kScanBlock, 0, st>>>(d_keys, d_carry, n_frames, d_gate);
    FWAV_LAUNCH_CHECK(ctx);
    const int g = grid_for(n_out);
    apply_gate_kernel<<<g, 256, 0, st>>>(d_signal, n, d_gate, fs, n_out, d_ranges, d_part);
    FWAV_LAUNCH_CHECK(ctx);
    sum_partials_kernel<<<1, 32, 0, st>>>(d_part, g, d_sumsq);
    FWAV_LAUNCH_CHECK(ctx);
    return FWAV_OK;
}

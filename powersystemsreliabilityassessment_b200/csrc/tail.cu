// tail.cu -- tail-risk reduction over per-year ENS on sm_100a.
//
// Reference outputs: the per-year distribution vector and its histogram
// (GeneratingAdequacy/tail_risk.jl:18,84,168-175; Montecarlo_seq/seqMain.m:48,173,287).
// VaR / CVaR do not exist in the reference; the build-side spec is
//   VaR_a  = Julia Statistics.quantile(x, a) default (type 7): position (N-1)a (0-based),
//            linear interpolation between the two bracketing order statistics,
//   CVaR_a = mean(x_i : x_i >= VaR_a).
// Both are computed exactly from the integer per-year values: the order statistics by an
// MSB-first radix select (one privatised 256-bin histogram pass per byte), the tail by a
// count/sum pass.  No sort, no float accumulation on the device.
#include <math.h>

#include <vector>

#include "psra_internal.cuh"

// histogram of byte `shift/8` over the keys whose higher bytes equal `prefix`
__global__ void __launch_bounds__(256) radix_hist_kernel(const unsigned long long *__restrict__ v, long long n,
                                                         unsigned long long prefix, int shift,
                                                         unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long himask = (shift >= 56) ? 0ull : (~0ull << (shift + 8));
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long k = v[i];
        if ((k & himask) == prefix) atomicAdd(&sh[(k >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// count and integer sum of the values >= thr (thr = ceil(VaR): the values are integers)
__global__ void __launch_bounds__(256) tail_sum_kernel(const long long *__restrict__ v, long long n, long long thr,
                                                       unsigned long long *__restrict__ out /* count, sum, max */)
{
    unsigned long long cnt = 0, sum = 0, mx = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long x = v[i];
        if (x >= thr) { cnt++; sum += (unsigned long long)x; }
        if ((unsigned long long)x > mx) mx = (unsigned long long)x;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, d);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) { atomicAdd(&out[0], cnt); atomicAdd(&out[1], sum); }
        atomicMax(&out[2], mx);
    }
}

// fixed-width histogram with a block-private copy in shared memory (bin 0 is hot: ~45 % of
// RTS-79 years have ENS = 0)
__global__ void __launch_bounds__(256) fixed_hist_kernel(const long long *__restrict__ v, long long n,
                                                         long long bin_width, int n_bins,
                                                         unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned int shh[];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) shh[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long b = v[i] / bin_width;
        if (b >= n_bins) b = n_bins - 1;
        if (b < 0) b = 0;
        atomicAdd(&shh[b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (shh[i]) atomicAdd(&hist[i], (unsigned long long)shh[i]);
}

// k-th smallest (0-based) of the device vector by MSB-first radix select
static int radix_select(psra_handle *h, const unsigned long long *d_v, long long n, long long k, int top_shift,
                        unsigned long long *d_hist, int grid, unsigned long long *result)
{
    unsigned long long prefix = 0;
    unsigned long long hist[256];
    for (int shift = top_shift; shift >= 0; shift -= 8) {
        PSRA_CUDA(h, cudaMemsetAsync(d_hist, 0, sizeof(hist), h->stream));
        radix_hist_kernel<<<grid, 256, 0, h->stream>>>(d_v, n, prefix, shift, d_hist);
        PSRA_CUDA(h, cudaGetLastError());
        PSRA_CUDA(h, cudaMemcpyAsync(hist, d_hist, sizeof(hist), cudaMemcpyDeviceToHost, h->stream));
        PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
        int b = 0;
        for (; b < 256; b++) {
            if (k < (long long)hist[b]) break;
            k -= (long long)hist[b];
        }
        if (b == 256) return psra_fail(h, PSRA_E_INVALID, "radix select: rank out of range");
        prefix |= (unsigned long long)b << shift;
    }
    *result = prefix;
    return PSRA_OK;
}

extern "C" int psra_tail(psra_handle *h, const int64_t *values, int64_t n, const double *alphas,
                         int32_t n_alpha, psra_tail_out *out, int64_t *hist, int32_t n_bins, int64_t bin_width)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_CUDA(h, cudaSetDevice(h->device));
    const long long *d_v = nullptr;
    if (values) {
        PSRA_REQUIRE(h, n >= 1, "need at least one value");
        int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(int64_t) * (size_t)n);
        if (rc) return rc;
        PSRA_CUDA(h, cudaMemcpyAsync(h->d_scratch, values, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        d_v = (const long long *)h->d_scratch;
    } else {
        PSRA_REQUIRE(h, h->kept_n >= 1, "no per-year vector kept on the device (run psra_seq_mc with keep_on_device)");
        PSRA_REQUIRE(h, n == 0 || n == h->kept_n, "n does not match the kept vector");
        n = h->kept_n;
        d_v = (const long long *)h->d_ens;
    }
    PSRA_REQUIRE(h, n_alpha >= 0 && (n_alpha == 0 || (alphas && out)), "bad alpha arguments");
    int rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, sizeof(unsigned long long) * (256 + 8 + (size_t)(n_bins > 0 ? n_bins : 0)));
    if (rc) return rc;
    unsigned long long *d_hist = (unsigned long long *)h->d_scratch2, *d_ts = d_hist + 256, *d_fh = d_ts + 8;
    long long want = (n + 255) / 256;
    const int grid = (int)(want < (long long)h->sm_count * 8 ? want : (long long)h->sm_count * 8);

    // maximum (bounds the radix passes) -- also validates non-negativity via the unsigned view
    unsigned long long ts[3];
    PSRA_CUDA(h, cudaMemsetAsync(d_ts, 0, sizeof(unsigned long long) * 8, h->stream));
    tail_sum_kernel<<<grid, 256, 0, h->stream>>>(d_v, n, 0x7fffffffffffffffll, d_ts);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaMemcpyAsync(ts, d_ts, sizeof(ts), cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    PSRA_REQUIRE(h, ts[2] <= 0x7fffffffffffffffull, "values must be non-negative");
    int top_shift = 0;
    while (top_shift < 56 && (ts[2] >> (top_shift + 8))) top_shift += 8;

    for (int a = 0; a < n_alpha; a++) {
        const double alpha = alphas[a];
        PSRA_REQUIRE(h, alpha >= 0.0 && alpha <= 1.0, "alpha must be within [0, 1]");
        const double pos = (double)(n - 1) * alpha;          // type 7, 0-based position
        long long lo = (long long)floor(pos);
        if (lo > n - 1) lo = n - 1;
        const long long hi = lo + 1 < n ? lo + 1 : n - 1;
        const double g = pos - (double)lo;
        unsigned long long xlo = 0, xhi = 0;
        rc = radix_select(h, (const unsigned long long *)d_v, n, lo, top_shift, d_hist, grid, &xlo);
        if (rc) return rc;
        if (hi != lo) {
            rc = radix_select(h, (const unsigned long long *)d_v, n, hi, top_shift, d_hist, grid, &xhi);
            if (rc) return rc;
        } else {
            xhi = xlo;
        }
        const double var = (double)xlo + g * ((double)xhi - (double)xlo);
        const long long thr = (long long)ceil(var);           // integer x >= var  <=>  x >= ceil(var)
        PSRA_CUDA(h, cudaMemsetAsync(d_ts, 0, sizeof(unsigned long long) * 8, h->stream));
        tail_sum_kernel<<<grid, 256, 0, h->stream>>>(d_v, n, thr, d_ts);
        PSRA_CUDA(h, cudaGetLastError());
        PSRA_CUDA(h, cudaMemcpyAsync(ts, d_ts, sizeof(ts), cudaMemcpyDeviceToHost, h->stream));
        PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
        out[a].var = var;
        out[a].n_tail = (int64_t)ts[0];
        out[a].cvar = ts[0] ? (double)ts[1] / (double)ts[0] : var;
        out[a].x_lo = (int64_t)xlo;
        out[a].x_hi = (int64_t)xhi;
    }
    if (hist && n_bins > 0) {
        PSRA_REQUIRE(h, bin_width >= 1 && n_bins <= 8192, "bad histogram shape");
        PSRA_CUDA(h, cudaMemsetAsync(d_fh, 0, sizeof(unsigned long long) * (size_t)n_bins, h->stream));
        fixed_hist_kernel<<<grid, 256, sizeof(unsigned int) * (size_t)n_bins, h->stream>>>(d_v, n, bin_width, n_bins, d_fh);
        PSRA_CUDA(h, cudaGetLastError());
        PSRA_CUDA(h, cudaMemcpyAsync(hist, d_fh, sizeof(int64_t) * (size_t)n_bins, cudaMemcpyDeviceToHost, h->stream));
        PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return PSRA_OK;
}

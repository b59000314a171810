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

#include <algorithm>
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

// ------------------------------------------------------------------------------------------------------------
// Tail risk from the ENS histogram the sequential kernels keep (psra_seq_outputs.tail_hist): hist[e] = number of
// years whose ENS is e fixed-point MWh (1 <= e < bins; the `zeros` years without loss of load are implicit),
// hist[bins] / hist[bins + 1] = number / ENS sum of the years beyond the range.  With one-unit bins the order
// statistics and the tail sums are exact functions of the counts.  ONE launch: every block reduces its chunk of bins
// to {count, sum of e * count}; the last block to finish (ticket) scans the chunk totals, locates for every alpha the
// two order statistics that bracket the type-7 quantile position, and sums the tail >= ceil(VaR).
#define TH_THREADS 256
#define TH_MAX_BLOCKS 1024
#define TH_MAX_ALPHA 8

struct TailHistArgs {
    const unsigned long long *hist;
    long long bins, chunk, years, zeros;
    int n_alpha;
    long long r_lo[TH_MAX_ALPHA], r_hi[TH_MAX_ALPHA];     // 0-based ranks of the bracketing order statistics
    double g[TH_MAX_ALPHA];                                // interpolation weight of the upper one
    unsigned long long *partial;                           // [gridDim.x][2]
    unsigned int *ticket;
    psra_tail_out *out;                                    // [n_alpha]
    int *flag;                                             // != 0: a quantile lies beyond the histogram range
};

__device__ __forceinline__ unsigned long long th_block_sum(unsigned long long v, unsigned long long *sh)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < TH_THREADS / 32; w++) t += sh[w];
    return t;
}

// smallest bin e >= lo with (number of binned years <= e) > r, given `before` = number of binned years < lo;
// block-cooperative, every thread returns the result (or -1 if the bins up to `hi` do not reach rank r)
__device__ long long th_find(const unsigned long long *__restrict__ hist, long long lo, long long hi, unsigned long long before,
                             unsigned long long r, unsigned long long *sh, long long *sh_res)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *sh_res = -1;
    __syncthreads();
    for (long long base = lo; base < hi; base += TH_THREADS) {
        const long long i = base + threadIdx.x;
        const unsigned long long c = i < hi ? hist[i] : 0ull;
        unsigned long long v = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        __syncthreads();
        if (lane == 31) sh[warp] = v;
        __syncthreads();
        unsigned long long pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < TH_THREADS / 32; w++) { const unsigned long long t = sh[w]; tot += t; if (w < warp) pre += t; }
        const unsigned long long incl = before + pre + v;          // binned years <= bin i
        if (c && incl > r && incl - c <= r) *sh_res = i;           // exactly one thread
        __syncthreads();
        if (*sh_res >= 0) break;
        before += tot;
    }
    const long long res = *sh_res;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(TH_THREADS) tail_hist_kernel(const TailHistArgs a)
{
    __shared__ unsigned long long sh[TH_THREADS / 32];
    __shared__ unsigned long long pc[TH_MAX_BLOCKS + 1], ps[TH_MAX_BLOCKS + 1];   // chunk totals -> exclusive prefixes
    __shared__ long long sh_res;
    __shared__ bool is_last;
    const long long lo = (long long)blockIdx.x * a.chunk, hi = min(a.bins, lo + a.chunk);
    unsigned long long cnt = 0, sum = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += TH_THREADS) {
        const unsigned long long c = a.hist[i];
        cnt += c; sum += c * (unsigned long long)i;
    }
    cnt = th_block_sum(cnt, sh);
    sum = th_block_sum(sum, sh);
    if (threadIdx.x == 0) {
        a.partial[2 * blockIdx.x] = cnt; a.partial[2 * blockIdx.x + 1] = sum;
        __threadfence();
        is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    const int nb = (int)gridDim.x;
    if (threadIdx.x == 0) {                         // <= 1024 chunk totals: a serial scan is a few microseconds
        unsigned long long c = 0, s = 0;
        for (int b = 0; b < nb; b++) {
            pc[b] = c; ps[b] = s;
            c += ((volatile unsigned long long *)a.partial)[2 * b];
            s += ((volatile unsigned long long *)a.partial)[2 * b + 1];
        }
        pc[nb] = c; ps[nb] = s;
    }
    __syncthreads();
    const unsigned long long binned = pc[nb], binned_sum = ps[nb];
    const unsigned long long beyond = a.hist[a.bins], beyond_sum = a.hist[a.bins + 1];

    auto order_stat = [&](long long r, bool &ok) -> long long {      // r-th smallest per-year ENS (0-based)
        ok = true;
        if (r < a.zeros) return 0;
        const unsigned long long rr = (unsigned long long)(r - a.zeros);
        if (rr >= binned) { ok = false; return a.bins; }
        int b0 = 0, b1 = nb - 1;                                     // chunk with pc[b] <= rr < pc[b + 1]
        while (b0 < b1) { const int m = (b0 + b1 + 1) >> 1; if (pc[m] <= rr) b0 = m; else b1 = m - 1; }
        const long long clo = (long long)b0 * a.chunk, chi = min(a.bins, clo + a.chunk);
        return th_find(a.hist, clo, chi, pc[b0], rr, sh, &sh_res);
    };

    for (int k = 0; k < a.n_alpha; k++) {
        bool ok1 = true, ok2 = true;
        const long long xlo = order_stat(a.r_lo[k], ok1);
        long long xhi = xlo;
        if (a.r_hi[k] != a.r_lo[k]) xhi = order_stat(a.r_hi[k], ok2);
        const double var = (double)xlo + a.g[k] * ((double)xhi - (double)xlo);
        const long long thr = (long long)ceil(var);                  // integer x >= var  <=>  x >= ceil(var)
        unsigned long long n_tail, s_tail;
        if (thr <= 0) { n_tail = (unsigned long long)a.years; s_tail = binned_sum + beyond_sum; }
        else if (thr >= a.bins) { n_tail = beyond; s_tail = beyond_sum; }
        else {
            const int bt = (int)(thr / a.chunk);
            const long long chi = min(a.bins, ((long long)bt + 1) * a.chunk);
            unsigned long long c = 0, s2 = 0;
            for (long long i = thr + threadIdx.x; i < chi; i += TH_THREADS) {
                const unsigned long long q = a.hist[i];
                c += q; s2 += q * (unsigned long long)i;
            }
            c = th_block_sum(c, sh);
            s2 = th_block_sum(s2, sh);
            n_tail = c + (binned - pc[bt + 1]) + beyond;
            s_tail = s2 + (binned_sum - ps[bt + 1]) + beyond_sum;
        }
        if (threadIdx.x == 0) {
            psra_tail_out o;
            o.var = var; o.n_tail = (int64_t)n_tail; o.x_lo = xlo; o.x_hi = xhi;
            o.cvar = n_tail ? (double)s_tail / (double)n_tail : var;
            a.out[k] = o;
            if (!ok1 || !ok2) *a.flag = 1;
        }
        __syncthreads();
    }
}

// coarse histogram (n_bins of width `w`, last bin open-ended) from the one-unit bins
__global__ void __launch_bounds__(256) tail_rebin_kernel(const unsigned long long *__restrict__ hist, long long bins, long long zeros,
                                                         long long w, int n_bins, unsigned long long *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < bins; i += stride) {
        const unsigned long long c = hist[i];
        if (c) { long long b = i / w; if (b >= n_bins) b = n_bins - 1; atomicAdd(&out[b], c); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (zeros) atomicAdd(&out[0], (unsigned long long)zeros);
        if (hist[bins]) atomicAdd(&out[n_bins - 1], hist[bins]);
    }
}

int psra_tail_hist_prepare(psra_handle *h)
{
    int64_t bins = h->cfg.tail_bins > 0 ? (int64_t)h->cfg.tail_bins
                                        : std::max<int64_t>(1 << 16, std::min<int64_t>(1 << 24, 64 * h->total_cap));
    if (bins != h->tail_bins || !h->d_tail_hist) {
        if (h->d_tail_hist) cudaFree(h->d_tail_hist);
        h->d_tail_hist = nullptr; h->tail_bins = 0;
        PSRA_CUDA(h, cudaMalloc(&h->d_tail_hist, sizeof(unsigned long long) * (size_t)(bins + 2)));
        h->tail_bins = bins;
    }
    h->hist_years = 0;
    PSRA_CUDA(h, cudaMemsetAsync(h->d_tail_hist, 0, sizeof(unsigned long long) * (size_t)(bins + 2), h->stream));
    return PSRA_OK;
}

__global__ void __launch_bounds__(256) tail_hist_used_kernel(const unsigned long long *__restrict__ hist, long long bins,
                                                             unsigned long long *__restrict__ used)
{
    long long last = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < bins; i += stride)
        if (hist[i]) last = i + 1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, last, d); last = o > last ? o : last; }
    if ((threadIdx.x & 31) == 0 && last) atomicMax(used, (unsigned long long)last);
}

int psra_tail_hist_used(psra_handle *h, int64_t *used)
{
    PSRA_REQUIRE(h, h->d_tail_hist && h->tail_bins > 0, "no ENS histogram on the device");
    int rc = psra_reserve(h, &h->d_tail_work, &h->tail_work_cap, 256);
    if (rc) return rc;
    unsigned long long *d_u = (unsigned long long *)h->d_tail_work, u = 0;
    PSRA_CUDA(h, cudaMemsetAsync(d_u, 0, sizeof(u), h->stream));
    const int grid = (int)std::min<long long>((h->tail_bins + 255) / 256, (long long)h->sm_count * 8);
    tail_hist_used_kernel<<<grid, 256, 0, h->stream>>>(h->d_tail_hist, h->tail_bins, d_u);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaMemcpyAsync(&u, d_u, sizeof(u), cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    *used = (int64_t)u;
    return PSRA_OK;
}

static int tail_from_hist(psra_handle *h, const double *alphas, int32_t n_alpha, psra_tail_out *out, int64_t *hist,
                          int32_t n_bins, int64_t bin_width)
{
    PSRA_REQUIRE(h, n_alpha <= TH_MAX_ALPHA, "at most 8 alphas per call on the histogram path");
    const long long N = h->hist_years, zeros = N - h->hist_years_with_loss;
    const long long bins = h->tail_bins;
    int nblk = (int)std::min<long long>(TH_MAX_BLOCKS, (bins + 2047) / 2048);
    long long chunk = ((bins + nblk - 1) / nblk + TH_THREADS - 1) / TH_THREADS * TH_THREADS;
    nblk = (int)((bins + chunk - 1) / chunk);
    const size_t work = sizeof(unsigned long long) * 2 * TH_MAX_BLOCKS + 64 + sizeof(psra_tail_out) * TH_MAX_ALPHA +
                        sizeof(unsigned long long) * (size_t)(n_bins > 0 ? n_bins : 0);
    int rc = psra_reserve(h, &h->d_tail_work, &h->tail_work_cap, work);
    if (rc) return rc;
    unsigned char *wp = (unsigned char *)h->d_tail_work;
    TailHistArgs a{};
    a.hist = h->d_tail_hist; a.bins = bins; a.chunk = chunk; a.years = N; a.zeros = zeros; a.n_alpha = n_alpha;
    a.partial = (unsigned long long *)wp;
    a.ticket = (unsigned int *)(wp + sizeof(unsigned long long) * 2 * TH_MAX_BLOCKS);
    a.flag = (int *)(a.ticket + 1);
    a.out = (psra_tail_out *)(wp + sizeof(unsigned long long) * 2 * TH_MAX_BLOCKS + 64);
    unsigned long long *d_fh = (unsigned long long *)(a.out + TH_MAX_ALPHA);
    for (int k = 0; k < n_alpha; k++) {
        const double alpha = alphas[k];
        PSRA_REQUIRE(h, alpha >= 0.0 && alpha <= 1.0, "alpha must be within [0, 1]");
        const double pos = (double)(N - 1) * alpha;          // type 7, 0-based position
        long long lo = (long long)floor(pos);
        if (lo > N - 1) lo = N - 1;
        a.r_lo[k] = lo; a.r_hi[k] = lo + 1 < N ? lo + 1 : N - 1; a.g[k] = pos - (double)lo;
    }
    if (n_alpha > 0) {
        PSRA_CUDA(h, cudaMemsetAsync(a.ticket, 0, 64, h->stream));
        tail_hist_kernel<<<nblk, TH_THREADS, 0, h->stream>>>(a);
        PSRA_CUDA(h, cudaGetLastError());
    }
    if (hist && n_bins > 0) {
        PSRA_REQUIRE(h, bin_width >= 1 && n_bins <= (1 << 20), "bad histogram shape");
        PSRA_CUDA(h, cudaMemsetAsync(d_fh, 0, sizeof(unsigned long long) * (size_t)n_bins, h->stream));
        const int grid = (int)std::min<long long>((bins + 255) / 256, (long long)h->sm_count * 8);
        tail_rebin_kernel<<<grid, 256, 0, h->stream>>>(h->d_tail_hist, bins, zeros, bin_width, n_bins, d_fh);
        PSRA_CUDA(h, cudaGetLastError());
        PSRA_CUDA(h, cudaMemcpyAsync(hist, d_fh, sizeof(int64_t) * (size_t)n_bins, cudaMemcpyDeviceToHost, h->stream));
    }
    int flag = 0;
    if (n_alpha > 0) {
        PSRA_CUDA(h, cudaMemcpyAsync(out, a.out, sizeof(psra_tail_out) * (size_t)n_alpha, cudaMemcpyDeviceToHost, h->stream));
        PSRA_CUDA(h, cudaMemcpyAsync(&flag, a.flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag)
        return psra_fail(h, PSRA_E_OVERFLOW, "a requested quantile lies beyond the ENS histogram (%lld one-unit bins): raise psra_config.tail_bins",
                         bins);
    return PSRA_OK;
}

extern "C" int psra_tail_hist_export(psra_handle *h, int64_t *counts, int64_t max_bins, int64_t *n_used, int64_t *meta)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, counts && n_used && meta && max_bins >= 1, "null argument");
    PSRA_REQUIRE(h, h->hist_years >= 1 && h->d_tail_hist, "no ENS histogram on the device (run psra_seq_mc with tail_hist)");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    std::vector<unsigned long long> all((size_t)h->tail_bins + 2);
    PSRA_CUDA(h, cudaMemcpy(all.data(), h->d_tail_hist, sizeof(unsigned long long) * all.size(), cudaMemcpyDeviceToHost));
    int64_t used = h->tail_bins;
    while (used > 0 && all[(size_t)used - 1] == 0) used--;
    *n_used = used;
    meta[0] = h->hist_years; meta[1] = h->hist_years_with_loss;
    meta[2] = (int64_t)all[(size_t)h->tail_bins]; meta[3] = (int64_t)all[(size_t)h->tail_bins + 1];
    if (used > max_bins) return psra_fail(h, PSRA_E_OVERFLOW, "histogram has %lld used bins, buffer holds %lld", (long long)used, (long long)max_bins);
    for (int64_t i = 0; i < used; i++) counts[i] = (int64_t)all[(size_t)i];
    return PSRA_OK;
}

extern "C" int psra_tail_hist_import(psra_handle *h, const int64_t *counts, int64_t n, const int64_t *meta)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, (counts || n == 0) && meta && n >= 0, "null argument");
    PSRA_REQUIRE(h, h->U > 0, "psra_set_system has not been called");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    int rc = psra_tail_hist_prepare(h);
    if (rc) return rc;
    PSRA_REQUIRE(h, n <= h->tail_bins, "more bins than the histogram of this handle holds (psra_config.tail_bins)");
    if (n > 0) PSRA_CUDA(h, cudaMemcpyAsync(h->d_tail_hist, counts, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    const unsigned long long beyond[2] = {(unsigned long long)meta[2], (unsigned long long)meta[3]};
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_tail_hist + h->tail_bins, beyond, sizeof(beyond), cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    h->hist_years = meta[0]; h->hist_years_with_loss = meta[1];
    return PSRA_OK;
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
    } else if (h->hist_years >= 1) {
        PSRA_REQUIRE(h, n == 0 || n == h->hist_years, "n does not match the years behind the ENS histogram");
        PSRA_REQUIRE(h, n_alpha >= 0 && (n_alpha == 0 || (alphas && out)), "bad alpha arguments");
        return tail_from_hist(h, alphas, n_alpha, out, hist, n_bins, bin_width);
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

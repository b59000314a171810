// handle.cu -- lifetime, system / load upload (replaces struct Generator / LoadModel,
// GeneratingAdequacy/PowerSystemAdequacy.jl:20-45, for the device path).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <new>
#include <thread>
#include <vector>

#include <algorithm>

#include "psra_internal.cuh"

int psra_fail(psra_handle *h, int code, const char *fmt, ...)
{
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int psra_reserve(psra_handle *h, void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap && *p) return PSRA_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t want = bytes < 256 ? 256 : bytes;
    PSRA_CUDA(h, cudaMalloc(p, want));
    *cap = want;
    return PSRA_OK;
}

int psra_reserve_outputs(psra_handle *h, int64_t n)
{
    if (n <= h->out_cap) return PSRA_OK;
    if (h->d_lol) cudaFree(h->d_lol);
    if (h->d_ens) cudaFree(h->d_ens);
    if (h->d_ent) cudaFree(h->d_ent);
    h->d_lol = nullptr; h->d_ens = nullptr; h->d_ent = nullptr; h->out_cap = 0; h->kept_n = 0;
    PSRA_CUDA(h, cudaMalloc(&h->d_lol, sizeof(uint32_t) * (size_t)n));
    PSRA_CUDA(h, cudaMalloc(&h->d_ens, sizeof(int64_t) * (size_t)n));
    PSRA_CUDA(h, cudaMalloc(&h->d_ent, sizeof(uint32_t) * (size_t)n));
    h->out_cap = n;
    return PSRA_OK;
}

// running mean of the group sums, cum_lole / y of PSA.jl:203,264.  Two coalesced passes: (1) every block sums its
// contiguous chunk; (2) every block adds up the chunk sums in front of it and scans its chunk tile by tile
// (warp-shuffle scan, 1024 elements per tile, running carry).
#define HIST_THREADS 1024
__device__ __forceinline__ long long hist_block_sum(long long v, long long *sh)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < HIST_THREADS / 32; w++) t += sh[w];
    return t;
}

__global__ void __launch_bounds__(HIST_THREADS) history_partial_kernel(const long long *__restrict__ g, long long n, long long chunk,
                                                                       int block0, long long *__restrict__ partial)
{
    __shared__ long long sh[HIST_THREADS / 32];
    const int blk = block0 + (int)blockIdx.x;
    const long long lo = (long long)blk * chunk, hi = min(n, lo + chunk);
    long long s = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += HIST_THREADS) s += g[i];
    s = hist_block_sum(s, sh);
    if (threadIdx.x == 0) partial[blk] = s;
}

__global__ void __launch_bounds__(HIST_THREADS) history_scan_kernel(const long long *__restrict__ g, long long n, long long chunk,
                                                                    int block0, const long long *__restrict__ partial, int group,
                                                                    long long carry0, long long idx0, double *__restrict__ out)
{
    __shared__ long long sh[HIST_THREADS / 32];
    __shared__ long long wtot[HIST_THREADS / 32];
    const int blk = block0 + (int)blockIdx.x;
    long long carry = 0;
    for (int b = threadIdx.x; b < blk; b += HIST_THREADS) carry += partial[b];
    carry = hist_block_sum(carry, sh) + carry0;          // carry0 / idx0: what lies in front of this device's range (multi-GPU)
    const long long lo = (long long)blk * chunk, hi = min(n, lo + chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = lo; base < hi; base += HIST_THREADS) {
        const long long i = base + threadIdx.x;
        long long v = i < hi ? g[i] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        __syncthreads();
        if (lane == 31) wtot[warp] = v;
        __syncthreads();
        long long before = 0, total = 0;
        for (int w = 0; w < HIST_THREADS / 32; w++) { const long long t = wtot[w]; total += t; if (w < warp) before += t; }
        if (i < hi) out[i] = (double)(carry + before + v) / ((double)group * (double)(idx0 + i + 1));
        carry += total;
    }
}

// groups per scan block: a multiple of the 1024-element tile, about four blocks per SM over the whole sequence
int64_t psra_history_block(const psra_handle *h, int64_t nfull)
{
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((nfull + HIST_THREADS - 1) / HIST_THREADS, 4 * (int64_t)h->sm_count));
    return ((nfull + blocks - 1) / blocks + HIST_THREADS - 1) / HIST_THREADS * HIST_THREADS;
}

int psra_history_prepare(psra_handle *h, int64_t nfull)
{
    if (nfull <= 0) return PSRA_OK;
    const int64_t chunk = psra_history_block(h, nfull), blocks = (nfull + chunk - 1) / chunk;
    const size_t pin = sizeof(double) * (size_t)nfull;
    if (h->pin_cap < pin) {
        if (h->h_pin) cudaFreeHost(h->h_pin);
        h->h_pin = nullptr; h->pin_cap = 0;
        const size_t want = std::max(pin + pin / 4, (size_t)1 << 20);
        PSRA_CUDA(h, cudaHostAlloc(&h->h_pin, want, cudaHostAllocDefault));
        h->pin_cap = want;
    }
    h->pin_pending.clear();
    return psra_reserve(h, &h->d_hist, &h->hist_cap, sizeof(double) * (size_t)nfull + sizeof(long long) * (size_t)blocks);
}

int psra_history_range(psra_handle *h, const long long *d_group, int64_t nfull, int group, int64_t b0, int64_t b1,
                       double *history, cudaStream_t stream)
{
    if (nfull <= 0) return PSRA_OK;
    const int64_t chunk = psra_history_block(h, nfull), blocks = (nfull + chunk - 1) / chunk;
    b1 = std::min(b1, blocks);
    if (b1 <= b0) return PSRA_OK;
    double *d_hist = (double *)h->d_hist;
    long long *d_part = (long long *)(d_hist + nfull);
    history_partial_kernel<<<(unsigned)(b1 - b0), HIST_THREADS, 0, stream>>>(d_group, nfull, chunk, (int)b0, d_part);
    history_scan_kernel<<<(unsigned)(b1 - b0), HIST_THREADS, 0, stream>>>(d_group, nfull, chunk, (int)b0, d_part, group, h->hist_carry0, h->hist_idx0, d_hist);
    PSRA_CUDA(h, cudaGetLastError());
    const int64_t g0 = b0 * chunk, g1 = std::min(nfull, b1 * chunk);
    const size_t slot = h->pin_pending.size();
    if (slot >= sizeof(h->ev_pin) / sizeof(h->ev_pin[0]))
        return psra_fail(h, PSRA_E_CUDA, "internal error: too many staged history ranges");
    double *pin = (double *)h->h_pin + g0;
    PSRA_CUDA(h, cudaMemcpyAsync(pin, d_hist + g0, sizeof(double) * (size_t)(g1 - g0), cudaMemcpyDeviceToHost, stream));
    PSRA_CUDA(h, cudaEventRecord(h->ev_pin[slot], stream));
    h->pin_pending.push_back({history + g0, pin, sizeof(double) * (size_t)(g1 - g0), h->ev_pin[slot]});
    return PSRA_OK;
}

int psra_history_drain(psra_handle *h)
{
    for (const psra_handle::PinCopy &c : h->pin_pending) {
        PSRA_CUDA(h, cudaEventSynchronize(c.ev));
        memcpy(c.dst, c.src, c.bytes);
    }
    h->pin_pending.clear();
    return PSRA_OK;
}

int psra_history_to_host(psra_handle *h, const long long *d_group, int64_t nfull, int group, double *history)
{
    if (nfull <= 0) return PSRA_OK;
    int rc = psra_history_prepare(h, nfull);
    if (rc) return rc;
    rc = psra_history_range(h, d_group, nfull, group, 0, INT64_MAX / 2, history, h->stream);
    if (rc) return rc;
    return psra_history_drain(h);
}

extern "C" int psra_version(void) { return PSRA_VERSION; }

extern "C" const char *psra_last_error(const psra_handle *h) { return h ? h->err : "null handle"; }

extern "C" uint64_t psra_stream(const psra_handle *h) { return h ? (uint64_t)(uintptr_t)h->stream : 0; }

extern "C" int psra_device_info(const psra_handle *h, int32_t *sm_count, int32_t *sm_clock_khz)
{
    if (!h) return PSRA_E_INVALID;
    if (sm_count) *sm_count = h->sm_count;
    if (sm_clock_khz) *sm_clock_khz = h->sm_clock_khz;
    return PSRA_OK;
}

extern "C" int psra_last_counters(const psra_handle *h, uint64_t *out, int32_t n)
{
    if (!h || !out || n < 0) return PSRA_E_INVALID;
    for (int i = 0; i < n; i++) out[i] = i < ACC_COUNT_MAX ? h->last_acc[i] : 0;
    return PSRA_OK;
}

extern "C" int psra_create(psra_handle **out, const psra_config *cfg)
{
    if (!out) return PSRA_E_INVALID;
    *out = nullptr;
    psra_handle *h = new (std::nothrow) psra_handle();
    if (!h) return PSRA_E_INVALID;
    *out = h;  // returned even on failure so that psra_last_error() can be read
    if (cfg) h->cfg = *cfg;
    h->device = h->cfg.device;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return psra_fail(h, PSRA_E_CUDA, "no CUDA device available (%s); libpsra_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (h->device < 0 || h->device >= ndev)
        return psra_fail(h, PSRA_E_INVALID, "device %d out of range (%d devices)", h->device, ndev);
    PSRA_CUDA(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    PSRA_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10)
        return psra_fail(h, PSRA_E_CUDA, "device %s is sm_%d%d; this library holds sm_100a code only",
                         prop.name, prop.major, prop.minor);
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
    h->sm_clock_khz = khz;
    PSRA_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
        // the scans of finished history ranges must not queue behind the Monte Carlo blocks that are still waiting for an SM
        int lo = 0, hi = 0;
        PSRA_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PSRA_CUDA(h, cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi));
        PSRA_CUDA(h, cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
        PSRA_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    for (int i = 0; i < PSRA_MAX_CHUNKS; i++) PSRA_CUDA(h, cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming));
    for (int i = 0; i < PSRA_MAX_CHUNKS + 2; i++) PSRA_CUDA(h, cudaEventCreateWithFlags(&h->ev_pin[i], cudaEventDisableTiming));
    PSRA_CUDA(h, cudaEventCreate(&h->ev0));
    PSRA_CUDA(h, cudaEventCreate(&h->ev1));
    PSRA_CUDA(h, cudaMalloc(&h->d_acc, sizeof(unsigned long long) * ACC_COUNT));
    PSRA_CUDA(h, cudaMalloc(&h->d_redo, sizeof(unsigned long long) * (1 + 4096)));
    PSRA_CUDA(h, cudaMalloc(&h->d_red, sizeof(unsigned long long) * 32));
    if (h->cfg.ngpus > 1) return psra_multi_create(h);
    return PSRA_OK;
}

extern "C" void psra_destroy(psra_handle *h)
{
    if (!h) return;
    psra_multi_destroy(h);
    cudaSetDevice(h->device);
    void *bufs[] = {h->d_cap, h->d_mttf, h->d_mttr, h->d_for_thr, h->d_for, h->d_load, h->d_lmax,
                    h->d_load_sorted, h->d_load_suffix, h->d_acc, h->d_lol, h->d_ens, h->d_ent, h->d_fail,
                    h->d_group, h->d_scratch, h->d_scratch2, h->d_order, h->d_hist, h->d_lol_tab, h->d_byte_tab, h->d_wide_tab,
                    h->d_redo, h->d_tail_hist, h->d_tail_work, h->d_red};
    for (void *p : bufs)
        if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < PSRA_MAX_CHUNKS; i++) if (h->ev_chunk[i]) cudaEventDestroy(h->ev_chunk[i]);
    for (int i = 0; i < PSRA_MAX_CHUNKS + 2; i++) if (h->ev_pin[i]) cudaEventDestroy(h->ev_pin[i]);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// run `fn(peer)` for every peer device of a multi-GPU handle on its own host thread (uploads with their synchronisation
// are ~0.15 ms per device when done one after the other)
template <typename F>
static int for_each_peer(psra_handle *h, F fn)
{
    if (h->peers.empty()) return PSRA_OK;
    std::vector<int> rc(h->peers.size(), PSRA_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < h->peers.size(); i++) th.emplace_back([&, i]() { rc[i] = fn(h->peers[i]); });
    for (auto &t : th) t.join();
    for (size_t i = 0; i < rc.size(); i++)
        if (rc[i]) {
            char msg[400];
            snprintf(msg, sizeof(msg), "%.390s", h->peers[i]->err);
            return psra_fail(h, rc[i], "device %d: %s", h->peers[i]->device, msg);
        }
    return cudaSetDevice(h->device) == cudaSuccess ? PSRA_OK : psra_fail(h, PSRA_E_CUDA, "cudaSetDevice failed");
}

extern "C" int psra_set_system(psra_handle *h, const int32_t *cap_fp, const double *mttf_h,
                               const double *mttr_h, int32_t n_units)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, cap_fp && mttf_h && mttr_h, "null system arrays");
    PSRA_REQUIRE(h, n_units >= 1 && n_units <= (1 << 20), "unit count out of range");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    const int U = n_units;
    std::vector<float> mf(U), mr(U);
    std::vector<uint32_t> thr(U);
    std::vector<double> q(U);
    int64_t total = 0;
    int32_t max_unit = 0;
    double rate = 0.0;
    for (int u = 0; u < U; u++) {
        PSRA_REQUIRE(h, cap_fp[u] >= 0, "negative capacity");
        // upper bound: event times are 64-bit ticks of 2^-24 h and hour indices 32 bits (a duration is < 2^56 ticks)
        PSRA_REQUIRE(h, mttf_h[u] > 0 && mttr_h[u] > 0 && mttf_h[u] <= PSRA_MAX_MEAN_HOURS && mttr_h[u] <= PSRA_MAX_MEAN_HOURS,
                     "MTTF / MTTR must be positive (at most 1e8 hours)");
        total += cap_fp[u];
        if (cap_fp[u] > max_unit) max_unit = cap_fp[u];
        rate += 2.0 / (mttf_h[u] + mttr_h[u]);
        mf[u] = (float)mttf_h[u];
        mr[u] = (float)mttr_h[u];
        // PSA.jl:32-37: lambda = 1/MTTF, mu = 1/MTTR, FOR = lambda/(lambda+mu)
        const double lam = 1.0 / mttf_h[u], mu = 1.0 / mttr_h[u];
        q[u] = lam / (lam + mu);
        double t = floor(q[u] * 4294967296.0);
        thr[u] = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    }
    PSRA_REQUIRE(h, total <= 0x3fffffff, "installed capacity exceeds the int32 fixed-point range");
    if (U != h->U || !h->d_cap) {      // (re)allocate only when the unit count changes
        void *old[] = {h->d_cap, h->d_mttf, h->d_mttr, h->d_for_thr, h->d_for, h->d_order, h->d_wide_tab};
        for (void *p : old)
            if (p) cudaFree(p);
        h->d_cap = nullptr; h->d_mttf = nullptr; h->d_mttr = nullptr; h->d_for_thr = nullptr; h->d_for = nullptr; h->d_order = nullptr; h->d_wide_tab = nullptr;
        h->U = 0;
        PSRA_CUDA(h, cudaMalloc(&h->d_cap, sizeof(int32_t) * U));
        PSRA_CUDA(h, cudaMalloc(&h->d_mttf, sizeof(float) * U));
        PSRA_CUDA(h, cudaMalloc(&h->d_mttr, sizeof(float) * U));
        PSRA_CUDA(h, cudaMalloc(&h->d_for_thr, sizeof(uint32_t) * U));
        PSRA_CUDA(h, cudaMalloc(&h->d_for, sizeof(double) * U));
        PSRA_CUDA(h, cudaMalloc(&h->d_order, sizeof(int32_t) * ((U + 31) & ~31)));      // padded to whole groups of 32 (seq_wide.cu)
        PSRA_CUDA(h, cudaMalloc(&h->d_wide_tab, sizeof(uint4) * ((U + 31) & ~31)));
    }
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_cap, cap_fp, sizeof(int32_t) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_mttf, mf.data(), sizeof(float) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_mttr, mr.data(), sizeof(float) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_for_thr, thr.data(), sizeof(uint32_t) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_for, q.data(), sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    const int Upad = (U + 31) & ~31;
    std::vector<int32_t> order(Upad, 0);
    for (int u = 0; u < U; u++) order[u] = u;
    std::stable_sort(order.begin(), order.begin() + U, [&](int32_t x, int32_t y) { return mttf_h[x] + mttr_h[x] < mttf_h[y] + mttr_h[y]; });
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_order, order.data(), sizeof(int32_t) * Upad, cudaMemcpyHostToDevice, h->stream));
    h->cycle_sorted.assign(U, 0.0);
    for (int k = 0; k < U; k++) h->cycle_sorted[k] = mttf_h[order[k]] + mttr_h[order[k]];
    // one 16-byte record per queue position for seq_wide.cu (the means already in ticks: a power-of-two scaling, exact)
    // (positions U .. Upad - 1: capacity 0; the kernel keeps their events out of the year)
    std::vector<uint4> wtab(Upad, make_uint4(0u, 0x4B800000u, 0x4B800000u, 0u));
    for (int k = 0; k < U; k++) {
        const int u = order[k];
        const float mup = mf[u] * 16777216.0f, mdn = mr[u] * 16777216.0f;
        uint32_t bu, bd;
        memcpy(&bu, &mup, 4); memcpy(&bd, &mdn, 4);
        wtab[k] = make_uint4((uint32_t)cap_fp[u], bu, bd, thr[u]);
    }
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_wide_tab, wtab.data(), sizeof(uint4) * Upad, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    h->U = U;
    h->total_cap = total;
    h->max_unit_cap = max_unit;
    h->events_per_hour = rate;
    h->tab_valid = false;
    h->hist_years = 0;
    // multi-GPU handle: every device holds the system
    return for_each_peer(h, [&](psra_handle *p) { return psra_set_system(p, cap_fp, mttf_h, mttr_h, n_units); });
}

extern "C" int psra_set_load(psra_handle *h, const int32_t *load_fp, int32_t n_hours)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, load_fp, "null load");
    PSRA_REQUIRE(h, n_hours >= 1 && n_hours <= (1 << 20), "hour count out of range");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    const int H = n_hours, Wd = (H + 31) / 32;
    std::vector<int32_t> pad((size_t)Wd * 32, 0), lmax(Wd, 0);
    int32_t mx = 0;
    for (int i = 0; i < H; i++) {
        PSRA_REQUIRE(h, load_fp[i] >= 0 && load_fp[i] <= 0x3fffffff, "load out of the int32 fixed-point range");
        pad[i] = load_fp[i];
        if (load_fp[i] > lmax[i >> 5]) lmax[i >> 5] = load_fp[i];
        if (load_fp[i] > mx) mx = load_fp[i];
    }
    if (Wd != h->Wd || !h->d_load) {   // (re)allocate only when the padded length changes
        if (h->d_load) cudaFree(h->d_load);
        if (h->d_lmax) cudaFree(h->d_lmax);
        h->d_load = nullptr; h->d_lmax = nullptr; h->H = 0; h->Wd = 0;
        PSRA_CUDA(h, cudaMalloc(&h->d_load, sizeof(int32_t) * pad.size()));
        PSRA_CUDA(h, cudaMalloc(&h->d_lmax, sizeof(int32_t) * Wd));
    }
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_load, pad.data(), sizeof(int32_t) * pad.size(), cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_lmax, lmax.data(), sizeof(int32_t) * Wd, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    h->H = H; h->Wd = Wd; h->max_load = mx;
    h->tab_valid = false;
    return for_each_peer(h, [&](psra_handle *p) { return psra_set_load(p, load_fp, n_hours); });
}

// multi.cu -- one handle, several B200s (psra_config.ngpus > 1): the year / sample range of psra_seq_mc and
// psra_nonseq_mc is cut into contiguous shards, one per device of this process, and the small results are
// combined with NCCL over NVLink.
//
// The path has no data-path exchange: every chain / sample owns its Philox streams keyed on its GLOBAL index,
// so the per-year integers do not depend on the number of devices.  What is exchanged, once per call:
//   * the integer accumulators (13 values, the 128-bit sum of ENS^2 as 32-bit limbs)    ncclAllReduce(sum, uint64)
//   * the per-hour failure counts fail_count[H] (tail_risk.jl:81,88)                      ncclAllReduce(sum, uint32)
//   * the per-year ENS histogram up to its last used bin (tail_risk.jl:168-175)          ncclAllReduce(sum, uint64)
// after which every device holds the totals (psra_tail then runs on device 0).  Per-year vectors are
// disjoint ranges of the caller's buffers and are written by the devices directly; the running-mean history
// (PowerSystemAdequacy.jl:263-265) is scanned per device with the LOL hours of the devices in front of it as carry.
// One host thread per device drives the single-device engine (run_seq / run_nonseq), so the kernels of all
// devices run concurrently; NCCL is loaded at run time (libnccl.so.2), the library itself does not link it.
#include <dlfcn.h>

#include <algorithm>
#include <numeric>
#include <chrono>
#include <thread>
#include <vector>

#include "psra_internal.cuh"

// ---- the few NCCL entry points used, resolved with dlopen (types as in nccl.h 2.x)
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;                 // ncclSuccess = 0
enum { NCCL_UINT32 = 3, NCCL_UINT64 = 5, NCCL_SUM = 0 };

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    char why[256] = {0};
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    // a library of that name the process has already loaded (a host that runs torch.distributed) is reused
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) { snprintf(api.why, sizeof(api.why), "%s", dlerror()); return &api; }
#define PSRA_SYM(field, name)                                                                      \
    *(void **)(&api.field) = dlsym(api.lib, name);                                                 \
    if (!api.field) { snprintf(api.why, sizeof(api.why), "symbol %s not found", name); api.lib = nullptr; return &api; }
    PSRA_SYM(GetVersion, "ncclGetVersion")
    PSRA_SYM(CommInitAll, "ncclCommInitAll")
    PSRA_SYM(CommDestroy, "ncclCommDestroy")
    PSRA_SYM(AllReduce, "ncclAllReduce")
    PSRA_SYM(GroupStart, "ncclGroupStart")
    PSRA_SYM(GroupEnd, "ncclGroupEnd")
    PSRA_SYM(GetErrorString, "ncclGetErrorString")
#undef PSRA_SYM
    return &api;
}

#define PSRA_NCCL(h, call)                                                                         \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != 0)                                                                              \
            return psra_fail((h), PSRA_E_NCCL, "%s failed: %s (%s:%d)", #call,                     \
                             nccl_api()->GetErrorString(r__), __FILE__, __LINE__);                 \
    } while (0)

static int n_devices(const psra_handle *h) { return 1 + (int)h->peers.size(); }
static psra_handle *device_of(psra_handle *h, int g) { return g == 0 ? h : h->peers[(size_t)g - 1]; }

int psra_multi_create(psra_handle *h)
{
    const int G = h->cfg.ngpus;
    int ndev = 0;
    PSRA_CUDA(h, cudaGetDeviceCount(&ndev));
    if (h->device + G > ndev)
        return psra_fail(h, PSRA_E_INVALID, "ngpus = %d from device %d, but the process sees %d device(s)", G, h->device, ndev);
    NcclApi *api = nccl_api();
    if (!api->lib) return psra_fail(h, PSRA_E_NCCL, "ngpus = %d needs NCCL: libnccl.so.2 could not be loaded (%s)", G, api->why);
    for (int g = 1; g < G; g++) {
        psra_config c = h->cfg;
        c.device = h->device + g;
        c.ngpus = 1;
        psra_handle *p = nullptr;
        const int rc = psra_create(&p, &c);
        if (rc) {
            psra_fail(h, rc, "device %d: %s", c.device, p ? p->err : "psra_create failed");
            if (p) psra_destroy(p);
            return rc;
        }
        h->peers.push_back(p);
    }
    std::vector<int> devs(G);
    for (int g = 0; g < G; g++) devs[g] = h->device + g;
    ncclComm_t *comms = new ncclComm_t[G]();
    h->nccl_comms = comms;
    PSRA_NCCL(h, api->CommInitAll(comms, G, devs.data()));
    return PSRA_OK;
}

void psra_multi_destroy(psra_handle *h)
{
    if (h->nccl_comms) {
        ncclComm_t *comms = (ncclComm_t *)h->nccl_comms;
        NcclApi *api = nccl_api();
        for (int g = 0; g < n_devices(h); g++)
            if (comms[g] && api->lib) api->CommDestroy(comms[g]);
        delete[] comms;
        h->nccl_comms = nullptr;
    }
    for (psra_handle *p : h->peers) psra_destroy(p);
    h->peers.clear();
}

// element-wise sum over the devices of one buffer per device (same count), result on every device
static int all_reduce(psra_handle *h, void *const *bufs, size_t count, int dtype)
{
    NcclApi *api = nccl_api();
    ncclComm_t *comms = (ncclComm_t *)h->nccl_comms;
    const int G = n_devices(h);
    PSRA_NCCL(h, api->GroupStart());
    for (int g = 0; g < G; g++) {
        psra_handle *d = device_of(h, g);
        PSRA_CUDA(h, cudaSetDevice(d->device));
        PSRA_NCCL(h, api->AllReduce(bufs[g], bufs[g], count, dtype, NCCL_SUM, comms[g], d->stream));
    }
    PSRA_NCCL(h, api->GroupEnd());
    return PSRA_OK;
}

static int sync_all(psra_handle *h)
{
    for (int g = 0; g < n_devices(h); g++) {
        psra_handle *d = device_of(h, g);
        PSRA_CUDA(h, cudaSetDevice(d->device));
        PSRA_CUDA(h, cudaStreamSynchronize(d->stream));
    }
    PSRA_CUDA(h, cudaSetDevice(h->device));
    return PSRA_OK;
}

// accumulators -> 16 uint64 that can be summed element-wise over any number of devices without losing a carry
// (the 128-bit sum of squares travels as four 32-bit limbs)
#define RED_N 16
static void pack_red(unsigned long long *v, long long n, long long lol, long long ens, long long ent, long long ywl,
                     unsigned long long lol2, unsigned long long e2lo, unsigned long long e2hi, unsigned long long events,
                     long long redone, unsigned long long beyond_n, unsigned long long beyond_sum)
{
    v[0] = (unsigned long long)n; v[1] = (unsigned long long)lol; v[2] = (unsigned long long)ens; v[3] = (unsigned long long)ent;
    v[4] = (unsigned long long)ywl; v[5] = lol2;
    v[6] = e2lo & 0xffffffffull; v[7] = e2lo >> 32; v[8] = e2hi & 0xffffffffull; v[9] = e2hi >> 32;
    v[10] = events; v[11] = (unsigned long long)redone; v[12] = beyond_n; v[13] = beyond_sum; v[14] = 0; v[15] = 0;
}

static void unpack_e2(const unsigned long long *v, uint64_t *lo, uint64_t *hi)
{
    // limbs may exceed 32 bits after the sum: propagate
    unsigned long long l0 = v[6], l1 = v[7], l2 = v[8], l3 = v[9];
    l1 += l0 >> 32; l0 &= 0xffffffffull;
    l2 += l1 >> 32; l1 &= 0xffffffffull;
    l3 += l2 >> 32; l2 &= 0xffffffffull;
    *lo = l0 | (l1 << 32);
    *hi = l2 | (l3 << 32);
}

static long long lcm_ll(long long a, long long b) { return a / std::gcd(a, b) * b; }

// first error of the per-device calls -> the root handle
static int first_error(psra_handle *h, const std::vector<int> &rc)
{
    for (int g = 0; g < (int)rc.size(); g++)
        if (rc[(size_t)g] != PSRA_OK) {
            psra_handle *d = device_of(h, g);
            if (g == 0) return rc[0];                      // the message is already in h->err
            char msg[400];
            snprintf(msg, sizeof(msg), "%s", d->err);
            return psra_fail(h, rc[(size_t)g], "device %d: %s", d->device, msg);
        }
    return PSRA_OK;
}

// PSRA_TRACE=1 in the environment: wall-clock phases of a multi-device call on stderr (host side; the kernels' own time is
// psra_seq_summary.kernel_ms)
static bool trace_on()
{
    static const bool on = getenv("PSRA_TRACE") != nullptr;
    return on;
}
static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Running-mean history of a multi-device call: every device scans its groups (carry = LOL hours of the devices in front of
// it) and stages them through its pinned buffer; one host thread per device, so the scans, the link transfers and the host
// copies into the caller's buffer run side by side.
static int multi_history(psra_handle *h, const std::vector<long long> &idx0, const std::vector<long long> &nf,
                         const std::vector<long long> &carry, long long group, double *history)
{
    const int G = n_devices(h);
    std::vector<int> rc((size_t)G, PSRA_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        if (nf[(size_t)g] <= 0) continue;
        th.emplace_back([&, g]() {
            psra_handle *d = device_of(h, g);
            int r = PSRA_OK;
            if (cudaSetDevice(d->device) != cudaSuccess) r = psra_fail(d, PSRA_E_CUDA, "cudaSetDevice failed");
            d->hist_carry0 = carry[(size_t)g]; d->hist_idx0 = idx0[(size_t)g];
            if (r == PSRA_OK) r = psra_history_prepare(d, nf[(size_t)g]);
            if (r == PSRA_OK) r = psra_history_range(d, d->d_group, nf[(size_t)g], (int)group, 0, INT64_MAX / 2, history + idx0[(size_t)g], d->stream);
            if (r == PSRA_OK) r = psra_history_drain(d);
            d->hist_carry0 = 0; d->hist_idx0 = 0;
            rc[(size_t)g] = r;
        });
    }
    for (auto &t : th) t.join();
    return first_error(h, rc);
}

int psra_multi_seq_mc(psra_handle *h, int64_t year0, int64_t nyears, uint64_t seed, int32_t init_mode, int32_t ypc,
                      const psra_seq_outputs *out, psra_seq_summary *summary)
{
    PSRA_REQUIRE(h, summary != nullptr, "summary must not be NULL");
    PSRA_REQUIRE(h, h->U > 0 && h->H > 0, "psra_set_system / psra_set_load have not been called");
    PSRA_REQUIRE(h, !(out && out->keep_on_device), "keep_on_device is per device: a multi-GPU handle keeps the ENS histogram instead (tail_hist)");
    const int G = n_devices(h);
    memset(summary, 0, sizeof(*summary));
    summary->years = nyears;
    h->kept_n = 0; h->hist_years = 0;
    if (nyears == 0) return PSRA_OK;
    const bool want_group = out && (out->group_lol || out->history);
    PSRA_REQUIRE(h, !want_group || out->group >= 1, "group must be >= 1");
    const long long group = want_group ? out->group : 1;
    const long long nchains = nyears / ypc, chain0 = year0 / ypc;
    // shard = whole chains and whole history groups
    const long long unit = lcm_ll(group, ypc) / ypc;
    const long long per = ((nchains + G - 1) / G + unit - 1) / unit * unit;
    std::vector<long long> c0((size_t)G), cn((size_t)G);
    for (int g = 0; g < G; g++) {
        c0[(size_t)g] = std::min(nchains, (long long)g * per);
        cn[(size_t)g] = std::min(per, nchains - c0[(size_t)g]);
    }
    const bool want_fail = out && out->fail_count, want_hist = out && out->tail_hist;

    const double t_begin = trace_on() ? now_ms() : 0.0;
    double t_touch = 0.0, t_join = 0.0, t_reduce = 0.0, t_hist = 0.0;
    std::vector<psra_seq_summary> sums((size_t)G);
    std::vector<int> rc((size_t)G, PSRA_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        th.emplace_back([&, g]() {
            psra_handle *d = device_of(h, g);
            const long long y0 = c0[(size_t)g] * ypc;
            psra_seq_outputs og{};
            if (out) {
                og = *out;
                if (out->lol_hours) og.lol_hours = out->lol_hours + y0;
                if (out->ens_fp) og.ens_fp = out->ens_fp + y0;
                if (out->entries) og.entries = out->entries + y0;
                if (out->group_lol) og.group_lol = out->group_lol + y0 / group;
            }
            d->multi_defer = true;
            int r = PSRA_OK;
            if (cn[(size_t)g] > 0) {
                r = psra_run_seq_range(d, chain0 + c0[(size_t)g], cn[(size_t)g], ypc, init_mode, seed, out ? &og : nullptr, &sums[(size_t)g]);
            } else {                                         // more devices than shards: contribute zeros
                memset(&sums[(size_t)g], 0, sizeof(psra_seq_summary));
                if (cudaSetDevice(d->device) != cudaSuccess) r = psra_fail(d, PSRA_E_CUDA, "cudaSetDevice failed");
                if (r == PSRA_OK && want_fail) r = psra_seq_prepare_fail(d);
                if (r == PSRA_OK && want_hist) r = psra_tail_hist_prepare(d);
            }
            d->multi_defer = false;
            rc[(size_t)g] = r;
        });
    }
    // While the devices compute, this thread touches every page of the caller's history buffer (a fresh allocation costs
    // a page fault per 4 KB at its first write: ~2 ms per 10 MB, which would otherwise sit behind the kernels; the buffer
    // is an output that the call overwrites completely).
    if (out && out->history && nyears / group > 0) memset(out->history, 0, sizeof(double) * (size_t)(nyears / group));
    if (trace_on()) t_touch = now_ms();
    for (auto &t : th) t.join();
    if (trace_on()) t_join = now_ms();
    int e = first_error(h, rc);
    if (e) return e;

    // ---- accumulators (+ the histogram's beyond-range slots), per-hour failure counts, ENS histogram: NCCL sums
    std::vector<void *> bufs((size_t)G);
    int64_t used = 0;
    for (int g = 0; g < G; g++) {
        psra_handle *d = device_of(h, g);
        PSRA_CUDA(h, cudaSetDevice(d->device));
        const psra_seq_summary &s = sums[(size_t)g];
        unsigned long long beyond[2] = {0ull, 0ull};
        if (want_hist) {
            PSRA_CUDA(h, cudaMemcpy(beyond, d->d_tail_hist + d->tail_bins, sizeof(beyond), cudaMemcpyDeviceToHost));
            int64_t u = 0;
            const int r = psra_tail_hist_used(d, &u);
            if (r) return psra_fail(h, r, "device %d: %s", d->device, d->err);
            used = std::max(used, u);
        }
        unsigned long long v[RED_N];
        pack_red(v, s.years, s.sum_lol_hours, s.sum_ens_fp, s.sum_entries, s.years_with_loss, s.sum_lol_sq, s.sum_ens_sq_lo,
                 s.sum_ens_sq_hi, s.events, s.redone, beyond[0], beyond[1]);
        PSRA_CUDA(h, cudaMemcpyAsync(d->d_red, v, sizeof(v), cudaMemcpyHostToDevice, d->stream));
        PSRA_CUDA(h, cudaStreamSynchronize(d->stream));      // v is a stack array
        bufs[(size_t)g] = d->d_red;
    }
    e = all_reduce(h, bufs.data(), RED_N, NCCL_UINT64);
    if (e) return e;
    if (want_fail) {
        for (int g = 0; g < G; g++) bufs[(size_t)g] = device_of(h, g)->d_fail;
        e = all_reduce(h, bufs.data(), (size_t)h->Wd * 32, NCCL_UINT32);
        if (e) return e;
    }
    if (want_hist && used > 0) {
        for (int g = 0; g < G; g++) bufs[(size_t)g] = device_of(h, g)->d_tail_hist;
        e = all_reduce(h, bufs.data(), (size_t)used, NCCL_UINT64);
        if (e) return e;
    }
    unsigned long long tot[RED_N];
    PSRA_CUDA(h, cudaSetDevice(h->device));
    PSRA_CUDA(h, cudaMemcpyAsync(tot, h->d_red, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
    if (want_fail) PSRA_CUDA(h, cudaMemcpyAsync(out->fail_count, h->d_fail, sizeof(uint32_t) * (size_t)h->H, cudaMemcpyDeviceToHost, h->stream));

    if (trace_on()) t_reduce = now_ms();
    // ---- running mean of the LOL hours (PSA.jl:263-265): every device scans its groups, carry = LOL hours in front of it
    if (out && out->history) {
        const long long nfull = nyears / group;
        long long carry = 0;
        std::vector<long long> hi0((size_t)G), hnf((size_t)G), hcarry((size_t)G);
        for (int g = 0; g < G; g++) {
            hi0[(size_t)g] = c0[(size_t)g] * ypc / group;
            const long long ng = (cn[(size_t)g] * ypc + group - 1) / group;
            hnf[(size_t)g] = std::max(0ll, std::min(ng, nfull - hi0[(size_t)g]));
            hcarry[(size_t)g] = carry;
            carry += sums[(size_t)g].sum_lol_hours;
        }
        e = multi_history(h, hi0, hnf, hcarry, group, out->history);
        if (e) return e;
        PSRA_CUDA(h, cudaSetDevice(h->device));
    }
    e = sync_all(h);
    if (e) return e;
    if (trace_on()) {
        t_hist = now_ms();
        double kmax = 0.0;
        for (int g = 0; g < G; g++) kmax = std::max(kmax, (double)sums[(size_t)g].kernel_ms);
        fprintf(stderr, "[psra] seq_mc on %d devices, %lld years: first touch of the history buffer %.2f ms | devices done after %.2f ms (kernels %.2f ms) | "
                        "all-reduce %.2f ms | history scan + read-back %.2f ms | total %.2f ms\n", G, (long long)nyears, t_touch - t_begin,
                t_join - t_begin, kmax, t_reduce - t_join, t_hist - t_reduce, t_hist - t_begin);
    }

    summary->years = (int64_t)tot[0];
    summary->sum_lol_hours = (int64_t)tot[1];
    summary->sum_ens_fp = (int64_t)tot[2];
    summary->sum_entries = (int64_t)tot[3];
    summary->years_with_loss = (int64_t)tot[4];
    summary->sum_lol_sq = tot[5];
    unpack_e2(tot, &summary->sum_ens_sq_lo, &summary->sum_ens_sq_hi);
    summary->events = tot[10];
    summary->redone = (int32_t)std::min<unsigned long long>(tot[11], 0x7fffffffull);
    for (int g = 0; g < G; g++) summary->kernel_ms = std::max(summary->kernel_ms, sums[(size_t)g].kernel_ms);
    if (summary->years != nyears) return psra_fail(h, PSRA_E_NCCL, "all-reduce returned %lld years, expected %lld", (long long)summary->years, (long long)nyears);
    if (want_hist) {
        // every device now holds the histogram of all years; psra_tail works on device 0
        for (int g = 0; g < G; g++) {
            psra_handle *d = device_of(h, g);
            const unsigned long long beyond[2] = {tot[12], tot[13]};
            PSRA_CUDA(h, cudaSetDevice(d->device));
            PSRA_CUDA(h, cudaMemcpy(d->d_tail_hist + d->tail_bins, beyond, sizeof(beyond), cudaMemcpyHostToDevice));
            d->hist_years = nyears; d->hist_years_with_loss = summary->years_with_loss;
        }
        PSRA_CUDA(h, cudaSetDevice(h->device));
    }
    memset(h->last_acc, 0, sizeof(h->last_acc));
    h->last_acc[ACC_LOL] = tot[1]; h->last_acc[ACC_ENS] = tot[2]; h->last_acc[ACC_ENT] = tot[3]; h->last_acc[ACC_YWL] = tot[4];
    h->last_acc[ACC_EVENTS] = tot[10]; h->last_acc[ACC_REDO] = tot[11];
    return PSRA_OK;
}

int psra_multi_nonseq_mc(psra_handle *h, int64_t sample0, int64_t n, uint64_t seed, const psra_nonseq_outputs *out,
                         psra_nonseq_summary *summary)
{
    PSRA_REQUIRE(h, summary != nullptr, "summary must not be NULL");
    PSRA_REQUIRE(h, h->U > 0 && h->H > 0, "psra_set_system / psra_set_load have not been called");
    PSRA_REQUIRE(h, n >= 0 && sample0 >= 0, "negative sample range");
    const int G = n_devices(h);
    memset(summary, 0, sizeof(*summary));
    summary->samples = n;
    if (n == 0) return PSRA_OK;
    const bool want_group = out && (out->group_lol || out->history);
    PSRA_REQUIRE(h, !want_group || out->group >= 1, "group must be >= 1");
    const long long group = want_group ? out->group : 1;
    const long long per = ((n + G - 1) / G + group - 1) / group * group;
    const int W = (h->U + 31) / 32;
    std::vector<long long> i0((size_t)G), cn((size_t)G);
    for (int g = 0; g < G; g++) {
        i0[(size_t)g] = std::min((long long)n, (long long)g * per);
        cn[(size_t)g] = std::min(per, (long long)n - i0[(size_t)g]);
    }
    std::vector<psra_nonseq_summary> sums((size_t)G);
    std::vector<int> rc((size_t)G, PSRA_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        th.emplace_back([&, g]() {
            psra_handle *d = device_of(h, g);
            const long long o = i0[(size_t)g];
            psra_nonseq_outputs og{};
            if (out) {
                og = *out;
                if (out->lol_hours) og.lol_hours = out->lol_hours + o;
                if (out->ens_fp) og.ens_fp = out->ens_fp + o;
                if (out->cap_avail) og.cap_avail = out->cap_avail + o;
                if (out->states) og.states = out->states + o * W;
                if (out->group_lol) og.group_lol = out->group_lol + o / group;
            }
            d->multi_defer = true;
            rc[(size_t)g] = psra_run_nonseq_range(d, sample0 + o, cn[(size_t)g], seed, out ? &og : nullptr, &sums[(size_t)g]);
            d->multi_defer = false;
        });
    }
    if (out && out->history && n / group > 0) memset(out->history, 0, sizeof(double) * (size_t)(n / group));   // see psra_multi_seq_mc
    for (auto &t : th) t.join();
    int e = first_error(h, rc);
    if (e) return e;

    std::vector<void *> bufs((size_t)G);
    for (int g = 0; g < G; g++) {
        psra_handle *d = device_of(h, g);
        PSRA_CUDA(h, cudaSetDevice(d->device));
        const psra_nonseq_summary &s = sums[(size_t)g];
        unsigned long long v[RED_N];
        pack_red(v, s.samples, s.sum_lol_hours, s.sum_ens_fp, 0, s.samples_with_loss, s.sum_lol_sq, s.sum_ens_sq_lo, s.sum_ens_sq_hi,
                 0, 0, 0, 0);
        PSRA_CUDA(h, cudaMemcpyAsync(d->d_red, v, sizeof(v), cudaMemcpyHostToDevice, d->stream));
        PSRA_CUDA(h, cudaStreamSynchronize(d->stream));
        bufs[(size_t)g] = d->d_red;
    }
    e = all_reduce(h, bufs.data(), RED_N, NCCL_UINT64);
    if (e) return e;
    unsigned long long tot[RED_N];
    PSRA_CUDA(h, cudaSetDevice(h->device));
    PSRA_CUDA(h, cudaMemcpyAsync(tot, h->d_red, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
    if (out && out->history) {
        const long long nfull = n / group;
        long long carry = 0;
        std::vector<long long> hi0((size_t)G), hnf((size_t)G), hcarry((size_t)G);
        for (int g = 0; g < G; g++) {
            hi0[(size_t)g] = i0[(size_t)g] / group;
            const long long ng = (cn[(size_t)g] + group - 1) / group;
            hnf[(size_t)g] = std::max(0ll, std::min(ng, nfull - hi0[(size_t)g]));
            hcarry[(size_t)g] = carry;
            carry += sums[(size_t)g].sum_lol_hours;
        }
        e = multi_history(h, hi0, hnf, hcarry, group, out->history);
        if (e) return e;
        PSRA_CUDA(h, cudaSetDevice(h->device));
    }
    e = sync_all(h);
    if (e) return e;
    summary->samples = (int64_t)tot[0];
    summary->sum_lol_hours = (int64_t)tot[1];
    summary->sum_ens_fp = (int64_t)tot[2];
    summary->samples_with_loss = (int64_t)tot[4];
    summary->sum_lol_sq = tot[5];
    unpack_e2(tot, &summary->sum_ens_sq_lo, &summary->sum_ens_sq_hi);
    for (int g = 0; g < G; g++) summary->kernel_ms = std::max(summary->kernel_ms, sums[(size_t)g].kernel_ms);
    if (summary->samples != n) return psra_fail(h, PSRA_E_NCCL, "all-reduce returned %lld samples, expected %lld", (long long)summary->samples, (long long)n);
    return PSRA_OK;
}

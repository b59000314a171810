// detailed_mc.cu -- hourly-resampled Monte Carlo with maintenance windows, load-forecast uncertainty
// and energy-limited units.
//
// Replaces run_detailed_mc (GeneratingAdequacy/tail_risk.jl:12-91) == run_monte_carlo
// (MCvsMarkovProcess.jl:210-284) == generating_adequancy_comparative.jl:15-120:
// per year the ELU energy state is reset; per hour each unit not on maintenance is OUT iff
// rand() < FOR, load = base + randn()*sigma, unserved load drains the available energy-limited
// units (fully, or proportionally to capacity), an hour with deficit > 0 counts for the year and
// for the hour.
//
// B200 formulation: one thread per simulated year (the only sequential dependence is the ELU
// energy state inside a year); the 32 lanes of a warp walk the same hour together, so unit
// parameters, maintenance windows and the base load are warp-uniform broadcasts.  Random words:
// Philox4x32-10 keyed (seed; year, hour, block); unit u uses word u as an integer Bernoulli
// threshold test, words U and U+1 feed a fixed-sequence binary32 Box-Muller normal.  All load /
// capacity / energy arithmetic is FP64 in the reference's operation order (file compiled with
// --fmad=false), so injected uniform / normal matrices reproduce the reference loop bit for bit.
#include <math.h>

#include <vector>

#include "psra_internal.cuh"

#define DM_MAX_UNITS 32

struct DmArgs {
    int U, H;
    const double *cap; const double *for_rate; const uint32_t *for_thr;
    const int *mstart; const int *mweeks; const double *elim;
    const double *base_load; double lfu_std;
    uint32_t k0, k1; long long year0, nyears;
    const double *unif; const double *norm;     // injected (nullable)
    uint32_t *year_lole; uint32_t *hourly_fail;
};

// standard normal from two words: r = sqrt(2 max(E(x1), 0)), angle = (k + f) pi/2, fixed binary32 sequence
__device__ __forceinline__ float normal_u32x2(uint32_t x1, uint32_t x2)
{
    const float r = __fsqrt_rn(__fmul_rn(2.0f, fmaxf(neglog_u32(x1), 0.0f)));
    const uint32_t k = x2 >> 30;
    const float f = __fmul_rn((float)(2u * ((x2 >> 7) & 0x7FFFFFu) + 1u), 5.9604644775390625e-08f);
    const bool swap = f > 0.5f;
    const float g = swap ? __fadd_rn(1.0f, -f) : f;
    const float x = __fmul_rn(g, 1.57079637f);
    const float x2f = __fmul_rn(x, x);
    float s = __fmaf_rn(x2f, 2.75573192e-06f, -1.98412701e-04f);
    s = __fmaf_rn(x2f, s, 8.33333377e-03f);
    s = __fmaf_rn(x2f, s, -1.66666672e-01f);
    s = __fmaf_rn(__fmul_rn(x, x2f), s, x);
    float c = __fmaf_rn(x2f, 2.48015876e-05f, -1.38888892e-03f);
    c = __fmaf_rn(x2f, c, 4.16666679e-02f);
    c = __fmaf_rn(x2f, c, -0.5f);
    c = __fmaf_rn(x2f, c, 1.0f);
    const float sn = swap ? c : s, cs = swap ? s : c;
    const float v = (k == 0u) ? cs : (k == 1u) ? -sn : (k == 2u) ? -cs : sn;
    return __fmul_rn(r, v);
}

template <bool kInjected>
__global__ void __launch_bounds__(128) detailed_mc_kernel(const DmArgs a)
{
    __shared__ double s_cap[DM_MAX_UNITS], s_q[DM_MAX_UNITS], s_elim[DM_MAX_UNITS];
    __shared__ uint32_t s_thr[DM_MAX_UNITS];
    __shared__ int s_ms[DM_MAX_UNITS], s_me[DM_MAX_UNITS];
    if (threadIdx.x < a.U) {
        const int u = threadIdx.x;
        s_cap[u] = a.cap[u]; s_q[u] = a.for_rate[u]; s_elim[u] = a.elim[u]; s_thr[u] = a.for_thr[u];
        s_ms[u] = a.mstart[u]; s_me[u] = a.mstart[u] + a.mweeks[u];
    }
    __syncthreads();
    const long long y = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= a.nyears) return;
    const unsigned long long yy = (unsigned long long)(a.year0 + y);
    double energy[DM_MAX_UNITS];
#pragma unroll
    for (int i = 0; i < DM_MAX_UNITS; i++) energy[i] = 0.0;         // tail_risk.jl:27
    unsigned int count = 0;
    const int nblk = (a.U + 2 + 3) >> 2;
    for (int h = 0; h < a.H; h++) {
        const int week = h / 168 + 1;                               // div(h-1,168)+1, h 1-based (:31)
        double cap_unlimited = 0.0, cap_elu = 0.0;
        uint32_t elu_mask = 0;
        uint32_t x1 = 0, x2 = 0;
        auto unit_step = [&](int u, bool out) {
            if (week >= s_ms[u] && week < s_me[u]) return;          // maintenance (:39-42)
            if (out) return;                                        // random failure (:44)
            if (s_elim[u] < INFINITY) {                             // energy limit (:46-55)
                if (energy[u] >= s_elim[u]) return;
                cap_elu += s_cap[u];
                elu_mask |= 1u << u;
            } else {
                cap_unlimited += s_cap[u];
            }
        };
        double z;
        if constexpr (kInjected) {
            for (int u = 0; u < a.U; u++)
                unit_step(u, a.unif[((size_t)y * a.H + h) * a.U + u] < s_q[u]);
            z = a.norm[(size_t)y * a.H + h];
        } else {
            for (int b = 0; b < nblk; b++) {
                uint32_t w[4];
                philox4x32_10((uint32_t)yy, (uint32_t)(yy >> 32), (uint32_t)h, 0x444D0000u | (uint32_t)b, a.k0, a.k1, w);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int idx = 4 * b + q;
                    if (idx < a.U) unit_step(idx, w[q] < s_thr[idx]);
                    else if (idx == a.U) x1 = w[q];
                    else if (idx == a.U + 1) x2 = w[q];
                }
            }
            z = (double)normal_u32x2(x1, x2);
        }
        const double load = a.base_load[h] + z * a.lfu_std;         // :59
        double unserved = load - cap_unlimited;                     // :60
        if (unserved < 0.0) unserved = 0.0;
        double deficit = 0.0;
        if (unserved > 0) {
            if (unserved > cap_elu) {                               // :64-69
                deficit = unserved - cap_elu;
                for (uint32_t m = elu_mask; m; m &= m - 1) { const int u = __ffs(m) - 1; energy[u] += s_cap[u]; }
            } else {                                                // :70-76
                for (uint32_t m = elu_mask; m; m &= m - 1) {
                    const int u = __ffs(m) - 1;
                    const double share = unserved * (s_cap[u] / cap_elu);
                    energy[u] += share;
                }
            }
        }
        if (deficit > 0) {                                          // :79-84
            count++;
            if (a.hourly_fail) atomicAdd(&a.hourly_fail[h], 1u);
        }
    }
    a.year_lole[y] = count;
}

static int run_detailed(psra_handle *h, const psra_detailed_system *sys, const double *base_load, int H,
                        double lfu_std, long long year0, long long nyears, uint64_t seed, const double *unif,
                        const double *norm, uint32_t *year_lole, uint32_t *hourly_fail, float *kernel_ms)
{
    PSRA_REQUIRE(h, sys && base_load && year_lole, "null argument");
    PSRA_REQUIRE(h, sys->capacity_mw && sys->for_rate && sys->maint_start_week && sys->maint_weeks && sys->energy_limit_mwh,
                 "null system array");
    PSRA_REQUIRE(h, sys->n_units >= 1 && sys->n_units <= DM_MAX_UNITS - 2, "this kernel supports up to 30 units");
    PSRA_REQUIRE(h, H >= 1 && nyears >= 0 && year0 >= 0, "bad sizes");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    if (kernel_ms) *kernel_ms = 0.f;
    if (nyears == 0) return PSRA_OK;
    const int U = sys->n_units;
    const bool injected = unif != nullptr;
    std::vector<uint32_t> thr(U);
    for (int u = 0; u < U; u++) {
        PSRA_REQUIRE(h, sys->for_rate[u] >= 0.0 && sys->for_rate[u] <= 1.0, "FOR must be within [0, 1]");
        double t = floor(sys->for_rate[u] * 4294967296.0);
        thr[u] = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    }
    // device layout in scratch: cap | q | elim | base_load | (unif | norm) ; ints in scratch2: thr | ms | mw | year_lole | hourly
    size_t nd = 3 * (size_t)U + H;
    if (injected) nd += (size_t)nyears * H * U + (size_t)nyears * H;
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * nd);
    if (rc) return rc;
    rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, sizeof(uint32_t) * (3 * (size_t)U + (size_t)nyears + H));
    if (rc) return rc;
    double *d_cap = (double *)h->d_scratch, *d_q = d_cap + U, *d_el = d_q + U, *d_bl = d_el + U, *d_un = d_bl + H;
    double *d_no = d_un + (injected ? (size_t)nyears * H * U : 0);
    uint32_t *d_thr = (uint32_t *)h->d_scratch2; int *d_ms = (int *)(d_thr + U), *d_mw = d_ms + U;
    uint32_t *d_yl = (uint32_t *)(d_mw + U), *d_hf = d_yl + nyears;
    PSRA_CUDA(h, cudaMemcpyAsync(d_cap, sys->capacity_mw, sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_q, sys->for_rate, sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_el, sys->energy_limit_mwh, sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_bl, base_load, sizeof(double) * H, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_thr, thr.data(), sizeof(uint32_t) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_ms, sys->maint_start_week, sizeof(int) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_mw, sys->maint_weeks, sizeof(int) * U, cudaMemcpyHostToDevice, h->stream));
    if (injected) {
        PSRA_CUDA(h, cudaMemcpyAsync(d_un, unif, sizeof(double) * (size_t)nyears * H * U, cudaMemcpyHostToDevice, h->stream));
        PSRA_CUDA(h, cudaMemcpyAsync(d_no, norm, sizeof(double) * (size_t)nyears * H, cudaMemcpyHostToDevice, h->stream));
    }
    PSRA_CUDA(h, cudaMemsetAsync(d_hf, 0, sizeof(uint32_t) * (size_t)H, h->stream));
    DmArgs a{};
    a.U = U; a.H = H; a.cap = d_cap; a.for_rate = d_q; a.for_thr = d_thr; a.mstart = d_ms; a.mweeks = d_mw; a.elim = d_el;
    a.base_load = d_bl; a.lfu_std = lfu_std; a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.year0 = year0; a.nyears = nyears; a.unif = injected ? d_un : nullptr; a.norm = injected ? d_no : nullptr;
    a.year_lole = d_yl; a.hourly_fail = hourly_fail ? d_hf : nullptr;
    const unsigned grid = (unsigned)((nyears + 127) / 128);
    PSRA_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (injected) detailed_mc_kernel<true><<<grid, 128, 0, h->stream>>>(a);
    else detailed_mc_kernel<false><<<grid, 128, 0, h->stream>>>(a);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(year_lole, d_yl, sizeof(uint32_t) * (size_t)nyears, cudaMemcpyDeviceToHost, h->stream));
    if (hourly_fail) PSRA_CUDA(h, cudaMemcpyAsync(hourly_fail, d_hf, sizeof(uint32_t) * (size_t)H, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    PSRA_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (kernel_ms) *kernel_ms = ms;
    return PSRA_OK;
}

extern "C" int psra_detailed_mc(psra_handle *h, const psra_detailed_system *sys, const double *base_load_mw,
                                int32_t n_hours, double lfu_std_mw, int64_t year0, int64_t nyears, uint64_t seed,
                                uint32_t *year_lole, uint32_t *hourly_fail, float *kernel_ms)
{
    if (!h) return PSRA_E_INVALID;
    return run_detailed(h, sys, base_load_mw, n_hours, lfu_std_mw, year0, nyears, seed, nullptr, nullptr, year_lole,
                        hourly_fail, kernel_ms);
}

extern "C" int psra_detailed_eval_injected(psra_handle *h, const psra_detailed_system *sys, const double *base_load_mw,
                                           int32_t n_hours, double lfu_std_mw, int64_t nyears, const double *uniforms,
                                           const double *normals, uint32_t *year_lole, uint32_t *hourly_fail)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, uniforms && normals, "null injected matrices");
    return run_detailed(h, sys, base_load_mw, n_hours, lfu_std_mw, 0, nyears, 0, uniforms, normals, year_lole,
                        hourly_fail, nullptr);
}

// analytical.cu -- analytical cross-checks as FP64 CUDA kernels (sm_100a).
//
//  * psra_copt           : add_unit_convolution, GeneratingAdequacy/PowerSystemAdequacy.jl:67-111
//  * psra_copt_indices   : run_analytical's risk loop, PSA.jl:123-160
//  * psra_copt_indices_strict : calculate_indices, generating_adequacy_assessment.jl:113-146
//  * psra_fd_recursion   : add_unit_educational!, generating_adequacy_frequency.jl:76-129
//  * psra_markov2        : Markov_process.jl:89-110
//  * psra_dtmc_capacity  : Markov_process.jl:159-195
//  * psra_failure_times  : Markov_process.jl:39-60
//
// The file is compiled with --fmad=false and every product / sum is a single IEEE operation in
// the reference's order, so the COPT / F&D tables are bit-identical to the FP64 reference
// arithmetic; the index reductions use suffix sums and tree reductions (O(N + H) instead of the
// reference's O(H * tail) loop, PSA.jl:148-151) and agree to ~1e-14 relative.
#include <math.h>

#include <vector>

#include "psra_internal.cuh"

// --------------------------------------------------------------------------- COPT convolution
// PSA.jl:81-88: idx = Int(round(X/step)) + 1 (round-half-even), 0 outside the old table
__device__ __forceinline__ double copt_get(const double *p, int n, double x_val, double step)
{
    const double q = rint(x_val / step);
    if (q < 0.0 || q > (double)(n - 1)) return 0.0;
    return p[(int)q];
}

__global__ void copt_add_unit_kernel(const double *__restrict__ old_p, int n_old, double *__restrict__ new_p,
                                     int n_new, double step, double p, double q, double C, double C_lower,
                                     double C_upper, double q_lower, double q_upper, int exact)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_new) return;
    const double X = (double)i * step;
    if (exact) {   // PSA.jl:95-98
        new_p[i] = copt_get(old_p, n_old, X, step) * p + copt_get(old_p, n_old, X - C, step) * q;
    } else {       // PSA.jl:100-107, left-to-right sum of the three products
        new_p[i] = copt_get(old_p, n_old, X, step) * p + copt_get(old_p, n_old, X - C_lower, step) * q_lower +
                   copt_get(old_p, n_old, X - C_upper, step) * q_upper;
    }
}

extern "C" int psra_copt(psra_handle *h, const double *cap_mw, const double *for_rate, int32_t n_units,
                         double step, double *probs, int32_t max_len, int32_t *n_states)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, cap_mw && for_rate && probs && n_states, "null argument");
    PSRA_REQUIRE(h, n_units >= 0 && step > 0 && max_len >= 1, "bad COPT arguments");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    // table lengths first (PSA.jl:73-75: grid max of the OLD table + C, not the installed sum)
    std::vector<int> lens(n_units + 1);
    lens[0] = 1;
    for (int u = 0; u < n_units; u++) {
        PSRA_REQUIRE(h, cap_mw[u] >= 0, "negative capacity");
        const double max_old = (double)(lens[u] - 1) * step;
        const double len = ceil((max_old + cap_mw[u]) / step) + 1.0;
        if (len > (double)max_len)
            return psra_fail(h, PSRA_E_OVERFLOW, "COPT needs %.0f states, buffer holds %d", len, max_len);
        lens[u + 1] = (int)len;
    }
    const int n_final = lens[n_units];
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * (size_t)n_final);
    if (rc) return rc;
    rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, sizeof(double) * (size_t)n_final);
    if (rc) return rc;
    double *a = (double *)h->d_scratch, *b = (double *)h->d_scratch2;
    const double one = 1.0;
    PSRA_CUDA(h, cudaMemcpyAsync(a, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    for (int u = 0; u < n_units; u++) {
        const double C = cap_mw[u], q = for_rate[u], p = 1.0 - q;
        const int lower_idx = (int)floor(C / step);
        const double C_lower = lower_idx * step, C_upper = (lower_idx + 1) * step;
        const int exact = fabs(C - C_lower) < 1e-5;
        const double alpha = (C - C_lower) / step;
        const double q_upper = q * alpha, q_lower = q * (1.0 - alpha);
        const int n_new = lens[u + 1];
        copt_add_unit_kernel<<<(n_new + 255) / 256, 256, 0, h->stream>>>(a, lens[u], b, n_new, step, p, q, C,
                                                                         C_lower, C_upper, q_lower, q_upper, exact);
        PSRA_CUDA(h, cudaGetLastError());
        double *t = a; a = b; b = t;
    }
    PSRA_CUDA(h, cudaMemcpyAsync(probs, a, sizeof(double) * (size_t)n_final, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    *n_states = n_final;
    return PSRA_OK;
}

// ------------------------------------------------------------------------------ COPT indices
// suffix sums S_p[i] = sum_{k>=i} p_k and S_xp[i] = sum_{k>=i} (k*step) p_k, single block:
// each thread owns a contiguous run of states (from the top), then a block scan of the totals.
__global__ void __launch_bounds__(1024) copt_suffix_kernel(const double *__restrict__ p, int n, double step,
                                                          double *__restrict__ S_p, double *__restrict__ S_xp)
{
    __shared__ double sh_p[1024], sh_x[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int chunk = (n + T - 1) / T;
    // thread t owns reversed positions [t*chunk, (t+1)*chunk): state index k = n-1-pos
    double tp = 0.0, tx = 0.0;
    for (int j = 0; j < chunk; j++) {
        const int pos = t * chunk + j;
        if (pos < n) { const int k = n - 1 - pos; tp += p[k]; tx += ((double)k * step) * p[k]; }
    }
    sh_p[t] = tp; sh_x[t] = tx;
    __syncthreads();
    for (int d = 1; d < T; d <<= 1) {   // inclusive Hillis-Steele scan over threads
        double ap = 0.0, ax = 0.0;
        if (t >= d) { ap = sh_p[t - d]; ax = sh_x[t - d]; }
        __syncthreads();
        if (t >= d) { sh_p[t] += ap; sh_x[t] += ax; }
        __syncthreads();
    }
    double cp = sh_p[t] - tp, cx = sh_x[t] - tx;   // totals of the threads above
    for (int j = 0; j < chunk; j++) {
        const int pos = t * chunk + j;
        if (pos < n) {
            const int k = n - 1 - pos;
            cp += p[k]; cx += ((double)k * step) * p[k];
            S_p[k] = cp; S_xp[k] = cx;
        }
    }
    if (t == 0) { S_p[n] = 0.0; S_xp[n] = 0.0; }
}

// One thread per hour, deterministic block tree reduction, one partial per block.
template <bool kStrict>
__global__ void __launch_bounds__(256) copt_hours_kernel(const double *__restrict__ S_p, const double *__restrict__ S_xp,
                                                         int n, double step, double installed,
                                                         const double *__restrict__ load, int H,
                                                         double *__restrict__ part_lole, double *__restrict__ part_eue)
{
    __shared__ double sh_l[256], sh_e[256];
    const int hidx = blockIdx.x * blockDim.x + threadIdx.x;
    double l = 0.0, e = 0.0;
    if (hidx < H) {
        const double reserve = installed - load[hidx];
        if constexpr (!kStrict) {
            // PSA.jl:140: idx = floor(reserve/step) + 2 (1-based)
            const double idxd = floor(reserve / step) + 2.0;
            if (idxd <= (double)n && idxd >= 1.0) {
                const int k0 = (int)idxd - 1;
                l = S_p[k0];
                e = S_xp[k0] - reserve * S_p[k0];          // = sum_{k>=k0} (outage_k - reserve) p_k
            } else if (idxd < 1.0) {                       // PSA.jl:152-159
                l = 1.0;
                e = (load[hidx] - installed) + S_xp[0];
            }
        } else {
            // generating_adequacy_assessment.jl:131-139: states with outage > reserve (strict)
            double kd = floor(reserve / step);
            if (kd < -1.0) kd = -1.0;
            if (kd > (double)n) kd = (double)n;
            long long k = (long long)kd;
            while (k >= 0 && (double)k * step > reserve) k--;
            while (k + 1 < n && (double)(k + 1) * step <= reserve) k++;
            const int k0 = (int)(k + 1);                    // first state with outage > reserve
            if (k0 < n) {
                l = S_p[k0];
                e = S_xp[k0] - reserve * S_p[k0];
            }
        }
    }
    sh_l[threadIdx.x] = l; sh_e[threadIdx.x] = e;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) { sh_l[threadIdx.x] += sh_l[threadIdx.x + d]; sh_e[threadIdx.x] += sh_e[threadIdx.x + d]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_lole[blockIdx.x] = sh_l[0]; part_eue[blockIdx.x] = sh_e[0]; }
}

static int copt_indices_impl(psra_handle *h, bool strict, const double *probs, int n, double step,
                             double installed, const double *load, int H, double *lole, double *eue)
{
    PSRA_REQUIRE(h, probs && load && lole && eue, "null argument");
    PSRA_REQUIRE(h, n >= 1 && H >= 1 && step > 0, "bad sizes");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    const int nb = (H + 255) / 256;
    // scratch: probs[n] | S_p[n+1] | S_xp[n+1] ; scratch2: load[H] | partials[2*nb]
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * (3 * (size_t)n + 2));
    if (rc) return rc;
    rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, sizeof(double) * ((size_t)H + 2 * (size_t)nb));
    if (rc) return rc;
    double *d_p = (double *)h->d_scratch, *d_Sp = d_p + n, *d_Sx = d_Sp + n + 1;
    double *d_load = (double *)h->d_scratch2, *d_pl = d_load + H, *d_pe = d_pl + nb;
    PSRA_CUDA(h, cudaMemcpyAsync(d_p, probs, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_load, load, sizeof(double) * (size_t)H, cudaMemcpyHostToDevice, h->stream));
    copt_suffix_kernel<<<1, 1024, 0, h->stream>>>(d_p, n, step, d_Sp, d_Sx);
    PSRA_CUDA(h, cudaGetLastError());
    if (strict) {
        installed = (double)(n - 1) * step;   // generating_adequacy_assessment.jl:125
        copt_hours_kernel<true><<<nb, 256, 0, h->stream>>>(d_Sp, d_Sx, n, step, installed, d_load, H, d_pl, d_pe);
    } else {
        copt_hours_kernel<false><<<nb, 256, 0, h->stream>>>(d_Sp, d_Sx, n, step, installed, d_load, H, d_pl, d_pe);
    }
    PSRA_CUDA(h, cudaGetLastError());
    std::vector<double> part(2 * (size_t)nb);
    PSRA_CUDA(h, cudaMemcpyAsync(part.data(), d_pl, sizeof(double) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    double l = 0.0, e = 0.0;
    for (int b = 0; b < nb; b++) { l += part[b]; e += part[nb + b]; }   // fixed order: a few dozen partials
    *lole = l; *eue = e;
    return PSRA_OK;
}

extern "C" int psra_copt_indices(psra_handle *h, const double *probs, int32_t n_states, double step,
                                 double total_installed, const double *load_mw, int32_t n_hours,
                                 double *lole, double *eue)
{
    if (!h) return PSRA_E_INVALID;
    return copt_indices_impl(h, false, probs, n_states, step, total_installed, load_mw, n_hours, lole, eue);
}

extern "C" int psra_copt_indices_strict(psra_handle *h, const double *probs, int32_t n_states, double step,
                                        const double *ldc_mw, int32_t n_hours, double *lole, double *eue)
{
    if (!h) return PSRA_E_INVALID;
    return copt_indices_impl(h, true, probs, n_states, step, 0.0, ldc_mw, n_hours, lole, eue);
}

// ---------------------------------------------------------- frequency & duration recursion
// generating_adequacy_frequency.jl:76-98: P(x<0)=1, F=0; first level >= x; (0,0) beyond the table
__device__ __forceinline__ void fd_get(const double *P, const double *F, int n, double x, double &p, double &f)
{
    if (x < 0) { p = 1.0; f = 0.0; return; }
    const double idx = ceil(x);
    if (idx >= (double)n) { p = 0.0; f = 0.0; return; }
    p = P[(int)idx]; f = F[(int)idx];
}

__global__ void fd_add_unit_kernel(const double *__restrict__ P_old, const double *__restrict__ F_old, int n_old,
                                   double *__restrict__ P_new, double *__restrict__ F_new, int n_new, double C,
                                   double p, double q, double lam)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_new) return;
    const double X = (double)i;
    double PX, FX, PXC, FXC;
    fd_get(P_old, F_old, n_old, X, PX, FX);
    fd_get(P_old, F_old, n_old, X - C, PXC, FXC);
    P_new[i] = (p * PX) + (q * PXC);                         // :110
    const double term1 = p * FX, term2 = q * FXC;            // :113-116
    const double term3 = lam * p * (PXC - PX);
    F_new[i] = term1 + term2 + term3;
}

extern "C" int psra_fd_recursion(psra_handle *h, const double *cap_mw, const double *mtbf_h,
                                 const double *mttr_h, int32_t n_units, double *cum_prob, double *cum_freq,
                                 int32_t max_len, int32_t *n_states)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, cap_mw && mtbf_h && mttr_h && cum_prob && cum_freq && n_states, "null argument");
    PSRA_REQUIRE(h, n_units >= 0 && max_len >= 1, "bad arguments");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    std::vector<int> lens(n_units + 1);
    lens[0] = 1;
    for (int u = 0; u < n_units; u++) {
        // collect(0.0:1.0:current_max + capacity), generating_adequacy_frequency.jl:62-70
        const double len = floor((double)(lens[u] - 1) + cap_mw[u]) + 1.0;
        if (len > (double)max_len)
            return psra_fail(h, PSRA_E_OVERFLOW, "F&D table needs %.0f states, buffer holds %d", len, max_len);
        lens[u + 1] = (int)len;
    }
    const int n_final = lens[n_units];
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * 2 * (size_t)n_final);
    if (rc) return rc;
    rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, sizeof(double) * 2 * (size_t)n_final);
    if (rc) return rc;
    double *Pa = (double *)h->d_scratch, *Fa = Pa + n_final;
    double *Pb = (double *)h->d_scratch2, *Fb = Pb + n_final;
    const double one = 1.0, zero = 0.0;
    PSRA_CUDA(h, cudaMemcpyAsync(Pa, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(Fa, &zero, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    for (int u = 0; u < n_units; u++) {
        const double lam = 8760.0 / mtbf_h[u], mu = 8760.0 / mttr_h[u];   // :26-27
        const double q = lam / (lam + mu), p = 1.0 - q;                   // :30-31
        const int n_new = lens[u + 1];
        fd_add_unit_kernel<<<(n_new + 255) / 256, 256, 0, h->stream>>>(Pa, Fa, lens[u], Pb, Fb, n_new, cap_mw[u], p, q, lam);
        PSRA_CUDA(h, cudaGetLastError());
        double *t = Pa; Pa = Pb; Pb = t;
        t = Fa; Fa = Fb; Fb = t;
    }
    PSRA_CUDA(h, cudaMemcpyAsync(cum_prob, Pa, sizeof(double) * (size_t)n_final, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(cum_freq, Fa, sizeof(double) * (size_t)n_final, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    *n_states = n_final;
    return PSRA_OK;
}

// ----------------------------------------------------------------------------- Markov demos
// Markov_process.jl:89-110: P = [[1-p01, p01], [p10, 1-p10]], pi(t+1) = pi(t) P from [1, 0]
__global__ void markov2_kernel(double lambda, double mu, double dt, int steps, double *prob_down)
{
    if (blockIdx.x || threadIdx.x) return;
    const double p01 = 1 - exp(-lambda * dt), p10 = 1 - exp(-mu * dt);
    const double p00 = 1 - p01, p11 = 1 - p10;
    double up = 1.0, down = 0.0;
    for (int t = 0; t < steps; t++) {
        const double nu = up * p00 + down * p10;
        const double nd = up * p01 + down * p11;
        up = nu; down = nd;
        prob_down[t] = down;
    }
}

extern "C" int psra_markov2(psra_handle *h, double lambda, double mu, double dt, int32_t steps, double *prob_down)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, prob_down && steps >= 1, "bad arguments");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * (size_t)steps);
    if (rc) return rc;
    markov2_kernel<<<1, 32, 0, h->stream>>>(lambda, mu, dt, steps, (double *)h->d_scratch);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaMemcpyAsync(prob_down, h->d_scratch, sizeof(double) * (size_t)steps, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    return PSRA_OK;
}

// Markov_process.jl:159-195.  Thread = unit (strided): the unit's chain over time is sequential,
// units are independent; pass 1 writes the UP capacity contribution of each (t, unit), pass 2
// (same kernel, after a block barrier per tile of hours) sums over units in unit order.
__global__ void __launch_bounds__(256) dtmc_kernel(int U, const double *__restrict__ mttf, const double *__restrict__ mttr,
                                                   const double *__restrict__ cap, int T, const double *__restrict__ r,
                                                   unsigned char *__restrict__ up, double *__restrict__ avail)
{
    for (int i = threadIdx.x; i < U; i += blockDim.x) {
        const double p01 = 1 - exp(-(1 / mttf[i])), p10 = 1 - exp(-(1 / mttr[i]));   // :166-167
        int state = 0;
        for (int t = 0; t < T; t++) {
            const double x = r[(size_t)t * U + i];
            if (state == 0) { if (x < p01) state = 1; }       // :176-179
            else            { if (x < p10) state = 0; }       // :180-183
            up[(size_t)t * U + i] = (unsigned char)(state == 0);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double c = 0.0;
        for (int i = 0; i < U; i++) if (up[(size_t)t * U + i]) c += cap[i];   // :188-193 unit order
        avail[t] = c;
    }
}

extern "C" int psra_dtmc_capacity(psra_handle *h, const double *mttf_h, const double *mttr_h, const double *cap_mw,
                                  int32_t n_units, const double *r, int32_t n_steps, double *avail_mw)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, mttf_h && mttr_h && cap_mw && r && avail_mw, "null argument");
    PSRA_REQUIRE(h, n_units >= 1 && n_steps >= 1, "bad sizes");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    const size_t U = n_units, T = n_steps;
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * (3 * U + T * U + T));
    if (rc) return rc;
    rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, T * U);
    if (rc) return rc;
    double *d_mttf = (double *)h->d_scratch, *d_mttr = d_mttf + U, *d_cap = d_mttr + U, *d_r = d_cap + U, *d_av = d_r + T * U;
    PSRA_CUDA(h, cudaMemcpyAsync(d_mttf, mttf_h, sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_mttr, mttr_h, sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_cap, cap_mw, sizeof(double) * U, cudaMemcpyHostToDevice, h->stream));
    PSRA_CUDA(h, cudaMemcpyAsync(d_r, r, sizeof(double) * T * U, cudaMemcpyHostToDevice, h->stream));
    dtmc_kernel<<<1, 256, 0, h->stream>>>(n_units, d_mttf, d_mttr, d_cap, n_steps, d_r, (unsigned char *)h->d_scratch2, d_av);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaMemcpyAsync(avail_mw, d_av, sizeof(double) * T, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    return PSRA_OK;
}

// Markov_process.jl:39-60 "why constant rate = exponential distribution": thread = component; every dt hours the
// component fails with probability lambda * dt (rand() < lambda * dt, :54); t accumulates by repeated addition
// (:58) and the component is dropped once t > max_time (:59).  rand() is either the injected matrix r[i*K + k]
// or draw k of the Philox stream keyed (seed; component, 0x46540000 | block): u = word / 2^32 (exact in FP64).
__global__ void __launch_bounds__(256) failure_time_kernel(double thr, double dt, double max_time, long long n, uint32_t k0,
                                                           uint32_t k1, const double *__restrict__ r, int K, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = 0.0, res = -1.0;
    uint32_t x[4];
    for (int k = 0;; k++) {
        double u;
        if (r) {
            if (k >= K) { res = -2.0; break; }               // injected uniforms exhausted (reported as an error)
            u = r[(size_t)i * K + k];
        } else {
            if ((k & 3) == 0) philox4x32_10((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), 0x46540000u, (uint32_t)(k >> 2), k0, k1, x);
            u = (double)x[k & 3] * 2.3283064365386963e-10;    // / 2^32
        }
        if (u < thr) { res = t; break; }
        t += dt;
        if (t > max_time) break;
    }
    out[i] = res;
}

extern "C" int psra_failure_times(psra_handle *h, double lambda, double dt, double max_time, int64_t n, uint64_t seed,
                                  const double *r, int32_t K, double *failure_time)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, failure_time && n >= 1, "bad output / component count");
    PSRA_REQUIRE(h, lambda > 0 && dt > 0 && max_time >= 0 && max_time / dt < 1e8, "bad rate / time step / horizon");
    PSRA_REQUIRE(h, !r || K >= 1, "injected uniforms need K >= 1");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    const size_t nr = r ? (size_t)n * (size_t)K : 0;
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(double) * ((size_t)n + nr));
    if (rc) return rc;
    double *d_out = (double *)h->d_scratch, *d_r = nullptr;
    if (r) {
        d_r = d_out + n;
        PSRA_CUDA(h, cudaMemcpyAsync(d_r, r, sizeof(double) * nr, cudaMemcpyHostToDevice, h->stream));
    }
    failure_time_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(lambda * dt, dt, max_time, n, (uint32_t)seed,
                                                                           (uint32_t)(seed >> 32), d_r, K, d_out);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaMemcpyAsync(failure_time, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < n; i++)
        if (failure_time[i] == -2.0) return psra_fail(h, PSRA_E_OVERFLOW, "injected uniforms exhausted: component %lld needed more than K=%d draws", (long long)i, K);
    return PSRA_OK;
}

// nonseq_mc.cu -- non-sequential state-sampling Monte Carlo on sm_100a.
//
// Replaces the body of run_non_sequential_mc (GeneratingAdequacy/PowerSystemAdequacy.jl:169-208):
// per sample, unit u is UP iff rand() >= FOR_u (PSA.jl:183; MATLAB twin DOWN iff rand < U,
// Montecarlo_nsq_single/mc_sampling.m:35), available capacity = sum of UP capacities
// (PSA.jl:184), and that ONE capacity is evaluated against all H hours: hours with
// cap < load counted, deficits summed (PSA.jl:191-197).
//
// B200 formulation: one thread per sample.  Unit states come from Philox4x32-10 words keyed
// (seed; sample, unit/4) compared against integer thresholds floor(FOR * 2^32) and are kept
// bit-packed (bit u%32 of word u/32 = UP); the capacity is the masked sum of the unit
// capacities staged in shared memory.  The reference's O(H) hour loop per sample collapses to
// one binary search in the ascending-sorted load curve (staged in shared memory) plus one
// suffix-sum lookup:  LOL hours = #{h : load_h > cap},  ENS = sum_{load_h > cap} load_h
// - cap * LOL hours -- integer arithmetic, hence identical to the hour loop bit for bit.
//
// Systems of <= 32 units whose installed capacity (fixed point) stays below 2^16 take nonseq_fast_kernel:
// the 32 Bernoulli tests are unrolled (thresholds read four at a time from shared memory), the capacity of
// the packed state word comes from four 256-entry byte tables (sum of the capacities of every byte pattern)
// and the LOL hours from a table indexed by the integer capacity -- 4 + 1 shared-memory lookups instead of
// 32 predicated adds and a 14-step binary search.  Same Philox words, same integers.
#include <algorithm>
#include <vector>

#include "psra_internal.cuh"

struct NsArgs {
    int U, H, W, group;
    const int32_t *cap; const uint32_t *for_thr; const double *for_rate;
    const int32_t *sorted;        // [H] ascending
    const uint16_t *lol_tab;      // [total_cap+1] (fast path)
    const int32_t *byte_tab;      // [4][256]      (fast path)
    int total_cap;
    const long long *suffix;      // [H+1] suffix[i] = sum_{k>=i} sorted[k]
    uint32_t k0, k1;
    long long i0, n;
    const uint32_t *in_states;    // injected packed states or nullptr
    const double *in_uniforms;    // injected uniforms or nullptr
    uint32_t *lol; long long *ens; int32_t *cap_out; uint32_t *states;
    unsigned long long *group_lol; unsigned long long *acc;
};

enum { NS_PHILOX = 0, NS_STATES = 1, NS_UNIFORMS = 2 };

// U <= 32, total_cap < 65536, H < 65536.  kBlk = Philox blocks per sample, ceil(U / 4), a compile-time constant: the blocks of a
// sample are independent, and only in ONE basic block does ptxas interleave their 2 kBlk multiply chains (with a run-time
// count every block was its own basic block: two chains in flight per warp, issue slots 57 % busy)
#ifndef NSF_BPS
#define NSF_BPS 2
#endif
template <int kBlk>
__global__ void __launch_bounds__(256, NSF_BPS) nonseq_fast_kernel(const NsArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *s_btab = reinterpret_cast<int32_t *>(smem_raw);                      // [4][256]
    uint32_t *s_thr = reinterpret_cast<uint32_t *>(s_btab + 1024);                // [32]
    uint16_t *s_lol = reinterpret_cast<uint16_t *>(s_thr + 32);                   // [total_cap+1]
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_btab[i] = a.byte_tab[i];
    if (threadIdx.x < 32) s_thr[threadIdx.x] = threadIdx.x < a.U ? a.for_thr[threadIdx.x] : 0u;
    for (int i = threadIdx.x; i <= a.total_cap; i += blockDim.x) s_lol[i] = a.lol_tab[i];
    __syncthreads();
    const uint32_t umask = a.U >= 32 ? 0xffffffffu : ((1u << a.U) - 1u);

    unsigned long long acc_lol = 0, acc_swl = 0, acc_lol2 = 0, acc_e2lo = 0, acc_e2hi = 0;
    long long acc_ens = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const unsigned long long s = (unsigned long long)(a.i0 + i);
        uint32_t word = 0;
#pragma unroll
        for (int g = 0; g < kBlk; g++) {
            uint32_t x[4];
            philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)g, 0x4E53u, a.k0, a.k1, x);
            const uint4 t = *reinterpret_cast<const uint4 *>(s_thr + 4 * g);
            word |= (x[0] >= t.x ? 1u : 0u) << (4 * g);             // PSA.jl:183 in integer form
            word |= (x[1] >= t.y ? 1u : 0u) << (4 * g + 1);
            word |= (x[2] >= t.z ? 1u : 0u) << (4 * g + 2);
            word |= (x[3] >= t.w ? 1u : 0u) << (4 * g + 3);
        }
        word &= umask;
        const int cap = s_btab[word & 255u] + s_btab[256 + ((word >> 8) & 255u)] + s_btab[512 + ((word >> 16) & 255u)] +
                        s_btab[768 + (word >> 24)];
        const unsigned int lolh = s_lol[cap];
        long long ens = 0;
        if (lolh) ens = __ldg(&a.suffix[a.H - (int)lolh]) - (long long)cap * (long long)lolh;
        if (a.states) a.states[i] = word;
        if (a.lol) a.lol[i] = lolh;
        if (a.ens) a.ens[i] = ens;
        if (a.cap_out) a.cap_out[i] = cap;
        if (a.group_lol && lolh) atomicAdd(&a.group_lol[i / a.group], (unsigned long long)lolh);
        acc_lol += lolh; acc_ens += ens; acc_swl += lolh ? 1 : 0;
        acc_lol2 += (unsigned long long)lolh * lolh;
        const unsigned long long e = (unsigned long long)ens;
        const unsigned long long plo = e * e, phi = __umul64hi(e, e);
        const unsigned long long nlo = acc_e2lo + plo;
        acc_e2hi += phi + (nlo < acc_e2lo ? 1ull : 0ull);
        acc_e2lo = nlo;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc_lol += __shfl_xor_sync(0xffffffffu, acc_lol, d);
        acc_ens += __shfl_xor_sync(0xffffffffu, acc_ens, d);
        acc_swl += __shfl_xor_sync(0xffffffffu, acc_swl, d);
        acc_lol2 += __shfl_xor_sync(0xffffffffu, acc_lol2, d);
    }
    warp_sum_u128(acc_e2lo, acc_e2hi);       // one 128-bit atomic per warp, not per thread (they all hit one address)
    if ((threadIdx.x & 31) == 0) {
        if (acc_e2lo | acc_e2hi) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], acc_e2lo, acc_e2hi);
        if (acc_lol) atomicAdd(&a.acc[ACC_LOL], acc_lol);
        if (acc_ens) atomicAdd(&a.acc[ACC_ENS], (unsigned long long)acc_ens);
        if (acc_swl) atomicAdd(&a.acc[ACC_YWL], acc_swl);
        if (acc_lol2) atomicAdd(&a.acc[ACC_LOL2], acc_lol2);
    }
}

template <int kMode>
__global__ void __launch_bounds__(256) nonseq_kernel(const NsArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *s_sorted = reinterpret_cast<int32_t *>(smem_raw);
    int32_t *s_cap = s_sorted + a.H;
    uint32_t *s_thr = reinterpret_cast<uint32_t *>(s_cap + a.U);
    for (int i = threadIdx.x; i < a.H; i += blockDim.x) s_sorted[i] = a.sorted[i];
    for (int i = threadIdx.x; i < a.U; i += blockDim.x) { s_cap[i] = a.cap[i]; s_thr[i] = a.for_thr[i]; }
    __syncthreads();

    unsigned long long acc_lol = 0, acc_swl = 0, acc_lol2 = 0, acc_e2lo = 0, acc_e2hi = 0;
    long long acc_ens = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        int cap = 0;
        if constexpr (kMode == NS_PHILOX) {
            const unsigned long long s = (unsigned long long)(a.i0 + i);
            uint32_t word = 0;
            for (int u0 = 0; u0 < a.U; u0 += 4) {
                uint32_t x[4];
                philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(u0 >> 2), 0x4E53u, a.k0, a.k1, x);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int u = u0 + q;
                    if (u < a.U && x[q] >= s_thr[u]) {      // PSA.jl:183 in integer form
                        cap += s_cap[u];
                        word |= 1u << (u & 31);
                    }
                }
                if (((u0 + 4) & 31) == 0 || u0 + 4 >= a.U) {
                    if (a.states) a.states[(size_t)i * a.W + (u0 >> 5)] = word;
                    word = 0;
                }
            }
        } else if constexpr (kMode == NS_STATES) {
            for (int w = 0; w < a.W; w++) {
                uint32_t word = a.in_states[(size_t)i * a.W + w];
                if (w == a.W - 1 && (a.U & 31)) word &= (1u << (a.U & 31)) - 1u;
                if (a.states) a.states[(size_t)i * a.W + w] = word;
                for (; word; word &= word - 1) cap += s_cap[w * 32 + __ffs(word) - 1];
            }
        } else {
            uint32_t word = 0;
            for (int u = 0; u < a.U; u++) {
                if (a.in_uniforms[(size_t)i * a.U + u] >= a.for_rate[u]) {   // PSA.jl:183
                    cap += s_cap[u];
                    word |= 1u << (u & 31);
                }
                if (((u + 1) & 31) == 0 || u + 1 == a.U) {
                    if (a.states) a.states[(size_t)i * a.W + (u >> 5)] = word;
                    word = 0;
                }
            }
        }
        // ub = #{load <= cap}; loss hours are the H - ub largest loads
        int lo = 0, hi = a.H;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_sorted[mid] <= cap) lo = mid + 1; else hi = mid;
        }
        const unsigned int lolh = (unsigned int)(a.H - lo);
        long long ens = 0;
        if (lolh) ens = __ldg(&a.suffix[lo]) - (long long)cap * (long long)lolh;
        if (a.lol) a.lol[i] = lolh;
        if (a.ens) a.ens[i] = ens;
        if (a.cap_out) a.cap_out[i] = cap;
        if (a.group_lol && lolh) atomicAdd(&a.group_lol[i / a.group], (unsigned long long)lolh);
        acc_lol += lolh; acc_ens += ens; acc_swl += lolh ? 1 : 0;
        acc_lol2 += (unsigned long long)lolh * lolh;
        const unsigned long long e = (unsigned long long)ens;
        const unsigned long long plo = e * e, phi = __umul64hi(e, e);
        const unsigned long long nlo = acc_e2lo + plo;
        acc_e2hi += phi + (nlo < acc_e2lo ? 1ull : 0ull);
        acc_e2lo = nlo;
    }
    // warp reduction of the integer accumulators, one atomic set per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc_lol += __shfl_xor_sync(0xffffffffu, acc_lol, d);
        acc_ens += __shfl_xor_sync(0xffffffffu, acc_ens, d);
        acc_swl += __shfl_xor_sync(0xffffffffu, acc_swl, d);
        acc_lol2 += __shfl_xor_sync(0xffffffffu, acc_lol2, d);
    }
    warp_sum_u128(acc_e2lo, acc_e2hi);       // one 128-bit atomic per warp, not per thread (they all hit one address)
    if ((threadIdx.x & 31) == 0) {
        if (acc_e2lo | acc_e2hi) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], acc_e2lo, acc_e2hi);
        if (acc_lol) atomicAdd(&a.acc[ACC_LOL], acc_lol);
        if (acc_ens) atomicAdd(&a.acc[ACC_ENS], (unsigned long long)acc_ens);
        if (acc_swl) atomicAdd(&a.acc[ACC_YWL], acc_swl);
        if (acc_lol2) atomicAdd(&a.acc[ACC_LOL2], acc_lol2);
    }
}

// sorted load + suffix sums (host-side preparation at first use after psra_set_load)
static int prepare_sorted(psra_handle *h)
{
    if (h->tab_valid) return PSRA_OK;
    std::vector<int32_t> ld((size_t)h->Wd * 32);
    PSRA_CUDA(h, cudaMemcpy(ld.data(), h->d_load, sizeof(int32_t) * ld.size(), cudaMemcpyDeviceToHost));
    ld.resize(h->H);
    std::sort(ld.begin(), ld.end());
    std::vector<long long> suf((size_t)h->H + 1, 0);
    for (int i = h->H - 1; i >= 0; i--) suf[i] = suf[i + 1] + ld[i];
    if (h->d_load_sorted) cudaFree(h->d_load_sorted);
    if (h->d_load_suffix) cudaFree(h->d_load_suffix);
    h->d_load_sorted = nullptr; h->d_load_suffix = nullptr;
    PSRA_CUDA(h, cudaMalloc(&h->d_load_sorted, sizeof(int32_t) * (size_t)h->H));
    PSRA_CUDA(h, cudaMalloc(&h->d_load_suffix, sizeof(long long) * ((size_t)h->H + 1)));
    PSRA_CUDA(h, cudaMemcpy(h->d_load_sorted, ld.data(), sizeof(int32_t) * (size_t)h->H, cudaMemcpyHostToDevice));
    PSRA_CUDA(h, cudaMemcpy(h->d_load_suffix, suf.data(), sizeof(long long) * suf.size(), cudaMemcpyHostToDevice));
    // tables of the fast path (<= 32 units): LOL hours of every integer capacity, capacity of every state byte
    if (h->d_lol_tab) cudaFree(h->d_lol_tab);
    if (h->d_byte_tab) cudaFree(h->d_byte_tab);
    h->d_lol_tab = nullptr; h->d_byte_tab = nullptr;
    if (h->U <= 32 && h->total_cap < 65536 && h->H < 65536) {
        std::vector<uint16_t> lt((size_t)h->total_cap + 1);
        size_t ub = 0;                                         // #{load <= c}
        for (long long c = 0; c <= h->total_cap; c++) {
            while (ub < ld.size() && ld[ub] <= c) ub++;
            lt[(size_t)c] = (uint16_t)(ld.size() - ub);
        }
        std::vector<int32_t> capv(h->U), bt(1024, 0);
        PSRA_CUDA(h, cudaMemcpy(capv.data(), h->d_cap, sizeof(int32_t) * (size_t)h->U, cudaMemcpyDeviceToHost));
        for (int b = 0; b < 4; b++)
            for (int v = 0; v < 256; v++)
                for (int k = 0; k < 8; k++)
                    if (((v >> k) & 1) && 8 * b + k < h->U) bt[256 * b + v] += capv[8 * b + k];
        PSRA_CUDA(h, cudaMalloc(&h->d_lol_tab, sizeof(uint16_t) * lt.size()));
        PSRA_CUDA(h, cudaMalloc(&h->d_byte_tab, sizeof(int32_t) * bt.size()));
        PSRA_CUDA(h, cudaMemcpy(h->d_lol_tab, lt.data(), sizeof(uint16_t) * lt.size(), cudaMemcpyHostToDevice));
        PSRA_CUDA(h, cudaMemcpy(h->d_byte_tab, bt.data(), sizeof(int32_t) * bt.size(), cudaMemcpyHostToDevice));
    }
    h->tab_valid = true;
    return PSRA_OK;
}

static int run_nonseq(psra_handle *h, int mode, const void *input, long long i0, long long n, uint64_t seed,
                      const psra_nonseq_outputs *out, psra_nonseq_summary *summary)
{
    PSRA_REQUIRE(h, h->U > 0, "psra_set_system has not been called");
    PSRA_REQUIRE(h, h->H > 0, "psra_set_load has not been called");
    PSRA_REQUIRE(h, summary != nullptr, "summary must not be NULL");
    PSRA_REQUIRE(h, n >= 0 && i0 >= 0, "negative sample range");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    memset(summary, 0, sizeof(*summary));
    summary->samples = n;
    if (n == 0) return PSRA_OK;
    int rc = prepare_sorted(h);
    if (rc) return rc;

    NsArgs a{};
    a.U = h->U; a.H = h->H; a.W = (h->U + 31) / 32;
    a.cap = h->d_cap; a.for_thr = h->d_for_thr; a.for_rate = h->d_for;
    a.sorted = h->d_load_sorted; a.suffix = (const long long *)h->d_load_suffix;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.i0 = i0; a.n = n; a.acc = h->d_acc; a.group = 1;

    a.lol_tab = h->d_lol_tab; a.byte_tab = h->d_byte_tab; a.total_cap = (int)h->total_cap;
    const size_t smem_fast = sizeof(int32_t) * 1024 + sizeof(uint32_t) * 32 + ((sizeof(uint16_t) * ((size_t)h->total_cap + 1) + 15) & ~(size_t)15);
    const bool fastp = mode == NS_PHILOX && h->d_lol_tab && smem_fast <= 100 * 1024 && !h->cfg.reserved[0];
    const size_t smem = fastp ? smem_fast : sizeof(int32_t) * ((size_t)h->H + h->U) + sizeof(uint32_t) * (size_t)h->U;
    PSRA_REQUIRE(h, smem <= h->smem_optin, "load curve + unit table exceed shared memory");

    if (mode == NS_STATES) {
        const size_t bytes = sizeof(uint32_t) * (size_t)n * a.W;
        rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, bytes);
        if (rc) return rc;
        PSRA_CUDA(h, cudaMemcpyAsync(h->d_scratch, input, bytes, cudaMemcpyHostToDevice, h->stream));
        a.in_states = (const uint32_t *)h->d_scratch;
    } else if (mode == NS_UNIFORMS) {
        const size_t bytes = sizeof(double) * (size_t)n * a.U;
        rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, bytes);
        if (rc) return rc;
        PSRA_CUDA(h, cudaMemcpyAsync(h->d_scratch, input, bytes, cudaMemcpyHostToDevice, h->stream));
        a.in_uniforms = (const double *)h->d_scratch;
    }
    const bool want_vec = out && (out->lol_hours || out->ens_fp);
    if (want_vec) {
        rc = psra_reserve_outputs(h, n);
        if (rc) return rc;
        a.lol = h->d_lol; a.ens = (long long *)h->d_ens;
    }
    h->kept_n = 0;
    size_t cap_bytes = 0, st_bytes = 0;
    if (out && (out->cap_avail || out->states)) {
        cap_bytes = out->cap_avail ? sizeof(int32_t) * (size_t)n : 0;
        st_bytes = out->states ? sizeof(uint32_t) * (size_t)n * a.W : 0;
        rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, cap_bytes + st_bytes);
        if (rc) return rc;
        if (out->cap_avail) a.cap_out = (int32_t *)h->d_scratch2;
        if (out->states) a.states = (uint32_t *)((unsigned char *)h->d_scratch2 + cap_bytes);
    }
    long long ngroups = 0;
    if (out && (out->group_lol || out->history)) {
        PSRA_REQUIRE(h, out->group >= 1, "group must be >= 1");
        a.group = out->group;
        ngroups = (n + a.group - 1) / a.group;
        if (h->group_cap < ngroups) {
            if (h->d_group) cudaFree(h->d_group);
            h->d_group = nullptr; h->group_cap = 0;
            PSRA_CUDA(h, cudaMalloc(&h->d_group, sizeof(long long) * (size_t)ngroups));
            h->group_cap = ngroups;
        }
        PSRA_CUDA(h, cudaMemsetAsync(h->d_group, 0, sizeof(long long) * (size_t)ngroups, h->stream));
        a.group_lol = (unsigned long long *)h->d_group;
    }
    PSRA_CUDA(h, cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * ACC_COUNT, h->stream));

    static void (*const fast_kernels[8])(NsArgs) = {nonseq_fast_kernel<1>, nonseq_fast_kernel<2>, nonseq_fast_kernel<3>, nonseq_fast_kernel<4>,
                                                     nonseq_fast_kernel<5>, nonseq_fast_kernel<6>, nonseq_fast_kernel<7>, nonseq_fast_kernel<8>};
    void (*kern)(NsArgs) = fastp ? fast_kernels[((h->U + 3) >> 2) - 1] : mode == NS_PHILOX ? nonseq_kernel<NS_PHILOX>
                         : mode == NS_STATES ? nonseq_kernel<NS_STATES> : nonseq_kernel<NS_UNIFORMS>;
    PSRA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    PSRA_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, smem));
    if (bps < 1) return psra_fail(h, PSRA_E_CUDA, "non-sequential kernel does not fit on an SM");
    long long grid = (long long)h->sm_count * bps;
    const long long need = (n + 255) / 256;
    if (grid > need) grid = need;

    PSRA_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    kern<<<(unsigned)grid, 256, smem, h->stream>>>(a);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaEventRecord(h->ev1, h->stream));

    unsigned long long acc[ACC_COUNT];
    PSRA_CUDA(h, cudaMemcpyAsync(acc, h->d_acc, sizeof(acc), cudaMemcpyDeviceToHost, h->stream));
    if (out) {
        if (out->lol_hours) PSRA_CUDA(h, cudaMemcpyAsync(out->lol_hours, h->d_lol, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        if (out->ens_fp)    PSRA_CUDA(h, cudaMemcpyAsync(out->ens_fp, h->d_ens, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        if (out->cap_avail) PSRA_CUDA(h, cudaMemcpyAsync(out->cap_avail, a.cap_out, cap_bytes, cudaMemcpyDeviceToHost, h->stream));
        if (out->states)    PSRA_CUDA(h, cudaMemcpyAsync(out->states, a.states, st_bytes, cudaMemcpyDeviceToHost, h->stream));
        if (out->group_lol) PSRA_CUDA(h, cudaMemcpyAsync(out->group_lol, h->d_group, sizeof(long long) * (size_t)ngroups, cudaMemcpyDeviceToHost, h->stream));
        if (out->history && !h->multi_defer) {
            rc = psra_history_to_host(h, h->d_group, n / a.group, a.group, out->history);
            if (rc) return rc;
        }
    }
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    PSRA_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    summary->kernel_ms = ms;
    summary->sum_lol_hours = (int64_t)acc[ACC_LOL];
    summary->sum_ens_fp = (int64_t)acc[ACC_ENS];
    summary->samples_with_loss = (int64_t)acc[ACC_YWL];
    summary->sum_lol_sq = acc[ACC_LOL2];
    summary->sum_ens_sq_lo = acc[ACC_ENS2_LO];
    summary->sum_ens_sq_hi = acc[ACC_ENS2_HI];
    return PSRA_OK;
}

int psra_run_nonseq_range(psra_handle *h, long long i0, long long n, uint64_t seed, const psra_nonseq_outputs *out,
                          psra_nonseq_summary *summary)
{
    return run_nonseq(h, NS_PHILOX, nullptr, i0, n, seed, out, summary);
}

extern "C" int psra_nonseq_mc(psra_handle *h, int64_t sample0, int64_t n, uint64_t seed,
                              const psra_nonseq_outputs *out, psra_nonseq_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    if (!h->peers.empty()) return psra_multi_nonseq_mc(h, sample0, n, seed, out, summary);
    return run_nonseq(h, NS_PHILOX, nullptr, sample0, n, seed, out, summary);
}

extern "C" int psra_nonseq_eval_states(psra_handle *h, const uint32_t *states, int64_t n,
                                       const psra_nonseq_outputs *out, psra_nonseq_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, states != nullptr || n == 0, "null state matrix");
    return run_nonseq(h, NS_STATES, states, 0, n, 0, out, summary);
}

extern "C" int psra_nonseq_eval_uniforms(psra_handle *h, const double *r, int64_t n,
                                         const psra_nonseq_outputs *out, psra_nonseq_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, r != nullptr || n == 0, "null uniform matrix");
    return run_nonseq(h, NS_UNIFORMS, r, 0, n, 0, out, summary);
}

// multi_area.cu -- multi-area sequential adequacy with tie-line support (SURVEY f-3):
// run_fast_sequential_simulation / solve_curtailment_fast of
// GeneratingAdequacy/AdequacyAssessmentII.jl:73-179,185-250 behind psra_multi_area_mc.
//
// One thread block owns one simulated year at a time, as in seq_wide.cu: a lane owns one unit and walks
// its Philox stream block by block, adding the integer-MW deltas of its state changes to the hour
// timeline of the unit's AREA (one dense int32 timeline per area in shared memory); finished lanes pull
// the next unit from the block's queue.  Evaluation: per area the warps reduce the hour deltas to word
// sums, one warp per area scans them into the capacity entering every 32-hour word, and a word is
// resolved hour by hour (lane = hour) only if some area can be short in it (capacity + negative deltas
// < maximum load of the word).  ISOLATED: curtailment = deficit of the area (:84-92).  INTERCONNECTED:
// the lanes whose hour has a deficit run the reference's augmenting-path loop (:96-168) on integers --
// first surplus area, first deficit area, BFS in area order over residual tie capacities, never
// re-entering the source, stop when that sink cannot be reached -- so results match the literal loop
// bit for bit (all quantities are whole fixed-point numbers, the 1e-4 thresholds become > 0).
#include <limits.h>

#include <algorithm>
#include <vector>

#include "psra_internal.cuh"
#include "seq_args.cuh"

#define AREA_THREADS 256
#define AREA_MAX PSRA_MAX_AREAS

struct AreaArgs {
    int A, U, H, Wd, policy, init_mode;
    uint32_t k0, k1;
    long long year0, nyears;
    const int32_t *cap; const float *mttf; const float *mttr; const uint32_t *for_thr; const int32_t *order;
    const int32_t *unit_area;   // [U]
    const int32_t *load;        // [A][Wd*32] zero padded
    const int32_t *lmax;        // [A][Wd]
    const int32_t *topo;        // [A][A]
    uint32_t *lol; long long *ens;          // per year, per area (optional)
    unsigned long long *acc;    // [2*A]: sum of LOL hours, sum of curtailed energy per area; [2*A]: events
};

struct AreaShared {             // one per year parity
    int queue_head;
    int pad;
    int cap0[AREA_MAX];         // capacity of the units that start the year UP
    unsigned int lolh[AREA_MAX];
    unsigned long long ens[AREA_MAX];
};

static size_t area_smem_bytes(int A, int Wd)
{
    size_t b = sizeof(int32_t) * ((size_t)A * Wd * 32 + 32);            // hour timelines + one dummy slot per lane
    b += 4 * sizeof(int32_t) * (size_t)A * ((Wd + 3) & ~3);              // word sums, negative sums, entering capacity, load maxima
    b += 2 * sizeof(AreaShared) + 16;
    return (b + 15) & ~(size_t)15;
}

// reference solve_curtailment_fast, INTERCONNECTED branch, on integers (AdequacyAssessmentII.jl:96-176)
__device__ void area_max_flow(int n, const int32_t *__restrict__ topo, int *m)
{
    int res[AREA_MAX * AREA_MAX];
    for (int i = 0; i < n * n; i++) res[i] = topo[i];
    for (;;) {
        int src = -1, snk = -1;
        for (int i = n - 1; i >= 0; i--) { if (m[i] > 0) src = i; if (m[i] < 0) snk = i; }
        if (src < 0 || snk < 0) break;
        int parent[AREA_MAX], queue[AREA_MAX], qh = 0, qt = 0;
        bool found = false;
        for (int i = 0; i < n; i++) parent[i] = -1;
        queue[qt++] = src;
        while (qh < qt) {
            const int u = queue[qh++];
            if (u == snk) { found = true; break; }
            for (int v = 0; v < n; v++)
                if (res[u * n + v] > 0 && parent[v] < 0 && v != src) { parent[v] = u; queue[qt++] = v; }
        }
        if (!found) break;
        int f = min(m[src], -m[snk]);
        for (int c = snk; c != src; c = parent[c]) f = min(f, res[parent[c] * n + c]);
        m[src] -= f; m[snk] += f;
        for (int c = snk; c != src; c = parent[c]) { res[parent[c] * n + c] -= f; res[c * n + parent[c]] += f; }
    }
}

__global__ void __launch_bounds__(AREA_THREADS, 1) multi_area_kernel(const AreaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int A = a.A, Hpad = a.Wd * 32, Wd4 = (a.Wd + 3) & ~3;
    int32_t *tl = reinterpret_cast<int32_t *>(smem_raw);                 // [A][Hpad] + 32
    int32_t *wsum = tl + (size_t)A * Hpad + 32;                          // [A][Wd4]
    int32_t *wneg = wsum + A * Wd4;
    int32_t *wpre = wneg + A * Wd4;                                      // capacity entering the word
    int32_t *s_lmax = wpre + A * Wd4;
    AreaShared *sh_all = reinterpret_cast<AreaShared *>(s_lmax + A * Wd4);
    const uint32_t tl_s = (uint32_t)__cvta_generic_to_shared(tl);
    const uint32_t dummy_s = tl_s + 4u * (uint32_t)(A * Hpad + lane);

    for (int i = threadIdx.x; i < A * Wd4; i += blockDim.x) {
        const int ar = i / Wd4, w = i - ar * Wd4;
        s_lmax[i] = w < a.Wd ? a.lmax[ar * a.Wd + w] : 0;
    }
    for (int i = threadIdx.x; i < A * Hpad + 32; i += blockDim.x) tl[i] = 0;
    if (threadIdx.x < 2) {
        AreaShared *z = sh_all + threadIdx.x;
        z->queue_head = (int)blockDim.x;
        for (int i = 0; i < AREA_MAX; i++) { z->cap0[i] = 0; z->lolh[i] = 0u; z->ens[i] = 0ull; }
    }
    __syncthreads();

    unsigned long long acc_lol[AREA_MAX], acc_ens[AREA_MAX];            // thread 0 only
#pragma unroll
    for (int i = 0; i < AREA_MAX; i++) { acc_lol[i] = 0ull; acc_ens[i] = 0ull; }
    unsigned int n_events = 0;
    const unsigned long long end_t = (unsigned long long)a.H << PSRA_TICK_SHIFT;
    const unsigned long long parked = 0x00800000ull << 32;
    const bool stationary = a.init_mode == PSRA_INIT_STATIONARY;

    int par = 0;
    for (long long yl = blockIdx.x; yl < a.nyears; yl += gridDim.x, par ^= 1) {
        const unsigned long long chain = (unsigned long long)(a.year0 + yl);
        AreaShared *sh = sh_all + par;

        // ---- generation: lane = unit (seq_wide.cu), deltas go to the timeline of the unit's area
        int pos = threadIdx.x;
        bool busy = pos < a.U;
        int u = 0, cu = 0, ar = 0;
        float mup = 1.f, mdn = 1.f;
        uint32_t thr = 0u, nb = 0u, tla_s = tl_s;
        bool s0u = true;
        unsigned long long t = 0ull;
        auto take_unit = [&]() {
            u = __ldg(&a.order[pos]);
            cu = __ldg(&a.cap[u]);
            ar = __ldg(&a.unit_area[u]);
            mup = __fmul_rn(__ldg(&a.mttf[u]), 16777216.0f);
            mdn = __fmul_rn(__ldg(&a.mttr[u]), 16777216.0f);
            thr = __ldg(&a.for_thr[u]);
            tla_s = tl_s + 4u * (uint32_t)(ar * Hpad);
            nb = 0u;
            t = 0ull;
        };
        if (busy) take_unit();
        while (__any_sync(0xffffffffu, busy)) {
            uint32_t x[4];
            philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)u, nb, a.k0, a.k1, x);
            const bool first = nb == 0u;
            if (first) {                    // draw 0 of a stream is the initial state
                s0u = !(stationary && x[0] < thr);
                if (busy && s0u) atomicAdd(&sh->cap0[ar], cu);
            }
            const float m_a = s0u ? mdn : mup;
            const float m_b = s0u ? mup : mdn;
            const unsigned long long p1 = first ? 0ull : dur_ticks(m_a, x[0]);
            const unsigned long long p2 = p1 + dur_ticks(m_b, x[1]);
            const unsigned long long p3 = p2 + dur_ticks(m_a, x[2]);
            const unsigned long long p4 = p3 + dur_ticks(m_b, x[3]);
            const unsigned long long bm1 = busy ? t - 1ull : parked;
            const int delta_a = s0u ? cu : -cu;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned long long tm1 = bm1 + (q == 0 ? p1 : q == 1 ? p2 : q == 2 ? p3 : p4);
                if (q == 0 && first) tm1 = parked;
                const uint32_t hs = __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
                const bool ok = hs < (uint32_t)a.H;
                const uint32_t ad = ok ? tla_s + 4u * hs : dummy_s;
                asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(ad), "r"((q & 1) ? -delta_a : delta_a) : "memory");
                n_events += ok ? 1u : 0u;
            }
            t += p4;
            nb++;
            if (busy && t > end_t) {
                pos = atomicAdd(&sh->queue_head, 1);
                busy = pos < a.U;
                if (busy) take_unit();
            }
        }
        __syncthreads();

        // ---- word sums per area: lane = word, hour index skewed by the lane (conflict-free)
        for (int job = warp; job < A * ((a.Wd + 31) >> 5); job += nwarps) {
            const int tiles = (a.Wd + 31) >> 5;
            const int ja = job / tiles, w = (job - ja * tiles) * 32 + lane;
            if (w < a.Wd) {
                const int32_t *row = tl + ja * Hpad + w * 32;
                int s = 0, n = 0;
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const int d = row[(j + lane) & 31];
                    s += d;
                    n += min(d, 0);
                }
                wsum[ja * Wd4 + w] = s; wneg[ja * Wd4 + w] = n;
            }
        }
        __syncthreads();

        // ---- capacity entering every word: one warp per area, lane = run of `wpl` consecutive words
        for (int ja = warp; ja < A; ja += nwarps) {
            const int wpl = (a.Wd + 31) >> 5;
            const int wb = lane * wpl;
            int loc = 0;
            for (int k = 0; k < wpl; k++)
                if (wb + k < a.Wd) loc += wsum[ja * Wd4 + wb + k];
            int c = sh->cap0[ja] + warp_incl_scan(loc, lane) - loc;
            for (int k = 0; k < wpl; k++)
                if (wb + k < a.Wd) { wpre[ja * Wd4 + wb + k] = c; c += wsum[ja * Wd4 + wb + k]; }
        }
        __syncthreads();

        // ---- evaluation: a warp resolves and clears the words it owns
        {
            unsigned int lolh[AREA_MAX];
            long long ens_lane[AREA_MAX];
#pragma unroll
            for (int i = 0; i < AREA_MAX; i++) { lolh[i] = 0u; ens_lane[i] = 0ll; }
            for (int w = warp; w < a.Wd; w += nwarps) {
                bool short_possible = false;
                if (lane < A) short_possible = wpre[lane * Wd4 + w] + wneg[lane * Wd4 + w] < s_lmax[lane * Wd4 + w];
                if (__any_sync(0xffffffffu, short_possible)) {       // lane = hour of the word
                    const int h = w * 32 + lane;
                    int m[AREA_MAX];
                    bool deficit = false;
#pragma unroll
                    for (int ja = 0; ja < AREA_MAX; ja++) {
                        m[ja] = 0;
                        if (ja < A) {
                            const int c = wpre[ja * Wd4 + w] + warp_incl_scan(tl[ja * Hpad + h], lane);
                            m[ja] = c - __ldg(&a.load[ja * Hpad + h]);       // margin = generation - load (:221)
                            deficit |= m[ja] < 0;
                        }
                    }
                    deficit = deficit && h < a.H;
                    if (deficit && a.policy != 0) area_max_flow(A, a.topo, m);
                    __syncwarp();
#pragma unroll
                    for (int ja = 0; ja < AREA_MAX; ja++) {
                        if (ja < A) {
                            const bool cut = deficit && m[ja] < 0;              // curtailment > 0 (:231)
                            lolh[ja] += __popc(__ballot_sync(0xffffffffu, cut));
                            if (cut) ens_lane[ja] += (long long)(-m[ja]);
                        }
                    }
                }
                for (int ja = 0; ja < A; ja++) tl[ja * Hpad + w * 32 + lane] = 0;
            }
#pragma unroll
            for (int ja = 0; ja < AREA_MAX; ja++) {
                if (ja < A && lolh[ja]) {                                      // uniform within the warp
                    const long long e = warp_sum_ll(ens_lane[ja]);
                    if (lane == 0) { atomicAdd(&sh->lolh[ja], lolh[ja]); atomicAdd(&sh->ens[ja], (unsigned long long)e); }
                }
            }
        }
        __syncthreads();

        // ---- per-year results: thread 0 writes the year out and re-arms its scalars for the year after next
        if (threadIdx.x == 0) {
            sh->queue_head = (int)blockDim.x;
            for (int ja = 0; ja < A; ja++) {
                const unsigned int l = sh->lolh[ja];
                const unsigned long long e = sh->ens[ja];
                sh->lolh[ja] = 0u; sh->ens[ja] = 0ull; sh->cap0[ja] = 0;
                if (a.lol) a.lol[yl * A + ja] = l;
                if (a.ens) a.ens[yl * A + ja] = (long long)e;
                acc_lol[ja] += l; acc_ens[ja] += e;
            }
        }
    }

    unsigned long long ev = n_events;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ev += __shfl_xor_sync(0xffffffffu, ev, d);
    if (lane == 0 && ev) atomicAdd(&a.acc[2 * AREA_MAX], ev);
    if (threadIdx.x == 0) {
        for (int ja = 0; ja < A; ja++) {
            if (acc_lol[ja]) atomicAdd(&a.acc[ja], acc_lol[ja]);
            if (acc_ens[ja]) atomicAdd(&a.acc[AREA_MAX + ja], acc_ens[ja]);
        }
    }
}

// -------------------------------------------------------------------------------------- host side
extern "C" int psra_multi_area_mc(psra_handle *h, const psra_area_system *sys, int32_t policy, int64_t year0,
                                  int64_t nyears, uint64_t seed, int32_t init_mode, const psra_area_outputs *out,
                                  psra_area_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, sys && summary, "null system / summary");
    PSRA_REQUIRE(h, sys->n_areas >= 1 && sys->n_areas <= PSRA_MAX_AREAS, "number of areas out of range (1..PSRA_MAX_AREAS)");
    PSRA_REQUIRE(h, sys->n_units >= 1 && sys->n_hours >= 1 && sys->n_hours <= (1 << 20), "bad unit / hour counts");
    PSRA_REQUIRE(h, sys->unit_area && sys->cap_fp && sys->mttf_h && sys->mttr_h && sys->load_fp && sys->topology_fp, "null array in psra_area_system");
    PSRA_REQUIRE(h, policy == PSRA_POLICY_ISOLATED || policy == PSRA_POLICY_INTERCONNECTED, "unknown support policy");
    PSRA_REQUIRE(h, init_mode == PSRA_INIT_ALL_UP || init_mode == PSRA_INIT_STATIONARY, "unknown init_mode");
    PSRA_REQUIRE(h, year0 >= 0 && nyears >= 0, "negative year range");
    const int A = sys->n_areas, U = sys->n_units, H = sys->n_hours, Wd = (H + 31) / 32, Hpad = Wd * 32;
    for (int u = 0; u < U; u++) PSRA_REQUIRE(h, sys->unit_area[u] >= 0 && sys->unit_area[u] < A, "unit_area out of range");
    for (int i = 0; i < A * A; i++) PSRA_REQUIRE(h, sys->topology_fp[i] >= 0 && sys->topology_fp[i] <= 0x3fffffff, "tie capacity out of range");
    memset(summary, 0, sizeof(*summary));
    summary->years = nyears; summary->n_areas = A;
    int rc = PSRA_OK;
    if (nyears == 0) return PSRA_OK;
    PSRA_CUDA(h, cudaSetDevice(h->device));

    // private tables of this call in one scratch buffer (the handle's own system and load -- psra_set_system /
    // psra_set_load -- stay as they are): [unit_area U][load A*Hpad][lmax A*Wd][topo A*A][cap U][mttf U][mttr U]
    // [FOR thresholds U][unit order U]; the sampler quantities are those of psra_set_system (binary32 means,
    // floor(FOR 2^32), units ordered by decreasing transition rate)
    std::vector<int32_t> buf((size_t)U + (size_t)A * Hpad + (size_t)A * Wd + (size_t)A * A + 5 * (size_t)U, 0);
    int32_t *b_area = buf.data(), *b_load = b_area + U, *b_lmax = b_load + (size_t)A * Hpad, *b_topo = b_lmax + (size_t)A * Wd;
    int32_t *b_cap = b_topo + (size_t)A * A, *b_mttf = b_cap + U, *b_mttr = b_mttf + U, *b_thr = b_mttr + U, *b_order = b_thr + U;
    {
        int64_t total = 0;
        for (int u = 0; u < U; u++) {
            PSRA_REQUIRE(h, sys->cap_fp[u] >= 0, "negative capacity");
            PSRA_REQUIRE(h, sys->mttf_h[u] > 0 && sys->mttr_h[u] > 0 && sys->mttf_h[u] <= PSRA_MAX_MEAN_HOURS && sys->mttr_h[u] <= PSRA_MAX_MEAN_HOURS,
                         "MTTF / MTTR must be positive (at most 1e8 hours)");
            total += sys->cap_fp[u];
            b_cap[u] = sys->cap_fp[u];
            const float mf = (float)sys->mttf_h[u], mr = (float)sys->mttr_h[u];
            memcpy(&b_mttf[u], &mf, 4); memcpy(&b_mttr[u], &mr, 4);
            const double lam = 1.0 / sys->mttf_h[u], mu = 1.0 / sys->mttr_h[u];      // AdequacyAssessmentII.jl:15-26 rates
            const double t = floor(lam / (lam + mu) * 4294967296.0);
            const uint32_t thr = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
            memcpy(&b_thr[u], &thr, 4);
            b_order[u] = u;
        }
        PSRA_REQUIRE(h, total <= 0x3fffffff, "installed capacity exceeds the int32 fixed-point range");
        std::stable_sort(b_order, b_order + U, [&](int32_t x, int32_t y) { return sys->mttf_h[x] + sys->mttr_h[x] < sys->mttf_h[y] + sys->mttr_h[y]; });
    }
    for (int u = 0; u < U; u++) b_area[u] = sys->unit_area[u];
    for (int ar = 0; ar < A; ar++)
        for (int i = 0; i < H; i++) {
            const int32_t v = sys->load_fp[(size_t)ar * H + i];
            PSRA_REQUIRE(h, v >= 0 && v <= 0x3fffffff, "load out of the int32 fixed-point range");
            b_load[(size_t)ar * Hpad + i] = v;
            b_lmax[(size_t)ar * Wd + (i >> 5)] = std::max(b_lmax[(size_t)ar * Wd + (i >> 5)], v);
        }
    for (int i = 0; i < A * A; i++) b_topo[i] = sys->topology_fp[i];
    rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, sizeof(int32_t) * buf.size());
    if (rc) return rc;
    PSRA_CUDA(h, cudaMemcpyAsync(h->d_scratch, buf.data(), sizeof(int32_t) * buf.size(), cudaMemcpyHostToDevice, h->stream));

    AreaArgs a{};
    a.A = A; a.U = U; a.H = H; a.Wd = Wd; a.policy = policy; a.init_mode = init_mode;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.year0 = year0; a.nyears = nyears;
    a.unit_area = (const int32_t *)h->d_scratch;
    a.load = a.unit_area + U; a.lmax = a.load + (size_t)A * Hpad; a.topo = a.lmax + (size_t)A * Wd;
    a.cap = a.topo + (size_t)A * A;
    a.mttf = reinterpret_cast<const float *>(a.cap + U); a.mttr = a.mttf + U;
    a.for_thr = reinterpret_cast<const uint32_t *>(a.mttr + U);
    a.order = reinterpret_cast<const int32_t *>(a.for_thr + U);
    a.acc = h->d_acc;
    const bool want_vec = out && (out->lol_hours || out->ens_fp);
    if (want_vec) {
        rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, (sizeof(uint32_t) + sizeof(long long)) * (size_t)nyears * A);
        if (rc) return rc;
        a.ens = (long long *)h->d_scratch2;
        a.lol = (uint32_t *)(a.ens + (size_t)nyears * A);
    }
    const size_t smem = area_smem_bytes(A, Wd);
    if (smem > h->smem_optin)
        return psra_fail(h, PSRA_E_INVALID, "multi-area timelines do not fit shared memory (%zu B needed, %zu B available): fewer areas or hours", smem, h->smem_optin);
    PSRA_CUDA(h, cudaFuncSetAttribute(multi_area_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    PSRA_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, multi_area_kernel, AREA_THREADS, smem));
    if (bps < 1) return psra_fail(h, PSRA_E_CUDA, "multi-area kernel does not fit on an SM (smem %zu B)", smem);
    long long grid = std::min<long long>((long long)h->sm_count * bps, nyears);
    PSRA_CUDA(h, cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * ACC_COUNT, h->stream));
    PSRA_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    multi_area_kernel<<<(unsigned)grid, AREA_THREADS, smem, h->stream>>>(a);
    PSRA_CUDA(h, cudaGetLastError());
    PSRA_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    unsigned long long acc[ACC_COUNT];
    PSRA_CUDA(h, cudaMemcpyAsync(acc, h->d_acc, sizeof(acc), cudaMemcpyDeviceToHost, h->stream));
    if (want_vec) {
        if (out->lol_hours) PSRA_CUDA(h, cudaMemcpyAsync(out->lol_hours, a.lol, sizeof(uint32_t) * (size_t)nyears * A, cudaMemcpyDeviceToHost, h->stream));
        if (out->ens_fp) PSRA_CUDA(h, cudaMemcpyAsync(out->ens_fp, a.ens, sizeof(long long) * (size_t)nyears * A, cudaMemcpyDeviceToHost, h->stream));
    }
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    PSRA_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    summary->kernel_ms = ms;
    for (int ar = 0; ar < A; ar++) {
        summary->sum_lol_hours[ar] = (int64_t)acc[ar];
        summary->sum_ens_fp[ar] = (int64_t)acc[AREA_MAX + ar];
    }
    summary->events = acc[2 * AREA_MAX];
    return PSRA_OK;
}

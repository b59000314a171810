// seq_team.cu -- sampler-driven sequential chronological MC for systems of more than 32 units
// (BASELINE config 5: 1024 units).  Same model, sampler and per-year integers as seq_fast.cu /
// seq_mc.cu (run_sequential_mc, GeneratingAdequacy/PowerSystemAdequacy.jl:214-269; indices per
// Montecarlo_seq/seqMain.m:160-176, Montecarlo_seq/calnlc.m:22-34).
//
// One thread block owns one chain of years at a time ("each block owns simulated years").  The
// units are split in groups of 32; the warps of the block share the groups.  For every timeline
// segment each warp runs the generation waves of seq_fast.cu for its groups (lane = one Philox
// block job of a unit of the group; prefix sums of tick durations; segmented shuffle scan) and
// scatters the integer MW deltas with shared-memory atomics into ONE block-shared ring of two
// timeline segments (int32 per hour) plus per-32-hour-word sums.  After a block barrier warp 0
// evaluates the segment exactly like seq_fast.cu (shuffle scan over word sums, conservative flag
// against the per-word maximum load, hour-by-hour ballot/popc resolution of flagged runs against
// the load curve staged in shared memory) while the other warps wait; the evaluated half is
// cleared by the whole block.  Far-future events go to a block-shared pending list (atomic append,
// double buffered).
#include <limits.h>

#include "psra_internal.cuh"
#include "seq_args.cuh"

#define TEAM_NB_MAX 4
// pending entries: (hour << 12) | (unit << 1) | (delta > 0); capacity per buffer scales with the unit count
#define TEAM_MAX_UNITS 2048
#define TEAM_WARPS 8

struct TeamShared {                   // scalars shared by the block
    int pend_cnt[2];
    int capacity;
    int overflow;
};

int seq_team_pend_cap(int U) { const int c = 4 * ((U + 31) & ~31); return c < 1024 ? 1024 : c; }

size_t seq_team_smem_bytes(int U, int Wd, int seg_words, bool two_halves)
{
    const int halves = two_halves ? 2 : 1;
    const size_t Upad = (size_t)((U + 31) & ~31);
    size_t b = sizeof(int32_t) * (size_t)((Wd + 3) & ~3);                                  // word maxima of the load
    b += Upad * (sizeof(unsigned long long) + sizeof(int32_t) + 2 * sizeof(float) + 2 * sizeof(uint32_t));  // t_run, cap, means, thr, nb
    b += sizeof(uint32_t) * (size_t)(((Upad / 32) + 3) & ~3);                             // initial-state masks
    b += sizeof(int32_t) * halves * (size_t)seg_words * 32;                               // ring timeline
    b += 2 * sizeof(int32_t) * (size_t)((halves * seg_words + 3) & ~3);                    // word sums
    b += two_halves ? 2 * sizeof(uint32_t) * (size_t)seq_team_pend_cap(U) : 0;             // pending lists
    b += TEAM_WARPS * 32 * TEAM_NB_MAX;                                                   // job maps
    b += sizeof(TeamShared) + 64;
    return (b + 15) & ~(size_t)15;
}

int seq_team_max_units() { return TEAM_MAX_UNITS; }

__global__ void __launch_bounds__(TEAM_WARPS * 32, 3) seq_team_kernel(const SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int Upad = (a.U + 31) & ~31, G = Upad >> 5;
    const bool two_halves = a.two_halves != 0;
    const int halves = two_halves ? 2 : 1;
    const int PEND_CAP = two_halves ? a.pend_cap : 0;
    const int seg_slots = a.seg_words * 32;
    const int ring_words = (halves * a.seg_words + 3) & ~3;
    // ---- shared-memory carve-up (8-byte items first)
    unsigned long long *t_run = reinterpret_cast<unsigned long long *>(smem_raw);          // [Upad]
    int32_t *s_lmax = reinterpret_cast<int32_t *>(t_run + Upad);                            // [pad4(Wd)]
    const int32_t *s_load = a.load;     // the hourly curve is only read when a flagged run is resolved: global / L1
    int32_t *s_cap = s_lmax + ((a.Wd + 3) & ~3);                                            // [Upad]
    float *s_mup = reinterpret_cast<float *>(s_cap + Upad);
    float *s_mdn = s_mup + Upad;
    uint32_t *s_thr = reinterpret_cast<uint32_t *>(s_mdn + Upad);
    uint32_t *s_nb = s_thr + Upad;                                                          // next Philox block per unit
    uint32_t *s_s0 = s_nb + Upad;                                                           // [G] initial-state masks
    int32_t *tl = reinterpret_cast<int32_t *>(s_s0 + ((G + 3) & ~3));                       // [2*seg_slots]
    int32_t *wsum = tl + halves * seg_slots;
    int32_t *wneg = wsum + ring_words;
    uint32_t *pend = reinterpret_cast<uint32_t *>(wneg + ring_words);                       // [2][CAP]
    unsigned char *jobmap = reinterpret_cast<unsigned char *>(pend + 2 * PEND_CAP) + warp * 32 * TEAM_NB_MAX;
    TeamShared *sh = reinterpret_cast<TeamShared *>(
        (reinterpret_cast<uintptr_t>(pend + 2 * PEND_CAP) + TEAM_WARPS * 32 * TEAM_NB_MAX + 15) & ~(uintptr_t)15);

    for (int i = threadIdx.x; i < a.Wd; i += blockDim.x) s_lmax[i] = a.lmax[i];
    for (int i = threadIdx.x; i < Upad; i += blockDim.x) {
        const bool v = i < a.U;
        s_cap[i] = v ? a.cap[i] : 0;
        s_mup[i] = v ? __fmul_rn(a.mttf[i], 16777216.0f) : 1.0f;
        s_mdn[i] = v ? __fmul_rn(a.mttr[i], 16777216.0f) : 1.0f;
        s_thr[i] = v ? a.for_thr[i] : 0u;
    }
    for (int i = threadIdx.x; i < halves * seg_slots; i += blockDim.x) tl[i] = 0;
    for (int i = threadIdx.x; i < ring_words; i += blockDim.x) { wsum[i] = 0; wneg[i] = 0; }
    if (threadIdx.x == 0) { sh->pend_cnt[0] = sh->pend_cnt[1] = 0; sh->capacity = 0; sh->overflow = 0; }
    __syncthreads();

    unsigned long long acc_lol = 0, acc_ent = 0, acc_ywl = 0, acc_lol2 = 0, acc_e2lo = 0, acc_e2hi = 0;
    long long acc_ens = 0;
    unsigned int n_events = 0, n_waves = 0, n_jobs = 0, n_opt = 0, n_flag = 0;
    const int chain_end_h = a.ypc * a.H;

    for (long long cl = blockIdx.x; cl < a.nchains; cl += gridDim.x) {
        const unsigned long long chain = (unsigned long long)(a.chain_base + cl);
        for (int i = threadIdx.x; i < Upad; i += blockDim.x) {
            t_run[i] = a.disc ? (1ull << PSRA_TICK_SHIFT) : 0ull;
            s_nb[i] = 0u;
        }
        if (threadIdx.x == 0) { sh->capacity = 0; sh->pend_cnt[0] = sh->pend_cnt[1] = 0; }
        int ring = 0, pb = 0;          // ring half of the current segment, active pending buffer
        bool init_wave = true;
        __syncthreads();

        for (int y = 0; y < a.ypc; y++) {
            unsigned int lolh = 0, entries = 0;      // meaningful in warp 0
            long long ens_lane = 0;
            for (int seg = 0; seg < a.nseg; seg++, ring = two_halves ? ring ^ 1 : 0) {
                const int seg_h0 = seg * seg_slots;
                const int seg_h1 = min(a.H, seg_h0 + seg_slots);
                const int abs0 = y * a.H + seg_h0, abs1 = y * a.H + seg_h1;
                const int nxt_h0 = (seg + 1 < a.nseg) ? seg_h0 + seg_slots : 0;
                const int abs2 = min(chain_end_h, abs1 + min(seg_slots, a.H - nxt_h0));
                const unsigned long long seg_end_t = (unsigned long long)abs1 << PSRA_TICK_SHIFT;
                const unsigned long long nxt_end_t = (unsigned long long)abs2 << PSRA_TICK_SHIFT;
                const uint32_t len_cur = (uint32_t)(seg_h1 - seg_h0);
                const uint32_t ring_len = (uint32_t)(abs2 - abs0);
                const int pad = seg_slots - (int)len_cur;
                const int ring_base = ring * seg_slots;
                const int wbase_cur = ring * a.seg_words;
                auto ring_slot = [&](uint32_t rel) -> int {
                    int idx = ring_base + (int)rel + (rel >= len_cur ? pad : 0);
                    return idx >= 2 * seg_slots ? idx - 2 * seg_slots : idx;
                };
                uint32_t *pend_in = pend + pb * PEND_CAP, *pend_out = pend + (pb ^ 1) * PEND_CAP;

                // ---- pending events: scatter those that now fall into the ring, keep the rest (other buffer)
                const int n_in = two_halves ? sh->pend_cnt[pb] : 0;
                if (n_in) {
                    for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
                        const uint32_t e = pend_in[i];
                        const int hs = (int)(e >> 12);
                        if (hs < abs2) {
                            const int c = s_cap[(e >> 1) & 2047];
                            const int slot = ring_slot((uint32_t)(hs - abs0));
                            atomicAdd(&tl[slot], (e & 1u) ? c : -c);
                            atomicAdd(&wsum[slot >> 5], (e & 1u) ? c : -c);
                            if (!(e & 1u)) atomicAdd(&wneg[slot >> 5], -c);
                        } else {
                            const int pos = atomicAdd(&sh->pend_cnt[pb ^ 1], 1);
                            if (pos < PEND_CAP) pend_out[pos] = e; else sh->overflow = 1;
                        }
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) sh->pend_cnt[pb] = 0;
                pb ^= 1;                                   // new events append to the buffer that now holds the kept ones
                uint32_t *pend_act = pend + pb * PEND_CAP;
                __syncthreads();

                // ---- generation: every warp serves its unit groups
                for (int g = warp; g < G; g += TEAM_WARPS) {
                    const int ub = g * 32;                         // first unit of the group
                    const int ug = ub + lane;                      // this lane's unit (owner role)
                    const bool unit_valid = ug < a.U;
                    const float inv_span = unit_valid ? __fdividef(0.5f * 16777216.0f, s_mup[ug] + s_mdn[ug]) : 0.f;
                    uint32_t nb = s_nb[ug];
                    uint32_t s0mask = init_wave ? 0u : s_s0[g];
                    bool first = init_wave;
                    while (true) {
                        const unsigned long long tlast = t_run[ug];
                        const bool is_short = unit_valid && tlast <= seg_end_t;
                        if (!__any_sync(0xffffffffu, is_short)) break;
                        int want = 0;
                        if (first) want = unit_valid ? 1 : 0;
                        else if (unit_valid && tlast <= nxt_end_t) {
                            const float rem_h = (float)(int)((nxt_end_t - tlast) >> PSRA_TICK_SHIFT);
                            want = min(TEAM_NB_MAX, 1 + (int)(rem_h * inv_span));
                        }
                        auto count_scan = [&](int n, int &excl, int &total) {
                            const uint32_t b0 = __ballot_sync(0xffffffffu, n & 1), b1 = __ballot_sync(0xffffffffu, n & 2),
                                           b2 = __ballot_sync(0xffffffffu, n & 4);
                            excl = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
                            total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                        };
                        int n_m = is_short ? want : 0;
                        int off_m, J1;
                        count_scan(n_m, off_m, J1);
                        if (J1 < 32 && !first) {       // spare lanes: one more block of slack for the short units
                            const int n_x = is_short ? min(TEAM_NB_MAX, want + 1) : 0;
                            int off_x, Jx;
                            count_scan(n_x, off_x, Jx);
                            if (Jx <= 32) { n_m = n_x; off_m = off_x; J1 = Jx; }
                        }
                        int n_u, off, J;
                        if (J1 >= 32 || first || !two_halves) {
                            off = off_m;
                            n_u = max(0, min(n_m, 32 - off));
                            J = min(J1, 32);
                        } else {
                            const int n_o = (!is_short && sh->pend_cnt[pb] <= PEND_CAP / 2) ? want : 0;
                            int off_o, J2;
                            count_scan(n_o, off_o, J2);
                            off_o += J1;
                            off = is_short ? off_m : off_o;
                            n_u = is_short ? n_m : max(0, min(n_o, 32 - off_o));
                            J = min(32, J1 + J2);
                            n_opt += J - J1;
                        }
                        n_jobs += J;
                        n_waves++;
#pragma unroll
                        for (int k = 0; k < TEAM_NB_MAX; k++)
                            if (k < n_u) jobmap[off + k] = (unsigned char)lane;
                        __syncwarp();
                        {
                            const bool act = lane < J;
                            const int ul = act ? (int)jobmap[lane] : 0;      // unit within the group
                            const int u = ub + ul;                           // global unit
                            const int offu = __shfl_sync(0xffffffffu, off, ul);
                            const int nu = __shfl_sync(0xffffffffu, n_u, ul);
                            const uint32_t b = __shfl_sync(0xffffffffu, nb, ul) + (uint32_t)(lane - offu);
                            const bool is_last = act && (lane - offu) == nu - 1;
                            uint32_t x[4];
                            philox4x32_10_rk((uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)u, b, a.rk, x);
                            bool s0u;
                            if (b == 0u) s0u = !(a.init_mode == PSRA_INIT_STATIONARY && x[0] < s_thr[u]);
                            else s0u = (s0mask >> ul) & 1u;
                            const float mup = s_mup[u], mdn = s_mdn[u];
                            const float m_a = s0u ? mdn : mup, m_b = s0u ? mup : mdn;
                            const unsigned long long p1 = (b == 0u) ? 0ull : dur_ticks_disc(m_a, x[0], !s0u, a.disc);
                            const unsigned long long p2 = p1 + dur_ticks_disc(m_b, x[1], s0u, a.disc);
                            const unsigned long long p3 = p2 + dur_ticks_disc(m_a, x[2], !s0u, a.disc);
                            const unsigned long long p4 = p3 + dur_ticks_disc(m_b, x[3], s0u, a.disc);
                            const unsigned long long tot = act ? p4 : 0ull;
                            unsigned long long inc2 = tot;
#pragma unroll
                            for (int d = 1; d <= 2; d <<= 1) {
                                const unsigned long long o = __shfl_up_sync(0xffffffffu, inc2, d);
                                if (lane - d >= offu) inc2 += o;
                            }
                            const unsigned long long base_t = t_run[u] + (inc2 - tot);
                            __syncwarp();
                            if (is_last) t_run[u] = base_t + tot;
                            const int cu = s_cap[u];
                            const int delta_a = s0u ? cu : -cu;
                            if (b == 0u && act) s0mask = s0u ? 1u : 0u;
                            const unsigned long long bm1 = base_t - 1ull;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const unsigned long long tm1 = bm1 + (q == 0 ? p1 : q == 1 ? p2 : q == 2 ? p3 : p4);
                                const uint32_t hs = __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
                                const bool valid = act && !(b == 0u && q == 0);
                                const uint32_t rel = hs - (uint32_t)abs0;
                                const int delta = (q & 1) ? -delta_a : delta_a;
                                const bool in_ring = valid && rel < ring_len;
                                if (in_ring) {
                                    const int slot = ring_slot(rel);
                                    atomicAdd(&tl[slot], delta);
                                    atomicAdd(&wsum[slot >> 5], delta);
                                    if (delta < 0) atomicAdd(&wneg[slot >> 5], delta);
                                }
                                const bool inhor = valid && hs < (uint32_t)chain_end_h;
                                const bool pnd = two_halves && inhor && !in_ring;
                                const uint32_t pm = __ballot_sync(0xffffffffu, pnd);
                                if (pm) {
                                    int basep = 0;
                                    if (lane == 0) basep = atomicAdd(&sh->pend_cnt[pb], __popc(pm));
                                    basep = __shfl_sync(0xffffffffu, basep, 0);
                                    const int pos = basep + __popc(pm & lt_mask);
                                    if (pnd) {
                                        if (pos < PEND_CAP)
                                            pend_act[pos] = (hs << 12) | ((uint32_t)u << 1) | (delta > 0 ? 1u : 0u);
                                        else sh->overflow = 1;
                                    }
                                }
                                n_events += inhor ? 1u : 0u;
                            }
                            __syncwarp();
                        }
                        nb += (uint32_t)n_u;
                        if (first) {
                            s0mask = __ballot_sync(0xffffffffu, unit_valid && (s0mask & 1u));
                            int cp = (unit_valid && ((s0mask >> lane) & 1u)) ? s_cap[ug] : 0;
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) cp += __shfl_xor_sync(0xffffffffu, cp, d);
                            if (lane == 0) { atomicAdd(&sh->capacity, cp); s_s0[g] = s0mask; }
                            first = false;
                        }
                    }
                    s_nb[ug] = nb;
                    __syncwarp();
                }
                init_wave = false;
                __syncthreads();

                // ---- evaluation of the current half by warp 0: lane = run of `wpl` consecutive words
                const int nwords = (seg_h1 - seg_h0 + 31) >> 5;
                if (warp == 0) {
                    int capacity = sh->capacity;
                    const int wpl = (nwords + 31) >> 5;
                    const int wb = lane * wpl;
                    int loc = 0, lmin = INT_MAX;
                    for (int k = 0; k < wpl; k++) {
                        const int w = wb + k;
                        if (w < nwords) {
                            lmin = min(lmin, loc + wneg[wbase_cur + w] - s_lmax[seg * a.seg_words + w]);
                            loc += wsum[wbase_cur + w];
                        }
                    }
                    const int incl = warp_incl_scan(loc, lane);
                    const int cs_lane = capacity + incl - loc;
                    const bool flagged = (lmin != INT_MAX) && (cs_lane + lmin < 0);
                    uint32_t fm = __ballot_sync(0xffffffffu, flagged);
                    n_flag += __popc(fm);
                    while (fm) {
                        const int src = __ffs(fm) - 1;
                        fm &= fm - 1;
                        int c_in = __shfl_sync(0xffffffffu, cs_lane, src);
                        for (int k = 0; k < wpl; k++) {
                            const int wq = src * wpl + k;
                            if (wq >= nwords) break;
                            const int c = c_in + warp_incl_scan(tl[ring_base + wq * 32 + lane], lane);
                            const int hy0 = seg_h0 + wq * 32;
                            const int L = __ldg(&s_load[hy0 + lane]);
                            const bool lol = c < L;                  // PSA.jl:253 strict
                            const uint32_t mask = __ballot_sync(0xffffffffu, lol);
                            if (mask) {
                                const uint32_t prev = (hy0 > 0 && c_in < __ldg(&s_load[hy0 - 1])) ? 1u : 0u;
                                lolh += __popc(mask);
                                entries += __popc(mask & ~((mask << 1) | prev));   // calnlc.m:22-34
                                if (lol) {
                                    ens_lane += (long long)(L - c);
                                    if (a.fail) atomicAdd(&a.fail[hy0 + lane], 1u);
                                }
                            }
                            c_in = __shfl_sync(0xffffffffu, c, 31);
                        }
                    }
                    capacity += __shfl_sync(0xffffffffu, incl, 31);
                    __syncwarp();
                    if (lane == 0) sh->capacity = capacity;
                }
                __syncthreads();
                {   // clear the evaluated half
                    int4 *t4 = reinterpret_cast<int4 *>(tl + ring_base);
                    for (int i = threadIdx.x; i < seg_slots / 4; i += blockDim.x) t4[i] = make_int4(0, 0, 0, 0);
                    for (int i = threadIdx.x; i < a.seg_words; i += blockDim.x) { wsum[wbase_cur + i] = 0; wneg[wbase_cur + i] = 0; }
                }
                __syncthreads();
            }

            // ---- per-year indices (warp 0)
            if (warp == 0) {
                long long ens = 0;
                if (lolh) ens = warp_sum_ll(ens_lane);
                const long long yi = cl * a.ypc + y;
                if (lane == 0) {
                    if (a.lol) a.lol[yi] = lolh;
                    if (a.ens) a.ens[yi] = ens;
                    if (a.ent) a.ent[yi] = entries;
                    if (a.group_lol && lolh) atomicAdd(&a.group_lol[yi / a.group], (unsigned long long)lolh);
                    if (lolh) seq_hist_add(a, ens);
                }
                acc_lol += lolh; acc_ens += ens; acc_ent += entries;
                acc_ywl += lolh ? 1 : 0;
                acc_lol2 += (unsigned long long)lolh * lolh;
                const unsigned long long e = (unsigned long long)ens;
                const unsigned long long plo = e * e, phi = __umul64hi(e, e);
                const unsigned long long nlo = acc_e2lo + plo;
                acc_e2hi += phi + (nlo < acc_e2lo ? 1ull : 0ull);
                acc_e2lo = nlo;
            }
        }
        __syncthreads();
    }

    unsigned long long ev = n_events;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ev += __shfl_xor_sync(0xffffffffu, ev, d);
    if (lane == 0) {
        if (ev) atomicAdd(&a.acc[ACC_EVENTS], ev);
        atomicAdd(&a.acc[ACC_WAVES], (unsigned long long)n_waves);
        atomicAdd(&a.acc[ACC_JOBS], (unsigned long long)n_jobs);
        atomicAdd(&a.acc[ACC_OPT_JOBS], (unsigned long long)n_opt);
        if (warp == 0) {
            if (acc_lol) atomicAdd(&a.acc[ACC_LOL], acc_lol);
            if (acc_ens) atomicAdd(&a.acc[ACC_ENS], (unsigned long long)acc_ens);
            if (acc_ent) atomicAdd(&a.acc[ACC_ENT], acc_ent);
            if (acc_ywl) atomicAdd(&a.acc[ACC_YWL], acc_ywl);
            if (acc_lol2) atomicAdd(&a.acc[ACC_LOL2], acc_lol2);
            if (acc_e2lo | acc_e2hi) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], acc_e2lo, acc_e2hi);
            atomicAdd(&a.acc[ACC_FLAGGED], (unsigned long long)n_flag);
            if (sh->overflow) atomicExch(&a.acc[ACC_OVERFLOW], 2ull);
        }
    }
}

cudaError_t seq_team_prepare(size_t smem, int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(seq_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, seq_team_kernel, TEAM_WARPS * 32, smem);
}

void seq_team_launch(const SeqArgs &a, unsigned grid, size_t smem, cudaStream_t stream)
{
    seq_team_kernel<<<grid, TEAM_WARPS * 32, smem, stream>>>(a);
}

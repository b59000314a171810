// seq_mc.cu -- sequential chronological Monte Carlo on sm_100a.
//
// Replaces the body of run_sequential_mc (GeneratingAdequacy/PowerSystemAdequacy.jl:214-269)
// and adds the per-year index definitions of Montecarlo_seq/seqMain.m:160-176 +
// Montecarlo_seq/calnlc.m:22-34 (DLC = LOL hours, ENS, NLC = deficit entries).
//
// Formulation.  The reference walks every hour and decrements every
// unit's residual time (PSA.jl:237-250).  This kernel uses the bit-equivalent event form:
// a unit whose residual is t > 0 after hour g toggles next in hour g + ceil(t); the residual
// left after that hour's decrement is r = fl(fl(t - (ceil(t) - 1)) - 1) (all earlier
// decrements are exact in FP64 because t >= 1 there); the next residual is fl(r + D) with D
// the next duration, and while it is <= 0 the unit toggles again in the same hour
// (PSA.jl:240-248).  One warp owns one chain of years at a time:
//   1. event generation: lane = unit (lane-strided over units when U > 32); each lane draws
//      durations (Philox4x32-10 keyed (seed; chain, unit), or an injected list) and scatters
//      the integer capacity delta of every toggle into the warp's shared-memory hour
//      timeline (atomicAdd) and sets the hour's bit in an event bitmap (atomicOr);
//   2. evaluation: lane = 32-hour word of the timeline.  A lane gathers the deltas of the
//      set bits of its word, a warp-shuffle prefix scan turns the word sums into the
//      capacity at every word start, and a word is "flagged" when a lower bound of its
//      minimum capacity is below the word's maximum load (table staged in shared memory).
//      Flagged words are rare; the warp resolves them hour by hour (lane = hour): shuffle
//      scan of the 32 deltas, compare against the staged load curve, __ballot_sync/__popc
//      for LOL hours and deficit entries, per-lane ENS accumulators reduced once per year.
// A year is processed in `nseg` timeline segments so that the per-warp shared-memory
// footprint stays small enough for >= 16 resident warps per SM; unit states live in
// registers between segments (U <= 32) or in shared memory (U > 32, multi-year chains).
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "psra_internal.cuh"
#include "seq_args.cuh"


struct UnitState {
    double r;        // residual after the decrement of hour `next` (<= 0)
    int next;        // chain-relative 0-based hour slot of the next toggle
    uint32_t j;      // next draw index of the unit's stream
    int status;      // 1 = UP
    uint32_t buf[4]; // unread words of the current Philox block (buf[0] is next)
};

template <bool kInjected>
__device__ __forceinline__ uint32_t next_word(const SeqArgs &a, long long chain, int u, UnitState &s)
{
    if ((s.j & 3u) == 0u)
        philox4x32_10((uint32_t)chain, (uint32_t)((unsigned long long)chain >> 32), (uint32_t)u, s.j >> 2,
                      a.k0, a.k1, s.buf);
    const uint32_t x = s.buf[0];
    s.buf[0] = s.buf[1]; s.buf[1] = s.buf[2]; s.buf[2] = s.buf[3];
    s.j++;
    return x;
}

// duration of the state the unit has just entered (PSA.jl:243 / :246 / :224)
template <bool kInjected>
__device__ __forceinline__ double next_duration(const SeqArgs &a, long long chain_local, long long chain, int u,
                                                UnitState &s, float mf, float mr)
{
    if constexpr (kInjected) {
        if (s.j >= (uint32_t)a.K) {
            atomicExch(&a.acc[ACC_OVERFLOW], 1ull);
            return 1.0e300;
        }
        const double d = __ldg(&a.dur[((size_t)chain_local * a.U + u) * (size_t)a.K + s.j]);
        s.j++;
        return d;
    } else {
        const uint32_t x = next_word<false>(a, chain, u, s);
        // tick-quantised duration (2^-24 h), same definition as seq_fast.cu / DESIGN.md "Sampler"
        const float mt = __fmul_rn(s.status ? mf : mr, 16777216.0f);
        return (double)dur_ticks_disc(mt, x, s.status != 0, a.disc) * 5.9604644775390625e-08;
    }
}

__device__ __forceinline__ void schedule(UnitState &s, double t, int base)
{
    // t > 0: next toggle ceil(t) hours after `base`; residual r = fl(fl(t-(c-1)) - 1)
    const double c = ceil(t);
    s.r = __dadd_rn(__dsub_rn(t, c - 1.0), -1.0);
    s.next = (c >= 1.0e9) ? INT_MAX : base + (int)c;
}

// number of set bits of the loss-of-load bitmap in the hour slots [s0, s1) of a segment
__device__ __forceinline__ unsigned int lol_bits(const uint32_t *bm, int s0, int s1)
{
    unsigned int c = 0;
    const int w0 = s0 >> 5, w1 = (s1 - 1) >> 5;
    for (int w = w0; w <= w1; w++) {
        uint32_t m = bm[w];
        if (w == w0) m &= 0xffffffffu << (s0 & 31);
        if (w == w1 && (s1 & 31)) m &= (1u << (s1 & 31)) - 1u;
        c += __popc(m);
    }
    return c;
}

template <bool kInjected, bool kOneUnit>
__global__ void __launch_bounds__(512, 1) seq_mc_kernel(const SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int Hpad = a.Wd * 32;
    int32_t *s_load = reinterpret_cast<int32_t *>(smem_raw);
    int32_t *s_lmax = s_load + Hpad;
    const int seg_slots = a.seg_words * 32;
    int32_t *tl_all = s_lmax + a.Wd;
    int32_t *tl = tl_all + (size_t)warp * seg_slots;
    uint32_t *bm = reinterpret_cast<uint32_t *>(tl_all + (size_t)wpb * seg_slots) + (size_t)warp * a.seg_words;
    // persistent unit states for the generic path (U > 32 with multi-year chains)
    double *st_r = nullptr; int *st_next = nullptr; uint32_t *st_j = nullptr;
    if (!kOneUnit && a.persist) {
        unsigned char *p = reinterpret_cast<unsigned char *>(tl_all + (size_t)wpb * seg_slots) +
                           sizeof(uint32_t) * (size_t)wpb * a.seg_words;
        p = reinterpret_cast<unsigned char *>(((uintptr_t)p + 7) & ~(uintptr_t)7);
        st_r = reinterpret_cast<double *>(p) + (size_t)warp * a.U;
        st_next = reinterpret_cast<int *>(reinterpret_cast<double *>(p) + (size_t)wpb * a.U) + (size_t)warp * a.U;
        st_j = reinterpret_cast<uint32_t *>(reinterpret_cast<int *>(reinterpret_cast<double *>(p) + (size_t)wpb * a.U) +
                                            (size_t)wpb * a.U) + (size_t)warp * a.U;
    }
    // weak-point statistic (a.imp): per-warp bitmap of the hours with loss of load of the current segment; sits behind
    // the event bitmaps (the host allows it only without persistent unit states)
    const bool want_imp = a.imp != nullptr;
    uint32_t *lolbm = reinterpret_cast<uint32_t *>(tl_all + (size_t)wpb * seg_slots) + (size_t)(wpb + warp) * a.seg_words;
    unsigned long long imp_acc = 0ull;      // kOneUnit: lane = unit

    // stage the load curve and the per-word maxima once per block
    for (int i = threadIdx.x; i < Hpad; i += blockDim.x) s_load[i] = a.load[i];
    for (int i = threadIdx.x; i < a.Wd; i += blockDim.x) s_lmax[i] = a.lmax[i];
    for (int i = lane; i < seg_slots; i += 32) tl[i] = 0;
    for (int i = lane; i < a.seg_words; i += 32) bm[i] = 0u;
    if (want_imp) { for (int i = lane; i < a.seg_words; i += 32) lolbm[i] = 0u; }
    __syncthreads();

    unsigned long long acc_lol = 0, acc_ent = 0, acc_ywl = 0, acc_lol2 = 0, acc_e2lo = 0, acc_e2hi = 0;
    long long acc_ens = 0;
    unsigned int n_events = 0;

    const long long gw = (long long)blockIdx.x * wpb + warp;
    const long long nw = (long long)gridDim.x * wpb;

    // one-unit path: per-lane constants
    int capu = 0; float mf = 1.f, mr = 1.f; uint32_t thr = 0;
    if (kOneUnit && lane < a.U) {
        capu = a.cap[lane]; mf = a.mttf[lane]; mr = a.mttr[lane]; thr = a.for_thr[lane];
    }

    for (long long cl = gw; cl < a.nchains; cl += nw) {
        const long long chain = a.chain_base + cl;
        UnitState st;
        st.r = 0.0; st.next = INT_MAX; st.j = 0; st.status = 1;
        st.buf[0] = st.buf[1] = st.buf[2] = st.buf[3] = 0;
        int capacity = 0;   // capacity after the last evaluated hour (warp-uniform)

        for (int y = 0; y < a.ypc; y++) {
            unsigned int lolh = 0, entries = 0;
            long long ens_lane = 0;
            for (int seg = 0; seg < a.nseg; seg++) {
                const int seg_h0 = seg * seg_slots;
                const int seg_h1 = min(a.H, seg_h0 + seg_slots);
                const int abs0 = y * a.H + seg_h0, abs1 = y * a.H + seg_h1;
                const bool chain_start = (y == 0 && seg == 0);
                int cap_part = 0;
                const UnitState st0 = st;        // state at the segment start (replayed for the weak-point statistic)
                bool seg_lol = false;

                // ---------------- 1. event generation ----------------
                for (int u = lane; u < (kOneUnit ? 32 : a.U); u += 32) {
                    if (kOneUnit && u >= a.U) break;
                    if constexpr (!kOneUnit) {
                        capu = a.cap[u]; mf = a.mttf[u]; mr = a.mttr[u]; thr = a.for_thr[u];
                    }
                    if (chain_start) {
                        st.j = 0; st.status = 1;
                        if constexpr (!kInjected) {
                            const uint32_t x0 = next_word<false>(a, chain, u, st);
                            if (a.init_mode == PSRA_INIT_STATIONARY && x0 < thr) st.status = 0;
                        }
                        const double d0 = next_duration<kInjected>(a, cl, chain, u, st, mf, mr);
                        schedule(st, (!kInjected && a.disc) ? d0 + 1.0 : d0, -1);   // MATLAB: DOWN from hour d+1
                        cap_part += st.status ? capu : 0;
                    } else if constexpr (!kOneUnit) {
                        st.r = st_r[u]; st.next = st_next[u];
                        const uint32_t jj = st_j[u];
                        st.status = (int)(jj >> 31); st.j = jj & 0x7fffffffu;
                        if constexpr (!kInjected) {
                            const uint32_t jm = st.j & 3u;
                            if (jm) {   // re-create the partially consumed Philox block
                                philox4x32_10((uint32_t)chain, (uint32_t)((unsigned long long)chain >> 32),
                                              (uint32_t)u, st.j >> 2, a.k0, a.k1, st.buf);
                                for (uint32_t q = 0; q < jm; q++) {
                                    st.buf[0] = st.buf[1]; st.buf[1] = st.buf[2]; st.buf[2] = st.buf[3];
                                }
                            }
                        }
                    }
                    while (st.next < abs1) {
                        const int slot = st.next - abs0;
                        int delta = 0;
                        double t = st.r;
                        do {   // PSA.jl:240-248: toggle until the residual is positive again
                            st.status ^= 1;
                            delta += st.status ? capu : -capu;
                            t = __dadd_rn(t, next_duration<kInjected>(a, cl, chain, u, st, mf, mr));
                            n_events++;
                        } while (t <= 0.0);
                        atomicAdd(&tl[slot], delta);
                        atomicOr(&bm[slot >> 5], 1u << (slot & 31));
                        schedule(st, t, st.next);
                    }
                    if constexpr (!kOneUnit) {
                        if (a.persist) {
                            st_r[u] = st.r; st_next[u] = st.next;
                            st_j[u] = st.j | ((uint32_t)st.status << 31);
                        }
                    }
                }
                if (chain_start) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) cap_part += __shfl_xor_sync(0xffffffffu, cap_part, d);
                    capacity = cap_part;
                }
                __syncwarp();

                // ---------------- 2. evaluation of the segment ----------------
                const int nwords = (seg_h1 - seg_h0 + 31) >> 5;
                for (int base = 0; base < nwords; base += 32) {
                    const int w = base + lane;
                    const bool valid = w < nwords;
                    const uint32_t m = valid ? bm[w] : 0u;
                    int s = 0, neg = 0;
                    for (uint32_t mm = m; mm; mm &= mm - 1) {
                        const int d = tl[w * 32 + (__ffs(mm) - 1)];
                        s += d;
                        neg += min(d, 0);
                    }
                    const int incl = warp_incl_scan(s, lane);
                    const int cs = capacity + incl - s;          // capacity entering word w
                    const int wy = seg * a.seg_words + w;        // word index within the year
                    const bool flagged = valid && (cs + neg < s_lmax[min(wy, a.Wd - 1)]);
                    uint32_t fm = __ballot_sync(0xffffffffu, flagged);
                    while (fm) {   // rare: resolve the word hour by hour, lane = hour
                        const int src = __ffs(fm) - 1;
                        fm &= fm - 1;
                        const int wq = base + src;
                        const int csq = __shfl_sync(0xffffffffu, cs, src);
                        const int c = csq + warp_incl_scan(tl[wq * 32 + lane], lane);
                        const int hy0 = seg_h0 + wq * 32;        // year-relative hour of lane 0
                        const int L = s_load[hy0 + lane];
                        const bool lol = c < L;                  // PSA.jl:253 strict
                        const uint32_t mask = __ballot_sync(0xffffffffu, lol);
                        if (mask) {
                            if (want_imp) { if (lane == 0) lolbm[wq] = mask; seg_lol = true; }
                            const uint32_t prev = (hy0 > 0 && csq < s_load[hy0 - 1]) ? 1u : 0u;
                            lolh += __popc(mask);
                            entries += __popc(mask & ~((mask << 1) | prev));   // calnlc.m:22-34
                            if (lol) {
                                ens_lane += (long long)(L - c);
                                if (a.fail) atomicAdd(&a.fail[hy0 + lane], 1u);
                            }
                        }
                    }
                    __syncwarp();   // every lane has read the words it helped to resolve before they are cleared
                    for (uint32_t mm = m; mm; mm &= mm - 1) tl[w * 32 + (__ffs(mm) - 1)] = 0;
                    if (valid) bm[w] = 0u;
                    capacity += __shfl_sync(0xffffffffu, incl, 31);
                }
                __syncwarp();

                // ---------------- 2b. weak points (Montecarlo_seq/seqMain.m:140-150): hours with loss of load in which
                //                  a unit is DOWN.  Rare (the segment must contain loss of load): every lane replays its
                //                  units' events over the segment from the state at its start and counts the loss hours
                //                  inside the DOWN stretches.
                if (want_imp && seg_lol) {
                    for (int u = lane; u < (kOneUnit ? 32 : a.U); u += 32) {
                        if (kOneUnit && u >= a.U) break;
                        float mf_r = mf, mr_r = mr; uint32_t thr_r = thr;
                        if constexpr (!kOneUnit) { mf_r = a.mttf[u]; mr_r = a.mttr[u]; thr_r = a.for_thr[u]; }
                        UnitState sr = st0;
                        if (chain_start || !kOneUnit) {      // !kOneUnit: no persistent states here, every segment starts a chain
                            sr.j = 0; sr.status = 1;
                            if constexpr (!kInjected) {
                                const uint32_t x0 = next_word<false>(a, chain, u, sr);
                                if (a.init_mode == PSRA_INIT_STATIONARY && x0 < thr_r) sr.status = 0;
                            }
                            const double d0 = next_duration<kInjected>(a, cl, chain, u, sr, mf_r, mr_r);
                            schedule(sr, (!kInjected && a.disc) ? d0 + 1.0 : d0, -1);
                        }
                        unsigned long long cnt = 0ull;
                        int cur = abs0;                      // first hour of the present constant-state stretch
                        while (true) {
                            const int nx = min(sr.next, abs1);
                            if (!sr.status && nx > cur) cnt += lol_bits(lolbm, cur - abs0, nx - abs0);
                            if (sr.next >= abs1) break;
                            double t = sr.r;
                            do {
                                sr.status ^= 1;
                                t = __dadd_rn(t, next_duration<kInjected>(a, cl, chain, u, sr, mf_r, mr_r));
                            } while (t <= 0.0);
                            cur = sr.next;                   // the hour of a toggle already shows the new state (PSA.jl:249)
                            schedule(sr, t, sr.next);
                        }
                        if (cnt) {
                            if constexpr (kOneUnit) imp_acc += cnt;
                            else atomicAdd(&a.imp[u], cnt);
                        }
                    }
                    __syncwarp();
                    for (int i = lane; i < nwords; i += 32) lolbm[i] = 0u;
                    __syncwarp();
                }
            }

            // ---------------- 3. per-year indices ----------------
            long long ens = 0;
            if (lolh) ens = warp_sum_ll(ens_lane);
            const long long yi = cl * a.ypc + y;
            if (lane == 0) {
                if (a.lol) a.lol[yi] = lolh;
                if (a.ens) a.ens[yi] = ens;
                if (a.ent) a.ent[yi] = entries;
                if (a.group_lol && lolh) atomicAdd(&a.group_lol[(yi + a.group_phase) / a.group], (unsigned long long)lolh);
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

    if (kOneUnit && want_imp && lane < a.U && imp_acc) atomicAdd(&a.imp[lane], imp_acc);
    unsigned long long ev = n_events;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ev += __shfl_xor_sync(0xffffffffu, ev, d);
    if (lane == 0) {
        if (acc_lol) atomicAdd(&a.acc[ACC_LOL], acc_lol);
        if (acc_ens) atomicAdd(&a.acc[ACC_ENS], (unsigned long long)acc_ens);
        if (acc_ent) atomicAdd(&a.acc[ACC_ENT], acc_ent);
        if (acc_ywl) atomicAdd(&a.acc[ACC_YWL], acc_ywl);
        if (acc_lol2) atomicAdd(&a.acc[ACC_LOL2], acc_lol2);
        if (acc_e2lo | acc_e2hi) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], acc_e2lo, acc_e2hi);
        if (ev) atomicAdd(&a.acc[ACC_EVENTS], ev);
    }
}

// ------------------------------------------------------------------------------- host side
#define PSRA_REDO_CAP 4096

// launch geometry of the generic kernel of this file (injected durations, systems the sampler kernels do not take,
// replays of the redo list)
struct GenericGeom { int seg_words, nseg, persist, wpb; size_t smem; };

static GenericGeom generic_geom(const psra_handle *h, int ypc, bool want_imp, int wpb_cfg)
{
    GenericGeom g{};
    const bool one_unit = h->U <= 32;
    g.seg_words = h->Wd;
    if (one_unit) {
        if (h->cfg.seg_hours > 0) g.seg_words = std::max(1, std::min(h->Wd, (h->cfg.seg_hours + 31) / 32));
        else g.seg_words = std::max(1, std::min(h->Wd, (1120 + 31) / 32));
    }
    g.nseg = (h->Wd + g.seg_words - 1) / g.seg_words;
    g.persist = (!one_unit && (ypc > 1 || g.nseg > 1)) ? 1 : 0;
    auto smem_for = [&](int w) -> size_t {
        size_t b = sizeof(int32_t) * ((size_t)h->Wd * 32 + h->Wd);
        b += (size_t)w * (sizeof(int32_t) * (size_t)g.seg_words * 32 + sizeof(uint32_t) * (size_t)g.seg_words);
        if (g.persist) b += 8 + (size_t)w * h->U * (sizeof(double) + sizeof(int) + sizeof(uint32_t));
        if (want_imp) b += sizeof(uint32_t) * (size_t)w * g.seg_words;       // loss-of-load bitmaps
        return b;
    };
    g.wpb = std::max(1, std::min(16, wpb_cfg > 0 ? wpb_cfg : 16));
    while (g.wpb > 1 && smem_for(g.wpb) > h->smem_optin) g.wpb--;
    g.smem = smem_for(g.wpb);
    return g;
}

int psra_seq_prepare_fail(psra_handle *h)
{
    if (h->fail_cap < h->Wd * 32) {
        if (h->d_fail) cudaFree(h->d_fail);
        h->d_fail = nullptr; h->fail_cap = 0;
        PSRA_CUDA(h, cudaMalloc(&h->d_fail, sizeof(uint32_t) * (size_t)h->Wd * 32));
        h->fail_cap = h->Wd * 32;
    }
    PSRA_CUDA(h, cudaMemsetAsync(h->d_fail, 0, sizeof(uint32_t) * (size_t)h->Wd * 32, h->stream));
    return PSRA_OK;
}

static int run_seq(psra_handle *h, bool injected, const double *h_dur, int K, long long chain_base,
                   long long nchains, int ypc, int init_mode, uint64_t seed, const psra_seq_outputs *out,
                   psra_seq_summary *summary, uint64_t *imp_out = nullptr)
{
    PSRA_REQUIRE(h, h->U > 0, "psra_set_system has not been called");
    PSRA_REQUIRE(h, h->H > 0, "psra_set_load has not been called");
    PSRA_REQUIRE(h, summary != nullptr, "summary must not be NULL");
    PSRA_REQUIRE(h, nchains >= 0 && ypc >= 1, "bad year / chain counts");
    PSRA_REQUIRE(h, (long long)ypc * h->H < (1ll << 30), "years_per_chain * hours must stay below 2^30");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    memset(summary, 0, sizeof(*summary));
    const long long nyears = nchains * ypc;
    summary->years = nyears;
    h->kept_n = 0;
    h->hist_years = 0;
    if (nyears == 0) return PSRA_OK;

    const bool one_unit = h->U <= 32;
    const bool load16 = h->max_load <= 32767;
    // the weak-point statistic (imp_out) lives in the generic kernel of this file only
    const bool fast = !injected && one_unit && (long long)ypc * h->H < (1ll << 26) && !h->cfg.reserved[0] && !imp_out;
    const bool team = !injected && !one_unit && h->U <= seq_team_max_units() && (long long)ypc * h->H < (1ll << 20) &&
                      !h->cfg.reserved[0] && !imp_out;
    PSRA_REQUIRE(h, !imp_out || one_unit || ypc == 1, "unit importance for more than 32 units needs years_per_chain = 1");
    // seq_wide.cu: one block per year with a lane-level work queue over the units, when the whole year is one
    // shared-memory timeline (independent years, no explicit segment length); seq_team.cu covers the rest
    const bool wide = team && ypc == 1 && h->cfg.seg_hours == 0 && h->Wd * 32 <= 10240 && !h->cfg.reserved[2] &&
                      h->U <= SEQ_WIDE_MAX_UNITS;
    SeqArgs a{};
    a.U = h->U; a.H = h->H; a.Wd = h->Wd; a.ypc = ypc; a.init_mode = init_mode & ~PSRA_DISC_MATLAB; a.K = K;
    a.disc = (!injected && (init_mode & PSRA_DISC_MATLAB)) ? 1 : 0;
    a.order = h->d_order; a.wide_tab = h->d_wide_tab;
    a.cap = h->d_cap; a.mttf = h->d_mttf; a.mttr = h->d_mttr; a.for_thr = h->d_for_thr;
    a.load = h->d_load; a.lmax = h->d_lmax;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { a.rk[2 * r] = a.k0 + (uint32_t)r * 0x9E3779B9u; a.rk[2 * r + 1] = a.k1 + (uint32_t)r * 0xBB67AE85u; }
    a.chain_base = chain_base; a.nchains = nchains;
    a.acc = h->d_acc;
    a.load16 = load16 ? 1 : 0;
    a.redo = h->d_redo; a.redo_cap = PSRA_REDO_CAP;

    // launch geometry: segment length and warps per block under the shared-memory budget
    const GenericGeom gg = generic_geom(h, ypc, imp_out != nullptr, h->cfg.warps_per_block);
    int seg_words = h->Wd;
    if (team) {
        // whole year in one segment when its int32 timeline stays below 40 KB, else 1760-hour segments
        int seg_hours = h->cfg.seg_hours > 0 ? h->cfg.seg_hours : (h->Wd * 32 <= 10240 ? h->Wd * 32 : 1760);
        seg_words = std::max(1, std::min(h->Wd, (seg_hours + 31) / 32));
    }
    // seq_fast.cu: event lists sized 1.75x the expected transitions of a segment (+ initial draws + slack:
    // > 15 standard deviations for RTS-79; a chain whose list overflows goes to the redo list);
    // by default the whole year is one segment unless that list would exceed 2048 entries
    auto ev_cap_for = [&](int sw) -> int {
        if (h->cfg.ev_cap > 0) return (std::max(256, h->cfg.ev_cap) + 31) & ~31;     // cross-checks of the redo path
        const double e = h->events_per_hour * sw * 32.0 + h->U;
        const long long c = std::max(512ll, (long long)(1.75 * e) + 64);     // >= 1 static block + a few waves of 128 slots
        return (int)((c + 31) & ~31ll);
    };
    if (one_unit) {
        if (!fast) seg_words = gg.seg_words;
        else if (h->cfg.seg_hours > 0) seg_words = std::max(1, std::min(h->Wd, (h->cfg.seg_hours + 31) / 32));
        // event entries hold an 11-bit list link and a 14-bit hour: bound the list and the segment length
        if (fast) { while (seg_words > 1 && (ev_cap_for(seg_words) > 2016 || seg_words > 512)) seg_words = (seg_words + 1) / 2; }
    }
    a.seg_words = seg_words;
    a.nseg = (h->Wd + seg_words - 1) / seg_words;
    a.persist = (!one_unit && !team) ? gg.persist : 0;
    a.pend_cap = seq_team_pend_cap(h->U);
    a.two_halves = (a.nseg > 1 || ypc > 1) ? 1 : 0;
    a.ev_cap = ev_cap_for(seg_words);
    // seq_fast.cu single-segment mode: the word's sum and negative sum share one int32 (sum + 2^K * neg) when
    // |sum| <= installed capacity < 2^(K-1) and neg > -2^(31-K).  A list of ev_cap entries holds at most
    // (ev_cap + U) / 2 + 1 down events (the events of a unit alternate), so the bound below makes a silent
    // overflow impossible: the list overflow (redo list) would trigger first.
    a.pack_shift = 0;
    if (fast && !a.two_halves && h->cfg.reserved[1] != 1) {
        int K = 2;
        while ((1ll << (K - 1)) <= h->total_cap) K++;
        const long long worst = ((long long)(a.ev_cap + h->U) / 2 + 1) * (long long)h->max_unit_cap;
        if (K <= 20 && worst < (1ll << (31 - K))) a.pack_shift = K;
    }
    if (fast && a.ev_cap > 2016)
        return psra_fail(h, PSRA_E_INVALID, "unit transition rate too high for the sampler kernel (%d events per 32-hour word)", a.ev_cap);
    // seq_wide.cu: Philox blocks per unit that are generated without any scheduling (static phase), one count per group
    // of 32 queue positions (the units are sorted by transition rate, so a group's units have about the same demand).
    // A unit with m = 2 H / (MTTF + MTTR) expected transitions needs ceil((m' + 2) / 4) blocks, m' ~ m +- sqrt(m); a
    // static block costs ~140 warp-instructions per 32 units, a queued one ~200, so block b is worth scheduling
    // statically when it is needed with probability >= ~0.7: B = floor((m + 2 - 0.5 sqrt(m)) / 4), at least 1
    // (block 0 holds the initial state).  psra_config.reserved[3] = k > 0 moves the 0.5 to (k - 16) / 8 (sweeps).
    if (wide) {
        const double theta = h->cfg.reserved[3] > 0 ? (h->cfg.reserved[3] - 16) / 8.0 : 0.5;
        const int ngroups = (h->U + 31) / 32;
        for (int g = 0; g < ngroups; g++) {
            int B = 255;
            for (int k = g * 32; k < std::min(h->U, g * 32 + 32); k++) {
                const double m = 2.0 * h->H / h->cycle_sorted[(size_t)k];
                B = std::min(B, (int)std::floor((m + 2.0 - theta * std::sqrt(m)) / 4.0));
            }
            a.wide_sblk[g] = (uint8_t)std::max(1, std::min(255, B));
        }
    }
    // seq_fast.cu single-segment mode: the first blocks of every unit are generated lane = unit without any scheduling.
    // Their number: about 0.7 x the mean demand E[blocks] = (1 + 2 H / (MTTF + MTTR) + 1) / 4 per unit (3 for RTS-79;
    // a simulation of the scheduler puts the optimum of cost = 250 static + 425 per wave there), at least 1
    // (block 0 holds the initial state), and such that the list keeps room for the waves.
    a.static_blocks = 1;
    if (fast && !a.two_halves) {
        const double mean_blocks = (2.0 + h->events_per_hour * h->H / h->U) / 4.0 + 0.5;
        int sb = (int)(0.7 * mean_blocks);
        sb = std::max(1, std::min(sb, a.ev_cap / (8 * h->U)));            // at most half of the list
        if (h->cfg.reserved[3] > 0) sb = std::min(h->cfg.reserved[3], std::max(1, (a.ev_cap - 160) / (4 * h->U)));
        a.static_blocks = sb;
    }
    int wpb = h->cfg.warps_per_block > 0 ? h->cfg.warps_per_block : (fast ? 32 : 16);
    wpb = std::max(1, std::min(fast ? seq_fast_max_threads(a.two_halves != 0) / 32 : 16, wpb));
    auto smem_for = [&](int w) -> size_t {
        if (fast) return seq_fast_smem_bytes(h->Wd, seg_words, w, a.ev_cap, a.two_halves != 0, load16, a.pack_shift != 0);
        if (wide) return seq_wide_smem_bytes(h->Wd, h->U, w);
        if (team) return seq_team_smem_bytes(h->U, h->Wd, seg_words, a.two_halves != 0);
        return gg.smem;
    };
    if (team) {
        wpb = SEQ_TEAM_WARPS;
        if (wide) {
            // block size of seq_wide.cu (scripts/sweep_wide_small.py, round 2): 3 warps up to ~ 300 units (8.8 / 8.6 / 8.3 / 7.5 /
            // 6.7 e7 years/s at 64 / 96 / 128 / 192 / 256 units against 7.5 / 7.4 / 7.1 / 6.6 / 6.2 with 4), 4 warps above
            // (5.1 / 4.4 / 3.3 / 2.6 e7 at 384 / 512 / 768 / 1024 units; 3 and 5 warps are 2-8 % slower there);
            // psra_config.warps_per_block overrides
            const int wmax = seq_wide_max_warps();
            wpb = h->cfg.warps_per_block > 0 ? std::min(wmax, h->cfg.warps_per_block) : std::min(wmax, h->U <= 320 ? 3 : 4);
        }
    }
    if (!fast && !team) wpb = gg.wpb;
    while (fast && wpb > 1 && smem_for(wpb) > h->smem_optin) wpb--;
    const size_t smem = smem_for(wpb);
    if (smem > h->smem_optin)
        return psra_fail(h, PSRA_E_INVALID, "system too large for the shared-memory timeline (%zu B needed, %zu B available)",
                         smem, h->smem_optin);

    // device inputs / outputs
    if (injected) {
        const size_t bytes = sizeof(double) * (size_t)nchains * a.U * (size_t)K;
        int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, bytes);
        if (rc) return rc;
        PSRA_CUDA(h, cudaMemcpyAsync(h->d_scratch, h_dur, bytes, cudaMemcpyHostToDevice, h->stream));
        a.dur = (const double *)h->d_scratch;
    }
    const bool want_vec = out && (out->lol_hours || out->ens_fp || out->entries || out->keep_on_device);
    if (want_vec) {
        int rc = psra_reserve_outputs(h, nyears);
        if (rc) return rc;
        a.lol = h->d_lol; a.ens = (long long *)h->d_ens; a.ent = h->d_ent;
    }
    if (out && out->fail_count) {
        int rc = psra_seq_prepare_fail(h);
        if (rc) return rc;
        a.fail = h->d_fail;
    }
    if (out && out->tail_hist) {
        int rc = psra_tail_hist_prepare(h);
        if (rc) return rc;
        a.hist = (unsigned long long *)h->d_tail_hist; a.hist_bins = h->tail_bins;
    }
    long long ngroups = 0;
    if (out && (out->group_lol || out->history)) {
        PSRA_REQUIRE(h, out->group >= 1, "group must be >= 1");
        a.group = out->group;
        ngroups = (nyears + a.group - 1) / a.group;
        if (h->group_cap < ngroups) {
            if (h->d_group) cudaFree(h->d_group);
            h->d_group = nullptr; h->group_cap = 0;
            PSRA_CUDA(h, cudaMalloc(&h->d_group, sizeof(long long) * (size_t)ngroups));
            h->group_cap = ngroups;
        }
        PSRA_CUDA(h, cudaMemsetAsync(h->d_group, 0, sizeof(long long) * (size_t)ngroups, h->stream));
        a.group_lol = (unsigned long long *)h->d_group;
    } else {
        a.group = 1;
    }
    PSRA_CUDA(h, cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * ACC_COUNT, h->stream));
    PSRA_CUDA(h, cudaMemsetAsync(h->d_redo, 0, sizeof(unsigned long long), h->stream));
    if (imp_out) {
        int rc = psra_reserve(h, &h->d_scratch2, &h->scratch2_cap, sizeof(unsigned long long) * (size_t)h->U);
        if (rc) return rc;
        PSRA_CUDA(h, cudaMemsetAsync(h->d_scratch2, 0, sizeof(unsigned long long) * (size_t)h->U, h->stream));
        a.imp = (unsigned long long *)h->d_scratch2;
    }

    void (*kern)(SeqArgs) = nullptr;
    if (injected) kern = one_unit ? seq_mc_kernel<true, true> : seq_mc_kernel<true, false>;
    else          kern = one_unit ? seq_mc_kernel<false, true> : seq_mc_kernel<false, false>;
    int bps = 0;
    if (fast) {
        PSRA_CUDA(h, seq_fast_prepare(a.disc != 0, a.two_halves != 0, a.pack_shift != 0, smem, wpb * 32, &bps));
    } else if (wide) {
        PSRA_CUDA(h, seq_wide_prepare(a.disc != 0, smem, wpb * 32, &bps));
    } else if (team) {
        PSRA_CUDA(h, seq_team_prepare(smem, &bps));
    } else {
        PSRA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PSRA_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, wpb * 32, smem));
    }
    if (bps < 1) return psra_fail(h, PSRA_E_CUDA, "sequential kernel does not fit on an SM (smem %zu B)", smem);
    if (h->cfg.blocks_per_sm > 0) bps = std::min(bps, h->cfg.blocks_per_sm);
    long long grid = (long long)h->sm_count * bps;
    const long long need = team ? nchains : (nchains + wpb - 1) / wpb;
    if (grid > need) grid = need;

    // Launch plan.  Normally one launch.  When the convergence history of a long run is wanted, the years are cut into
    // up to PSRA_MAX_CHUNKS launches whose group ranges are whole history-scan blocks: the scan and the read-back
    // of a finished range run on stream2 while the next launch computes (the history is 8 B per `group` years --
    // 8 MB for 10^7 RTS-79 years -- and would otherwise be a serial tail of the call).
    const bool defer = h->multi_defer;     // device of a multi-GPU call: history, failure counts and histogram stay here
    const int64_t nfull = (out && out->history && !defer) ? nyears / a.group : 0;
    int nlaunch = 1;
    long long chunk_chains = nchains;
    int64_t hblock = 0, hblocks_per_launch = 0;
    if (nfull > 0) {
        int rc = psra_history_prepare(h, nfull);
        if (rc) return rc;
        hblock = psra_history_block(h, nfull);
        const int64_t hblocks = (nfull + hblock - 1) / hblock;
        // a launch must cover whole scan blocks: hblock groups = hblock * group years = whole chains
        if (!injected && nyears >= (1ll << 20) && (hblock * a.group) % ypc == 0 && hblocks >= 2 * PSRA_MAX_CHUNKS) {
            hblocks_per_launch = (hblocks + PSRA_MAX_CHUNKS - 1) / PSRA_MAX_CHUNKS;
            chunk_chains = hblocks_per_launch * hblock * a.group / ypc;
            // the last launch also takes the years behind the last full group (nyears % group of them)
            nlaunch = (int)std::min<long long>(PSRA_MAX_CHUNKS, (nchains + chunk_chains - 1) / chunk_chains);
        }
    }
    PSRA_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    // chunked run: the launches alternate between two streams, so the blocks of launch c + 1 fill the SMs that launch c
    // drains (the launches are independent: disjoint years, atomics on the shared accumulators)
    const bool alternate = nlaunch > 1 && !h->cfg.reserved2[0];
    if (alternate) {
        PSRA_CUDA(h, cudaEventRecord(h->ev_join, h->stream));               // the set-up (memsets, uploads) is on `stream`
        PSRA_CUDA(h, cudaStreamWaitEvent(h->stream3, h->ev_join, 0));
    }
    for (int c = 0; c < nlaunch; c++) {
        SeqArgs b = a;
        const long long c0 = (long long)c * chunk_chains;
        b.chain_base = a.chain_base + c0;
        b.nchains = (c == nlaunch - 1) ? nchains - c0 : chunk_chains;
        const long long y0 = c0 * ypc;
        if (a.lol) b.lol = a.lol + y0;
        if (a.ens) b.ens = a.ens + y0;
        if (a.ent) b.ent = a.ent + y0;
        if (a.group_lol) b.group_lol = a.group_lol + y0 / a.group;      // y0 is a multiple of the group when nlaunch > 1
        long long g = grid;
        const long long need_c = (team ? b.nchains : (b.nchains + wpb - 1) / wpb);
        if (g > need_c) g = need_c;
        cudaStream_t ls = (alternate && (c & 1)) ? h->stream3 : h->stream;
        if (fast) seq_fast_launch(b, (unsigned)g, wpb * 32, smem, ls);
        else if (wide) seq_wide_launch(b, (unsigned)g, wpb * 32, smem, ls);
        else if (team) seq_team_launch(b, (unsigned)g, smem, ls);
        else kern<<<(unsigned)g, wpb * 32, smem, ls>>>(b);
        PSRA_CUDA(h, cudaGetLastError());
        if (nlaunch > 1) PSRA_CUDA(h, cudaEventRecord(h->ev_chunk[c], ls));
    }
    if (alternate) {
        PSRA_CUDA(h, cudaEventRecord(h->ev_join, h->stream3));
        PSRA_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    PSRA_CUDA(h, cudaEventRecord(h->ev1, h->stream));

    unsigned long long acc[ACC_COUNT];
    unsigned long long n_redo = 0;
    if (nlaunch > 1) {
        // scan + read back the history range of every launch as soon as that launch has finished
        for (int c = 0; c < nlaunch; c++) {
            PSRA_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_chunk[c], 0));
            int rc = psra_history_range(h, h->d_group, nfull, a.group, c * hblocks_per_launch,
                                        c == nlaunch - 1 ? INT64_MAX / 2 : (c + 1) * hblocks_per_launch, out->history, h->stream2);
            if (rc) return rc;
        }
        // host side of the read-back: copy every staged range into the caller's buffer as soon as it has arrived, while the
        // later launches still compute
        int rc = psra_history_drain(h);
        if (rc) return rc;
    }
    // chains the sampler kernels handed back (event list of seq_fast.cu full):
    // replay each with a kernel that has no such limit, into the same accumulators and output slots.  Rare by
    // construction (see ev_cap_for above), so one small launch per chain is fine.
    PSRA_CUDA(h, cudaMemcpyAsync(&n_redo, h->d_redo, sizeof(n_redo), cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n_redo > PSRA_REDO_CAP)
        return psra_fail(h, PSRA_E_OVERFLOW, "%llu chains overflowed the fast sequential kernel (redo list holds %d): "
                         "set a shorter psra_config.seg_hours", n_redo, PSRA_REDO_CAP);
    if (n_redo > 0) {
        std::vector<unsigned long long> list((size_t)n_redo);
        PSRA_CUDA(h, cudaMemcpy(list.data(), h->d_redo + 1, sizeof(unsigned long long) * (size_t)n_redo, cudaMemcpyDeviceToHost));
        SeqArgs b = a;
        b.redo = nullptr;
        size_t smem_r = 0; int threads_r = 0;
        if (wide) {
            smem_r = smem; threads_r = wpb * 32;
        } else {
            b.seg_words = gg.seg_words; b.nseg = gg.nseg; b.persist = gg.persist;
            smem_r = gg.smem; threads_r = 32;
            const GenericGeom g1 = generic_geom(h, ypc, false, 1);
            smem_r = g1.smem;
            PSRA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
        }
        for (unsigned long long i = 0; i < n_redo; i++) {
            const long long cl = (long long)list[(size_t)i] - a.chain_base;
            if (cl < 0 || cl >= nchains) return psra_fail(h, PSRA_E_CUDA, "internal error: redo list entry out of range");
            SeqArgs r = b;
            r.chain_base = a.chain_base + cl; r.nchains = 1;
            const long long y0 = cl * ypc;
            if (a.lol) r.lol = a.lol + y0;
            if (a.ens) r.ens = a.ens + y0;
            if (a.ent) r.ent = a.ent + y0;
            if (a.group_lol) r.group_lol = a.group_lol + y0 / a.group;
            // group_lol is indexed by (local year) / group: keep the phase of the year inside its group
            r.group_phase = (int)(y0 % a.group);
            if (wide) seq_wide_launch(r, 1u, threads_r, smem_r, h->stream);
            else kern<<<1u, threads_r, smem_r, h->stream>>>(r);
            PSRA_CUDA(h, cudaGetLastError());
        }
        if (nfull > 0 && nlaunch > 1) {      // the overlapped history scan ran before the replays: scan again
            PSRA_CUDA(h, cudaStreamSynchronize(h->stream2));
            int rc = psra_history_range(h, h->d_group, nfull, a.group, 0, INT64_MAX / 2, out->history, h->stream);
            if (rc) return rc;
        }
    }
    PSRA_CUDA(h, cudaMemcpyAsync(acc, h->d_acc, sizeof(acc), cudaMemcpyDeviceToHost, h->stream));
    if (imp_out) PSRA_CUDA(h, cudaMemcpyAsync(imp_out, a.imp, sizeof(uint64_t) * (size_t)h->U, cudaMemcpyDeviceToHost, h->stream));
    if (out) {
        if (out->lol_hours) PSRA_CUDA(h, cudaMemcpyAsync(out->lol_hours, h->d_lol, sizeof(uint32_t) * (size_t)nyears, cudaMemcpyDeviceToHost, h->stream));
        if (out->ens_fp)    PSRA_CUDA(h, cudaMemcpyAsync(out->ens_fp, h->d_ens, sizeof(int64_t) * (size_t)nyears, cudaMemcpyDeviceToHost, h->stream));
        if (out->entries)   PSRA_CUDA(h, cudaMemcpyAsync(out->entries, h->d_ent, sizeof(uint32_t) * (size_t)nyears, cudaMemcpyDeviceToHost, h->stream));
        if (out->fail_count && !defer) PSRA_CUDA(h, cudaMemcpyAsync(out->fail_count, h->d_fail, sizeof(uint32_t) * (size_t)h->H, cudaMemcpyDeviceToHost, h->stream));
        if (out->group_lol) PSRA_CUDA(h, cudaMemcpyAsync(out->group_lol, h->d_group, sizeof(long long) * (size_t)ngroups, cudaMemcpyDeviceToHost, h->stream));
        if (nfull > 0 && nlaunch == 1) {
            int rc = psra_history_range(h, h->d_group, nfull, a.group, 0, INT64_MAX / 2, out->history, h->stream);
            if (rc) return rc;
        }
    }
    if (nlaunch > 1) PSRA_CUDA(h, cudaStreamSynchronize(h->stream2));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    {
        int rc = psra_history_drain(h);      // ranges staged after the launches (single launch, re-scan after replays)
        if (rc) return rc;
    }
    float ms = 0.f;
    PSRA_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    summary->kernel_ms = ms;
    summary->sum_lol_hours = (int64_t)acc[ACC_LOL];
    summary->sum_ens_fp = (int64_t)acc[ACC_ENS];
    summary->sum_entries = (int64_t)acc[ACC_ENT];
    summary->years_with_loss = (int64_t)acc[ACC_YWL];
    summary->sum_lol_sq = acc[ACC_LOL2];
    summary->sum_ens_sq_lo = acc[ACC_ENS2_LO];
    summary->sum_ens_sq_hi = acc[ACC_ENS2_HI];
    summary->events = acc[ACC_EVENTS];
    summary->redone = (int32_t)n_redo;
    memcpy(h->last_acc, acc, sizeof(acc));
    h->last_acc[ACC_REDO] = n_redo;
    if (want_vec && out->keep_on_device) h->kept_n = nyears;
    if (a.hist) { h->hist_years = nyears; h->hist_years_with_loss = (int64_t)acc[ACC_YWL]; }
    if (acc[ACC_OVERFLOW] == 3ull) {
        // seq_fast.cu, ring variant (several segments per year or multi-year chains): a segment's event list was full.
        // Never fail for that (the reference loop cannot, PSA.jl:230-266): repeat the call with the generic kernel.
        if (fast && !h->cfg.reserved[0]) {
            h->cfg.reserved[0] = 1;
            const int rc = run_seq(h, injected, h_dur, K, chain_base, nchains, ypc, init_mode, seed, out, summary, imp_out);
            h->cfg.reserved[0] = 0;
            if (rc == PSRA_OK) { summary->redone = (int32_t)std::min<long long>(nchains, 0x7fffffff); h->last_acc[ACC_REDO] = (unsigned long long)nchains; }
            return rc;
        }
        return psra_fail(h, PSRA_E_OVERFLOW, "event list of a timeline segment overflowed (%d entries): set a shorter psra_config.seg_hours", a.ev_cap);
    }
    if (acc[ACC_OVERFLOW] == 2ull)
        return psra_fail(h, PSRA_E_OVERFLOW, "internal error: pending-event list overflow in the sequential kernel");
    if (acc[ACC_OVERFLOW])
        return psra_fail(h, PSRA_E_OVERFLOW, "injected durations exhausted: a unit needed more than K=%d draws", K);
    return PSRA_OK;
}

int psra_run_seq_range(psra_handle *h, long long chain_base, long long nchains, int ypc, int init_mode, uint64_t seed,
                       const psra_seq_outputs *out, psra_seq_summary *summary)
{
    return run_seq(h, false, nullptr, 0, chain_base, nchains, ypc, init_mode, seed, out, summary);
}

extern "C" int psra_seq_mc(psra_handle *h, int64_t year0, int64_t nyears, uint64_t seed, int32_t init_mode,
                           int32_t years_per_chain, const psra_seq_outputs *out, psra_seq_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, years_per_chain >= 1, "years_per_chain must be >= 1");
    PSRA_REQUIRE(h, year0 >= 0 && nyears >= 0, "negative year range");
    PSRA_REQUIRE(h, year0 % years_per_chain == 0 && nyears % years_per_chain == 0,
                 "year0 and nyears must be multiples of years_per_chain");
    PSRA_REQUIRE(h, (init_mode & ~PSRA_DISC_MATLAB) == PSRA_INIT_ALL_UP || (init_mode & ~PSRA_DISC_MATLAB) == PSRA_INIT_STATIONARY,
                 "unknown init_mode");
    if (!h->peers.empty()) return psra_multi_seq_mc(h, year0, nyears, seed, init_mode, years_per_chain, out, summary);
    return run_seq(h, false, nullptr, 0, year0 / years_per_chain, nyears / years_per_chain, years_per_chain,
                   init_mode, seed, out, summary);
}

extern "C" int psra_seq_eval_injected(psra_handle *h, const double *durations, int64_t nchains,
                                      int32_t years_per_chain, int32_t K, const psra_seq_outputs *out,
                                      psra_seq_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, durations != nullptr && K >= 1 && nchains >= 0, "bad injected duration matrix");
    PSRA_REQUIRE(h, years_per_chain >= 1, "years_per_chain must be >= 1");
    PSRA_REQUIRE(h, h->U > 0, "psra_set_system has not been called");
    const size_t n = (size_t)nchains * h->U * (size_t)K;
    for (size_t i = 0; i < n; i++)
        if (!(durations[i] > 0.0)) return psra_fail(h, PSRA_E_INVALID, "injected durations must be > 0 (entry %zu)", i);
    return run_seq(h, true, durations, K, 0, nchains, years_per_chain, PSRA_INIT_ALL_UP, 0, out, summary);
}

extern "C" int psra_seq_unit_importance(psra_handle *h, int64_t year0, int64_t nyears, uint64_t seed, int32_t init_mode,
                                        int32_t years_per_chain, uint64_t *down_in_loss, const psra_seq_outputs *out,
                                        psra_seq_summary *summary)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, down_in_loss != nullptr, "down_in_loss must not be NULL");
    PSRA_REQUIRE(h, years_per_chain >= 1, "years_per_chain must be >= 1");
    PSRA_REQUIRE(h, year0 >= 0 && nyears >= 0, "negative year range");
    PSRA_REQUIRE(h, year0 % years_per_chain == 0 && nyears % years_per_chain == 0,
                 "year0 and nyears must be multiples of years_per_chain");
    PSRA_REQUIRE(h, (init_mode & ~PSRA_DISC_MATLAB) == PSRA_INIT_ALL_UP || (init_mode & ~PSRA_DISC_MATLAB) == PSRA_INIT_STATIONARY,
                 "unknown init_mode");
    PSRA_REQUIRE(h, h->U > 0, "psra_set_system has not been called");
    for (int u = 0; u < h->U; u++) down_in_loss[u] = 0;
    return run_seq(h, false, nullptr, 0, year0 / years_per_chain, nyears / years_per_chain, years_per_chain,
                   init_mode, seed, out, summary, down_in_loss);
}

// ------------------------------------------------------------------------------- sampler diagnostic
__global__ void sampler_durations_kernel(float mean_ticks, const uint32_t *__restrict__ draws, long long n,
                                         unsigned long long *__restrict__ ticks, uint32_t *__restrict__ e_bits)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t x = draws[i];
        if (e_bits) e_bits[i] = __float_as_uint(neglog_u32(x));
        if (ticks) ticks[i] = dur_ticks(mean_ticks, x);
    }
}

extern "C" int psra_sampler_durations(psra_handle *h, float mean_h, const uint32_t *draws, int64_t n, uint64_t *ticks, uint32_t *e_bits)
{
    if (!h) return PSRA_E_INVALID;
    PSRA_REQUIRE(h, draws && n >= 1 && (ticks || e_bits), "null argument / empty input");
    PSRA_REQUIRE(h, mean_h > 0.0f && mean_h <= (float)PSRA_MAX_MEAN_HOURS, "mean duration must be positive (at most 1e8 hours)");
    PSRA_CUDA(h, cudaSetDevice(h->device));
    int rc = psra_reserve(h, &h->d_scratch, &h->scratch_cap, (size_t)n * (sizeof(uint32_t) * 2 + sizeof(uint64_t)));
    if (rc) return rc;
    unsigned long long *d_t = (unsigned long long *)h->d_scratch;
    uint32_t *d_x = (uint32_t *)(d_t + n), *d_e = d_x + n;
    PSRA_CUDA(h, cudaMemcpyAsync(d_x, draws, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 8);
    sampler_durations_kernel<<<grid, 256, 0, h->stream>>>(mean_h * 16777216.0f, d_x, n, ticks ? d_t : nullptr, e_bits ? d_e : nullptr);
    PSRA_CUDA(h, cudaGetLastError());
    if (ticks) PSRA_CUDA(h, cudaMemcpyAsync(ticks, d_t, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    if (e_bits) PSRA_CUDA(h, cudaMemcpyAsync(e_bits, d_e, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    PSRA_CUDA(h, cudaStreamSynchronize(h->stream));
    return PSRA_OK;
}

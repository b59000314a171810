// seq_wide.cu -- sampler-driven sequential chronological MC for systems of more than 32 units when
// the whole year fits one shared-memory timeline (BASELINE config 5: 1024 units, 8736 hours).
// Same model, sampler and per-year integers as seq_fast.cu / seq_team.cu / seq_mc.cu
// (run_sequential_mc, GeneratingAdequacy/PowerSystemAdequacy.jl:214-269; indices per
// Montecarlo_seq/seqMain.m:160-176, Montecarlo_seq/calnlc.m:22-34).
//
// One thread block owns one simulated year at a time.  With hundreds of units there is no need to
// spread the Philox blocks of one unit over several lanes (the wave scheduler of seq_fast.cu /
// seq_team.cu): here a lane owns one unit and walks its stream block by block -- Philox4x32-10,
// four tick durations, a running 64-bit event time in registers, four integer-MW deltas added to the
// block's shared-memory hour timeline -- until the unit has passed the end of the year.  Then the
// lane takes the next unit from a block-wide work queue (shared-memory counter), so lanes stay busy
// although the units' event rates differ by 6x; the queue hands the units out in the host-sorted
// order "most transitions first" (longest job first), which keeps the tail of the year short.
// Out-of-year events are not branched around: they add into a per-lane dummy slot behind the year.
//
// The lane state is kept minimal (round 2): the running time is T - 1 of the last event, so the hour of
// an event is one funnel shift of the running sum and "the unit has left the year" is the in-year test
// of its fourth event; the initial-state draw (word 0 of block 0) is a duration of zero whose event
// falls on hour 2^32 - 1; a lane without a unit is parked at 2^55 ticks and needs no special casing.
//
// kPack: the hour timeline keeps two hours per 32-bit word (two's-complement halves: hour 2i in the low
// half, 2i + 1 in the high half; an event adds delta or delta << 16 with one 32-bit atomic, so the word ends
// as sum_lo + 2^16 sum_hi mod 2^32 whatever the order of the adds).  Half the shared memory per year =
// 8 instead of 5 resident blocks per SM.  The halves decode correctly as long as the net capacity
// change of every single hour stays inside int16.  Guard: a year in which a half leaves [-2^14, 2^14) is
// handed back to the host, which replays it with the int32 timeline (seq_mc.cu run_seq, "redo list"); so is
// a year whose checksum fails -- the capacity at the end of the year from the timeline must equal the
// capacity of the units the generators left UP (a half that wrapped past +-2^15 and came back into the
// accepted range shifts the sum by 65535).  The host only picks this variant when 2^14 is at least
// 8 + 4 x (events per hour) times the largest unit.
//
// Evaluation: the warps reduce the hour deltas to per-32-hour-word sums (lane = word, skewed so that the
// 32 lanes hit 32 different banks); after a barrier every warp scans the word sums into the capacity
// entering each run of words (redundantly: 9 loads per lane), flags the runs that can contain loss of
// load (capacity + negative hour deltas < maximum load of a word) and resolves / clears the runs it
// owns: flagged words hour by hour (shuffle scan, __ballot_sync / __popc for LOL hours and deficit
// entries, int64 ENS per lane).  The per-year sums meet in shared-memory scalars that are double
// buffered by year parity, so warp 0 writes year y out while the other warps already generate y + 1.
#include <limits.h>

#include "psra_internal.cuh"
#include "seq_args.cuh"

#define WIDE_THREADS 128
#define WIDE_BLOCKS_PER_SM(pack) ((pack) ? 8 : 5)

struct WideShared {     // one per year parity; 32 bytes (the queue-head address is computed by hand below)
    int capacity;       // sum of the capacities of the units that start the year UP
    int cap_end;        // ... of the units the generators left UP at the end of the year (checksum)
    int queue_head;     // next position of the unit order that has not been handed out
    int bad;            // the checksum failed: the year goes to the host's redo list
    unsigned int lolh, entries;
    unsigned long long ens;
};

static_assert(sizeof(WideShared) == 32, "the generation loop addresses queue_head of year parity p at + 32 p");

size_t seq_wide_smem_bytes(int Wd, bool pack)
{
    size_t b = sizeof(int32_t) * ((size_t)Wd * (pack ? 16 : 32) + 32);   // hour timeline + one dummy slot per lane
    b += 3 * sizeof(int32_t) * (size_t)((Wd + 3) & ~3);                   // word sums, negative sums, word maxima of the load
    b += 2 * sizeof(WideShared) + 32 * sizeof(int32_t) + 16;             // + one always-zero word per lane
    return (b + 15) & ~(size_t)15;
}

// add `delta` to the hour slot of an event, or to the lane's dummy slot when the event lies beyond the year
template <bool kPack>
__device__ __forceinline__ void wide_scatter(uint32_t tl_s, uint32_t dummy_s, uint32_t hs, uint32_t H, int delta, unsigned int &n_events)
{
    if constexpr (!kPack) {
        asm volatile("{\n .reg .pred p;\n .reg .b32 ad;\n"
                     " setp.lt.u32 p, %1, %2;\n"
                     " mad.lo.u32 ad, %1, 4, %3;\n"
                     " selp.b32 ad, ad, %4, p;\n"
                     " red.shared.add.s32 [ad], %5;\n"
                     " @p add.u32 %0, %0, 1;\n}\n"
                     : "+r"(n_events)
                     : "r"(hs), "r"(H), "r"(tl_s), "r"(dummy_s), "r"(delta)
                     : "memory");
    } else {
        // word hs >> 1; delta << 16 for an odd hour: the funnel shift in wrap mode takes its count mod 32, and
        // (16 hs) mod 32 = 16 (hs & 1)
        asm volatile("{\n .reg .pred p;\n .reg .b32 ad, w, s, d;\n"
                     " setp.lt.u32 p, %1, %2;\n"
                     " shr.u32 w, %1, 1;\n"
                     " mad.lo.u32 ad, w, 4, %3;\n"
                     " selp.b32 ad, ad, %4, p;\n"
                     " shl.b32 s, %1, 4;\n"
                     " shf.l.wrap.b32 d, 0, %5, s;\n"
                     " red.shared.add.s32 [ad], d;\n"
                     " @p add.u32 %0, %0, 1;\n}\n"
                     : "+r"(n_events)
                     : "r"(hs), "r"(H), "r"(tl_s), "r"(dummy_s), "r"(delta)
                     : "memory");
    }
}

// one sampler duration in ticks: RN_int64(max(mean_ticks * E(x), lo)) -- lo = 1 tick, or 0 together with
// mean_ticks = 0 for the initial-state draw, which is not a duration
template <bool kDisc>
__device__ __forceinline__ unsigned long long wide_dur(float mean_ticks, uint32_t x, float lo, bool up_state, uint32_t one_bits)
{
    unsigned long long t = (unsigned long long)__float2ll_rn(fmaxf(__fmul_rn(mean_ticks, neglog_u32(x, one_bits)), lo));
    if constexpr (kDisc) t = ((t + (up_state ? (1ull << 23) : ((1ull << 24) - 1ull))) >> 24) << 24;
    return t;
}

template <bool kDisc, bool kPack>
__global__ void __launch_bounds__(WIDE_THREADS, WIDE_BLOCKS_PER_SM(kPack)) seq_wide_kernel(const SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    constexpr int HPW = kPack ? 16 : 32;                                 // timeline integers per 32-hour word
    const int tl_len = a.Wd * HPW, Wd4 = (a.Wd + 3) & ~3;
    int32_t *tl = reinterpret_cast<int32_t *>(smem_raw);                 // [tl_len + 32]
    int32_t *wsum = tl + tl_len + 32;                                    // [Wd4]
    int32_t *wneg = wsum + Wd4;
    int32_t *s_lmax = wneg + Wd4;
    WideShared *sh_all = reinterpret_cast<WideShared *>(s_lmax + Wd4);   // [2], 8-byte aligned (Wd4 is a multiple of 4)
    const uint32_t tl_s = (uint32_t)__cvta_generic_to_shared(tl);
    uint32_t dummy_s = tl_s + 4u * (uint32_t)(tl_len + lane);
    // keep the two shared addresses in registers: left alone, ptxas rebuilds them (S2R SR_CgaCtaId, LEA, ...) in every
    // iteration of the generation loop; one_bits is the exponent pattern of 1.0f as an opaque register (see neglog_u32)
    uint32_t tl_o = tl_s, one_bits = 0x3F800000u;
    const uint32_t qh_s = (uint32_t)__cvta_generic_to_shared(&sh_all[0].queue_head);   // + sizeof(WideShared) for the odd years
    int32_t *zero32 = reinterpret_cast<int32_t *>(sh_all + 2);                         // [32], stays 0
    const uint32_t zero_s = (uint32_t)__cvta_generic_to_shared(zero32 + lane);
    asm volatile("" : "+r"(tl_o), "+r"(dummy_s), "+r"(one_bits));

    for (int i = threadIdx.x; i < a.Wd; i += blockDim.x) s_lmax[i] = a.lmax[i];
    for (int i = threadIdx.x; i < tl_len + 32; i += blockDim.x) tl[i] = 0;
    if (threadIdx.x < 32) zero32[threadIdx.x] = 0;
    __syncthreads();

    unsigned long long acc_lol = 0, acc_ent = 0, acc_ywl = 0, acc_lol2 = 0, acc_e2lo = 0, acc_e2hi = 0;
    long long acc_ens = 0;
    unsigned long long ev64 = 0ull;
    unsigned int n_jobs = 0, n_flag = 0;
    const uint32_t Hu = (uint32_t)a.H;
    const unsigned long long parked = 1ull << 55;                      // T - 1 of a lane without a unit: far beyond any year
    // T - 1 before the first duration.  MATLAB discretisation: a unit that fails after d whole hours is DOWN from
    // hour d + 1 (seq_mcsampling.m:63)
    const unsigned long long start_m1 = (kDisc ? (1ull << PSRA_TICK_SHIFT) : 0ull) - 1ull;
    const bool stationary = a.init_mode == PSRA_INIT_STATIONARY;

    if (threadIdx.x < 2) {
        WideShared *z = sh_all + threadIdx.x;
        z->capacity = 0; z->cap_end = 0; z->queue_head = (int)blockDim.x; z->bad = 0; z->lolh = 0u; z->entries = 0u; z->ens = 0ull;
    }
    __syncthreads();

    int par = 0;
    for (long long cl = blockIdx.x; cl < a.nchains; cl += gridDim.x, par ^= 1) {
        const unsigned long long chain = (unsigned long long)(a.chain_base + cl);
        WideShared *sh = sh_all + par;

        // ---- generation: lane = unit, block after block; finished lanes pull the next unit from the queue
        int pos = threadIdx.x;              // position in the unit order
        bool busy = pos < a.U;
        int u = 0, cu = 0, cap_up = 0, cap_end = 0;
        float mup = 1.f, mdn = 1.f;
        uint32_t thr = 0u, nb = 1u;
        bool sdn = false;                   // the unit's stream starts DOWN
        unsigned long long tm1 = parked;    // T - 1 of the unit's last event
        unsigned int ne = 0u;               // in-year events of the current unit
        auto take_unit = [&]() {
            const uint4 rec = __ldg(&a.wide_tab[pos]);      // one 16-byte record per queue position
            u = __ldg(&a.order[pos]);
            cu = (int)rec.x;
            mup = __uint_as_float(rec.y);
            mdn = __uint_as_float(rec.z);
            thr = rec.w;
            nb = 0u;
            tm1 = start_m1;
        };
        if (busy) take_unit();
        while (__any_sync(0xffffffffu, busy)) {
            uint32_t x[4];
            philox4x32_10_rk((uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)u, nb, a.rk, x);
            const bool first = nb == 0u;
            if (first) {                    // draw 0 of a stream is the initial state
                sdn = stationary && x[0] < thr;
                if (busy && !sdn) cap_up += cu;
            }
            const float m_a = sdn ? mup : mdn;      // draws 0, 2 of a block: the state opposite to the initial one
            const float m_b = sdn ? mdn : mup;      // draws 1, 3: the initial state
            const unsigned long long t1 = tm1 + wide_dur<kDisc>(first ? 0.f : m_a, x[0], first ? 0.f : 1.f, sdn, one_bits);
            const unsigned long long t2 = t1 + wide_dur<kDisc>(m_b, x[1], 1.f, !sdn, one_bits);
            const unsigned long long t3 = t2 + wide_dur<kDisc>(m_a, x[2], 1.f, sdn, one_bits);
            const unsigned long long t4 = t3 + wide_dur<kDisc>(m_b, x[3], 1.f, !sdn, one_bits);
            // hour of an event at tick T: ceil(T / 2^24) - 1 = (T - 1) >> 24
            uint32_t h1 = __funnelshift_r((uint32_t)t1, (uint32_t)(t1 >> 32), PSRA_TICK_SHIFT);
            const uint32_t h2 = __funnelshift_r((uint32_t)t2, (uint32_t)(t2 >> 32), PSRA_TICK_SHIFT);
            const uint32_t h3 = __funnelshift_r((uint32_t)t3, (uint32_t)(t3 >> 32), PSRA_TICK_SHIFT);
            const uint32_t h4 = __funnelshift_r((uint32_t)t4, (uint32_t)(t4 >> 32), PSRA_TICK_SHIFT);
            if (kDisc && first) h1 = 0xffffffffu;   // otherwise T - 1 = -1 already puts the non-event beyond the year
            const int d_a = sdn ? -cu : cu;         // draws 0, 2 end with the unit back in its initial state
            wide_scatter<kPack>(tl_o, dummy_s, h1, Hu, d_a, ne);
            wide_scatter<kPack>(tl_o, dummy_s, h2, Hu, -d_a, ne);
            wide_scatter<kPack>(tl_o, dummy_s, h3, Hu, d_a, ne);
            wide_scatter<kPack>(tl_o, dummy_s, h4, Hu, -d_a, ne);
            tm1 = t4;
            nb++;
            if (busy && h4 >= Hu) {         // unit done: take the next one of the block's queue
                ev64 += ne;
                n_jobs += nb;
                if (sdn == ((ne & 1u) != 0u)) cap_end += cu;     // UP at the end: started UP and toggled an even number of times, or ...
                ne = 0u;
                // One plain atomic per finishing lane.  Left alone, ptxas turns an atomic add of a constant to a
                // warp-uniform address into a warp-aggregated sequence (vote, leader election, popc, shuffle: ~25
                // instructions) that every iteration of the loop would pay for; the address therefore gets a per-lane
                // offset the compiler cannot see through (a shared-memory word per lane that always holds 0).
                asm volatile("{\n .reg .b32 z;\n ld.volatile.shared.u32 z, [%2];\n add.u32 z, z, %1;\n"
                             " atom.shared.add.u32 %0, [z], 1;\n}\n"
                             : "=r"(pos) : "r"(qh_s + 32u * (uint32_t)par), "r"(zero_s) : "memory");
                busy = pos < a.U;
                if (busy) take_unit();
                else { tm1 = parked; mup = 1.f; mdn = 1.f; nb = 1u; }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cap_up += __shfl_xor_sync(0xffffffffu, cap_up, d);
            cap_end += __shfl_xor_sync(0xffffffffu, cap_end, d);
        }
        if (lane == 0) {
            if (cap_up) atomicAdd(&sh->capacity, cap_up);
            if (cap_end) atomicAdd(&sh->cap_end, cap_end);
        }
        __syncthreads();

        // ---- word sums of the hour deltas: lane = word, index skewed by the lane (conflict-free)
        for (int w0 = warp * 32; w0 < a.Wd; w0 += nwarps * 32) {
            const int w = w0 + lane;
            if (w < a.Wd) {
                const int32_t *row = tl + w * HPW;
                int s = 0, n = 0;
                if constexpr (!kPack) {
#pragma unroll 8
                    for (int j = 0; j < 32; j++) {
                        const int d = row[(j + lane) & 31];
                        s += d;
                        n += min(d, 0);
                    }
                } else {
                    int rng = 0;
#pragma unroll 8
                    for (int j = 0; j < 16; j++) {
                        const int v = row[(j + (lane >> 1)) & 15];
                        const int lo = (int)(short)(v & 0xffff);
                        const int hi = (v - lo) >> 16;
                        s += lo + hi;
                        n += min(lo, 0) + min(hi, 0);
                        rng |= v ^ (v << 1);            // bit 15 / 31: the half is outside [-2^14, 2^14)
                    }
                    if (rng & 0x80008000) sh->bad = 1;  // a half may have wrapped: the year goes to the redo list
                }
                wsum[w] = s; wneg[w] = n;
            }
        }
        __syncthreads();

        // ---- evaluation: every warp scans the word sums (lane = run of `wpl` consecutive words), then resolves and
        //      clears the runs it owns
        {
            unsigned int lolh = 0, entries = 0;
            long long ens_lane = 0;
            const int nwords = a.Wd;
            const int wpl = (nwords + 31) >> 5;
            const int wb = lane * wpl;
            int loc = 0, lmin = INT_MAX;
            for (int k = 0; k < wpl; k++) {
                const int w = wb + k;
                if (w < nwords) {
                    lmin = min(lmin, loc + wneg[w] - s_lmax[w]);
                    loc += wsum[w];
                }
            }
            const int incl = warp_incl_scan(loc, lane);
            const int cap0 = sh->capacity;
            const int cs_lane = cap0 + incl - loc;                  // capacity entering the lane's run
            // checksum: the capacity the timeline ends the year with against the units the generators left UP
            const bool year_ok = (cap0 + __shfl_sync(0xffffffffu, incl, 31)) == sh->cap_end && !(kPack && sh->bad);
            const bool flagged = year_ok && (lmin != INT_MAX) && (cs_lane + lmin < 0);
            const uint32_t fm = __ballot_sync(0xffffffffu, flagged);
            if (warp == 0) {
                n_flag += __popc(fm);
                if (!year_ok && lane == 0) sh->bad = 1;
            }
            for (int src = warp; src < 32; src += nwarps) {         // runs owned by this warp
                if ((fm >> src) & 1u) {                             // rare: the run may contain loss of load
                    int c_in = __shfl_sync(0xffffffffu, cs_lane, src);
                    for (int k = 0; k < wpl; k++) {
                        const int wq = src * wpl + k;
                        if (wq >= nwords) break;
                        if (c_in + wneg[wq] < s_lmax[wq]) {         // resolve the word hour by hour, lane = hour
                            int dl;
                            if constexpr (!kPack) dl = tl[wq * 32 + lane];
                            else {
                                const int v = tl[wq * 16 + (lane >> 1)];
                                const int lo = (int)(short)(v & 0xffff);
                                dl = (lane & 1) ? ((v - lo) >> 16) : lo;
                            }
                            const int c = c_in + warp_incl_scan(dl, lane);
                            const int hy0 = wq * 32;
                            const int L = __ldg(&a.load[hy0 + lane]);
                            const bool lol = c < L;                 // PSA.jl:253 strict
                            const uint32_t mask = __ballot_sync(0xffffffffu, lol);
                            if (mask) {
                                const uint32_t prev = (hy0 > 0 && c_in < __ldg(&a.load[hy0 - 1])) ? 1u : 0u;
                                lolh += __popc(mask);
                                entries += __popc(mask & ~((mask << 1) | prev));   // calnlc.m:22-34
                                if (lol) {
                                    ens_lane += (long long)(L - c);
                                    if (a.fail) atomicAdd(&a.fail[hy0 + lane], 1u);
                                }
                            }
                        }
                        c_in += wsum[wq];
                    }
                }
                // clear the run (wpl 32-hour words; the dummy slots behind the year may keep their garbage)
                int4 *t4 = reinterpret_cast<int4 *>(tl + src * wpl * HPW);
                const int n4 = min(wpl, max(0, nwords - src * wpl)) * (HPW / 4);
                for (int i = lane; i < n4; i += 32) t4[i] = make_int4(0, 0, 0, 0);
            }
            if (lolh) {                                             // uniform within the warp
                const long long ens = warp_sum_ll(ens_lane);
                if (lane == 0) {
                    atomicAdd(&sh->lolh, lolh);
                    atomicAdd(&sh->entries, entries);
                    atomicAdd(&sh->ens, (unsigned long long)ens);
                }
            }
        }
        __syncthreads();

        // ---- per-year indices: warp 0 writes year `cl` out and re-arms its scalars for the year after next while the
        //      other warps already generate the next year with the other set
        if (threadIdx.x == 0) {
            const unsigned int lolh = sh->lolh, entries = sh->entries;
            const long long ens = (long long)sh->ens;
            const bool bad = sh->bad != 0;
            sh->capacity = 0; sh->cap_end = 0; sh->queue_head = (int)blockDim.x; sh->bad = 0; sh->lolh = 0u; sh->entries = 0u; sh->ens = 0ull;
            if (bad) {
                seq_redo_push(a, (long long)chain);                 // the host replays this year (int32 timeline)
            } else {
                if (a.lol) a.lol[cl] = lolh;
                if (a.ens) a.ens[cl] = ens;
                if (a.ent) a.ent[cl] = entries;
                if (a.group_lol && lolh) atomicAdd(&a.group_lol[(cl + a.group_phase) / a.group], (unsigned long long)lolh);
                if (lolh) seq_hist_add(a, ens);
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
    }

    unsigned long long ev = ev64, jb = n_jobs;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ev += __shfl_xor_sync(0xffffffffu, ev, d);
        jb += __shfl_xor_sync(0xffffffffu, jb, d);
    }
    if (lane == 0) {
        if (ev) atomicAdd(&a.acc[ACC_EVENTS], ev);
        atomicAdd(&a.acc[ACC_JOBS], jb);
        if (warp == 0) {
            if (acc_lol) atomicAdd(&a.acc[ACC_LOL], acc_lol);
            if (acc_ens) atomicAdd(&a.acc[ACC_ENS], (unsigned long long)acc_ens);
            if (acc_ent) atomicAdd(&a.acc[ACC_ENT], acc_ent);
            if (acc_ywl) atomicAdd(&a.acc[ACC_YWL], acc_ywl);
            if (acc_lol2) atomicAdd(&a.acc[ACC_LOL2], acc_lol2);
            if (acc_e2lo | acc_e2hi) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], acc_e2lo, acc_e2hi);
            atomicAdd(&a.acc[ACC_FLAGGED], (unsigned long long)n_flag);
        }
    }
}

int seq_wide_threads() { return WIDE_THREADS; }

static const void *wide_kernel_ptr(bool disc, bool pack)
{
    if (pack) return disc ? (const void *)seq_wide_kernel<true, true> : (const void *)seq_wide_kernel<false, true>;
    return disc ? (const void *)seq_wide_kernel<true, false> : (const void *)seq_wide_kernel<false, false>;
}

cudaError_t seq_wide_prepare(bool disc, bool pack, size_t smem, int threads, int *blocks_per_sm)
{
    const void *k = wide_kernel_ptr(disc, pack);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem);
}

void seq_wide_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream)
{
    void *args[] = {(void *)&a};
    cudaLaunchKernel(wide_kernel_ptr(a.disc != 0, a.wide_pack != 0), dim3(grid), dim3(threads), args, smem, stream);
}

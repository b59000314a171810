// seq_wide.cu -- sampler-driven sequential chronological MC for systems of more than 32 units when
// the whole year fits one shared-memory timeline (BASELINE config 5: 1024 units, 8736 hours).
// Same model, sampler and per-year integers as seq_fast.cu / seq_team.cu / seq_mc.cu
// (run_sequential_mc, GeneratingAdequacy/PowerSystemAdequacy.jl:214-269; indices per
// Montecarlo_seq/seqMain.m:160-176, Montecarlo_seq/calnlc.m:22-34).
//
// One thread block owns one simulated year at a time; its hour timeline (one int32 of net capacity
// change per hour) lives in shared memory.  A lane owns one unit and walks its stream block by block:
// Philox4x32-10, four tick durations, a running 64-bit event time in registers (T - 1 of the last
// event, so the hour of an event is one funnel shift and "the unit has left the year" is the in-year
// test of its fourth event), four integer-MW deltas added to the timeline with shared-memory atomics
// (out-of-year events are not branched around: they add into a per-lane dummy slot behind the year).
//
// Generation runs in two phases per warp (round 2, third version):
//  1. STATIC.  The units are sorted by transition rate (psra_set_system), so the 32 units of a group
//     of consecutive queue positions need about the same number of Philox blocks per year.  The host
//     gives every group a block count B_g that each of its units needs with probability >= ~0.7
//     (SeqArgs::wide_sblk).  The warp runs its groups (g = warp, warp + nwarps, ...) lane = unit through
//     block 0 (initial-state draw + three durations) and blocks 1 .. B_g - 1 in a loop with no per-lane
//     control flow at all: no work queue, no votes, no special case for the first block.  Units that
//     are still inside the year afterwards are appended (ballot / popc compaction) to the warp's
//     to-do list in shared memory: one 64-bit entry {T - 1 (38 bits), initial state, queue position,
//     next block}.
//  2. QUEUE.  The lanes take to-do entries from a per-warp queue (shared-memory counter) and walk them
//     to the end of the year; a finished lane takes the next entry.  This is the only place with
//     per-lane control flow, and it covers about a third of the blocks.
// Both phases are warp-private (the timeline is the only shared state), so the generation phase needs
// no block barrier.
//
// Evaluation: the warps reduce the hour deltas to per-32-hour-word sums (lane = word, 128-bit loads,
// skewed so that a quarter-warp covers all 32 banks) and scan them inside their block of 32 words;
// after a barrier lane = word again: the capacity entering the word is the capacity at hour 0 plus the
// sums of the blocks in front plus the prefix inside the block, and a word that can contain loss of load (capacity + negative hour deltas < maximum load of the
// word) is walked hour by hour by its lane (LOL hours, deficit entries, int64 ENS), and every lane clears
// its word -- so the rare resolution work is spread over all warps instead of serialising on one of them
// while the others wait at the barrier (measured: 15 % of the warp time).  The per-year sums meet in
// shared-memory scalars that are double buffered by year parity, so thread 0 writes year y out while the
// other warps already generate y + 1 (3 barriers per year, no serial phase).
#include <limits.h>

#include <algorithm>

#include "psra_internal.cuh"
#include "seq_args.cuh"

#define WIDE_MAX_WARPS 8
#ifndef WIDE_BPS4
#define WIDE_BPS4 5      // resident blocks per SM the 4-warp instantiation is compiled for (register budget)
#endif

// -DWIDE_PROFILE: per-phase clock64() sums over the warps into acc[16 ..] (scripts/wide_phases.py); off in the product build
#ifdef WIDE_PROFILE
#define WIDE_T(slot) do { const long long t_now = clock64(); prof[slot] += (unsigned long long)(t_now - t_prev); t_prev = t_now; } while (0)
#else
#define WIDE_T(slot) do { } while (0)
#endif

struct WideShared {     // one per year parity
    int capacity;       // sum of the capacities of the units that start the year UP
    unsigned int lolh, entries;
    int pad;
    unsigned long long ens;
};

static __host__ __device__ inline int wide_todo_cap(int U, int nwarps)   // to-do entries per warp
{
    const int ngroups = (U + 31) >> 5;
    return ((ngroups + nwarps - 1) / nwarps) * 32;
}

size_t seq_wide_smem_bytes(int Wd, int U, int nwarps)
{
    const size_t Wd4 = (size_t)((Wd + 3) & ~3);
    size_t b = sizeof(int32_t) * ((size_t)Wd * 32 + 32);                  // hour timeline + one dummy slot per lane
    // the to-do lists of the generation phase and the word sums / negative sums of the evaluation share one region
    b += std::max(2 * sizeof(int32_t) * Wd4, sizeof(unsigned long long) * (size_t)nwarps * wide_todo_cap(U, nwarps));
    b += 2 * sizeof(WideShared) + (32 + WIDE_MAX_WARPS + 16) * sizeof(int32_t) + 8 * sizeof(unsigned long long);   // + one always-zero word per lane, queue heads, block totals, word-block sums
    return (b + 15) & ~(size_t)15;
}

// add `delta` to the hour slot of an event, or to the lane's dummy slot when the event lies beyond the year
__device__ __forceinline__ void wide_scatter(uint32_t tl_s, uint32_t dummy_s, uint32_t hs, uint32_t H, int delta, unsigned int &n_events)
{
    asm volatile("{\n .reg .pred p;\n .reg .b32 ad;\n"
                 " setp.lt.u32 p, %1, %2;\n"
                 " mad.lo.u32 ad, %1, 4, %3;\n"
                 " selp.b32 ad, ad, %4, p;\n"
                 " red.shared.add.s32 [ad], %5;\n"
                 " @p add.u32 %0, %0, 1;\n}\n"
                 : "+r"(n_events)
                 : "r"(hs), "r"(H), "r"(tl_s), "r"(dummy_s), "r"(delta)
                 : "memory");
}

// one sampler duration in ticks: RN_int64(max(mean_ticks * E(x), 1 tick))
template <bool kDisc>
__device__ __forceinline__ unsigned long long wide_dur(float mean_ticks, uint32_t x, bool up_state, uint32_t one_bits)
{
    unsigned long long t = ticks_rn(fmaxf(__fmul_rn(mean_ticks, neglog_u32(x, one_bits)), 1.0f));
    if constexpr (kDisc) t = ((t + (up_state ? (1ull << 23) : ((1ull << 24) - 1ull))) >> 24) << 24;
    return t;
}

// hour of an event at tick T, from T - 1: ceil(T / 2^24) - 1 = (T - 1) >> 24
__device__ __forceinline__ uint32_t wide_hour(unsigned long long tm1)
{
    return __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
}

// one full Philox block of a unit's stream (draws 4 nb .. 4 nb + 3, nb >= 1): four durations, four events.
// m_a / d_a: mean of the state opposite to the initial one / delta of the event that ends it (draws 0, 2 of a block);
// m_b / -d_a: the initial state (draws 1, 3).  Returns the hour of the fourth event.
template <bool kDisc>
__device__ __forceinline__ uint32_t wide_block(const SeqArgs &a, uint32_t c_lo, uint32_t c_hi, uint32_t u, uint32_t nb,
                                               float m_a, float m_b, int d_a, bool sdn, unsigned long long &tm1,
                                               uint32_t tl_s, uint32_t dummy_s, uint32_t Hl, unsigned int &ne, uint32_t one_bits)
{
    uint32_t x[4];
    philox4x32_10_rk(c_lo, c_hi, u, nb, a.rk, x);
    const unsigned long long t1 = tm1 + wide_dur<kDisc>(m_a, x[0], sdn, one_bits);
    const unsigned long long t2 = t1 + wide_dur<kDisc>(m_b, x[1], !sdn, one_bits);
    const unsigned long long t3 = t2 + wide_dur<kDisc>(m_a, x[2], sdn, one_bits);
    const unsigned long long t4 = t3 + wide_dur<kDisc>(m_b, x[3], !sdn, one_bits);
    const uint32_t h4 = wide_hour(t4);
    wide_scatter(tl_s, dummy_s, wide_hour(t1), Hl, d_a, ne);
    wide_scatter(tl_s, dummy_s, wide_hour(t2), Hl, -d_a, ne);
    wide_scatter(tl_s, dummy_s, wide_hour(t3), Hl, d_a, ne);
    wide_scatter(tl_s, dummy_s, h4, Hl, -d_a, ne);
    tm1 = t4;
    return h4;
}

// kWarps = warps per block the instantiation is compiled for (register budget): 4 -> 5 blocks per SM (96 registers),
// 6 -> 4 blocks (80), 8 -> 3 blocks (80)
template <bool kDisc, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, kWarps <= 4 ? WIDE_BPS4 : kWarps <= 6 ? 4 : 3) seq_wide_kernel(const SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int tl_len = a.Wd * 32, Wd4 = (a.Wd + 3) & ~3;
    const int ngroups = (a.U + 31) >> 5, todo_cap = wide_todo_cap(a.U, nwarps);
    int32_t *tl = reinterpret_cast<int32_t *>(smem_raw);                 // [tl_len + 32]
    // one region, two lives: the warps' to-do lists while the year is generated, the word sums while it is evaluated
    // (block barriers separate the two)
    int32_t *wsum = tl + tl_len + 32;                                    // [Wd4]
    int32_t *wneg = wsum + Wd4;
    unsigned long long *todo = reinterpret_cast<unsigned long long *>(wsum) + (size_t)warp * todo_cap;   // 16-byte aligned
    const int region = max(2 * Wd4, 2 * nwarps * todo_cap);              // in int32 (both terms are multiples of 4)
    const int32_t *__restrict__ s_lmax = a.lmax;                         // word maxima of the load: read-only, L1-resident (1 KB)
    WideShared *sh_all = reinterpret_cast<WideShared *>(wsum + region);
    int32_t *zero32 = reinterpret_cast<int32_t *>(sh_all + 2);           // [32], stays 0
    int32_t *qheads = zero32 + 32;                                       // [WIDE_MAX_WARPS]
    const uint32_t tl_s = (uint32_t)__cvta_generic_to_shared(tl);
    uint32_t dummy_s = tl_s + 4u * (uint32_t)(tl_len + lane);
    // keep the shared addresses in registers: left alone, ptxas rebuilds them in every iteration of the generation
    // loops; one_bits is the exponent pattern of 1.0f as an opaque register (see neglog_u32)
    uint32_t tl_o = tl_s, one_bits = 0x3F800000u;
    const uint32_t qh_s = (uint32_t)__cvta_generic_to_shared(qheads + warp);
    const uint32_t zero_s = (uint32_t)__cvta_generic_to_shared(zero32 + lane);
    asm volatile("" : "+r"(tl_o), "+r"(dummy_s), "+r"(one_bits));

    for (int i = threadIdx.x; i < tl_len + 32; i += blockDim.x) tl[i] = 0;
    if (threadIdx.x < 32) zero32[threadIdx.x] = 0;
    if (threadIdx.x < 2) {
        WideShared *z = sh_all + threadIdx.x;
        z->capacity = 0; z->lolh = 0u; z->entries = 0u; z->ens = 0ull;
    }
    __syncthreads();

    // block totals (thread 0 only; in shared memory: they would cost 14 registers for one update per year)
    unsigned long long *bacc = reinterpret_cast<unsigned long long *>(qheads + WIDE_MAX_WARPS);   // [7]: LOL, ENS, ENT, YWL, LOL2, ENS2 lo / hi
    int32_t *bsum = reinterpret_cast<int32_t *>(bacc + 8);               // [16]: sums of the 32-word blocks of the year
    if (threadIdx.x < 7) bacc[threadIdx.x] = 0ull;
    unsigned long long ev64 = 0ull, n_jobs = 0ull;
    unsigned int n_flag = 0;
    const uint32_t Hu = (uint32_t)a.H;
    const unsigned long long parked = 1ull << 55;                      // T - 1 of a lane without a unit: far beyond any year
    // T - 1 before the first duration.  MATLAB discretisation: a unit that fails after d whole hours is DOWN from
    // hour d + 1 (seq_mcsampling.m:63)
    const unsigned long long start_m1 = (kDisc ? (1ull << PSRA_TICK_SHIFT) : 0ull) - 1ull;
    const uint32_t thr_mask = a.init_mode == PSRA_INIT_STATIONARY ? 0xffffffffu : 0u;
    const uint32_t lt_mask = (1u << lane) - 1u;

#ifdef WIDE_PROFILE
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#endif
    int par = 0;
    for (long long cl = blockIdx.x; cl < a.nchains; cl += gridDim.x, par ^= 1) {
        const unsigned long long chain = (unsigned long long)(a.chain_base + cl);
        const uint32_t c_lo = (uint32_t)chain, c_hi = (uint32_t)(chain >> 32);
        WideShared *sh = sh_all + par;
        unsigned int ne = 0u;               // in-year events of this lane's units
        int cap_up = 0, list_n = 0;

        // ---- generation, static phase: lane = unit of group g, blocks 0 .. B_g - 1
        // (groups are sorted by demand; rounds alternate the direction so that every warp gets the same mix)
        for (int g0 = 0, rnd = 0; g0 < ngroups; g0 += nwarps, rnd ^= 1) {
            const int g = g0 + (rnd ? nwarps - 1 - warp : warp);
            if (g >= ngroups) continue;
            const int pos = g * 32 + lane;
            const uint4 rec = __ldg(&a.wide_tab[pos]);      // one 16-byte record per queue position (padded to whole groups)
            const uint32_t u = (uint32_t)__ldg(&a.order[pos]);
            const int B = a.wide_sblk[g];
            const bool valid = pos < a.U;
            const uint32_t Hl = valid ? Hu : 0u;            // a padding lane keeps all its events out of the year
            const int cu = (int)rec.x;
            uint32_t x[4];
            philox4x32_10_rk(c_lo, c_hi, u, 0u, a.rk, x);
            const bool sdn = x[0] < (rec.w & thr_mask);     // draw 0 of a stream is the initial state
            cap_up += sdn ? 0 : cu;
            const float m_a = __uint_as_float(sdn ? rec.y : rec.z);   // draws 0, 2 of a block: the state opposite to the initial one
            const float m_b = __uint_as_float(sdn ? rec.z : rec.y);   // draws 1, 3: the initial state
            const int d_a = sdn ? -cu : cu;                 // draws 0, 2 end with the unit back in its initial state
            const unsigned long long t2 = start_m1 + wide_dur<kDisc>(m_b, x[1], !sdn, one_bits);
            const unsigned long long t3 = t2 + wide_dur<kDisc>(m_a, x[2], sdn, one_bits);
            unsigned long long tm1 = t3 + wide_dur<kDisc>(m_b, x[3], !sdn, one_bits);
            uint32_t h4 = wide_hour(tm1);
            wide_scatter(tl_o, dummy_s, wide_hour(t2), Hl, -d_a, ne);
            wide_scatter(tl_o, dummy_s, wide_hour(t3), Hl, d_a, ne);
            wide_scatter(tl_o, dummy_s, h4, Hl, -d_a, ne);
            for (int b = 1; b < B; b++)
                h4 = wide_block<kDisc>(a, c_lo, c_hi, u, (uint32_t)b, m_a, m_b, d_a, sdn, tm1, tl_o, dummy_s, Hl, ne, one_bits);
            // still inside the year: to the warp's to-do list
            const bool more = h4 < Hl;
            const uint32_t mm = __ballot_sync(0xffffffffu, more);
            if (more) {
                const uint32_t hi = (uint32_t)(tm1 >> 32) | (sdn ? 0x40u : 0u) | ((uint32_t)pos << 8) | ((uint32_t)B << 20);
                todo[list_n + __popc(mm & lt_mask)] = ((unsigned long long)hi << 32) | (uint32_t)tm1;
            }
            list_n += __popc(mm);
            n_jobs += valid ? (unsigned)B : 0u;
        }
        if (lane == 0) qheads[warp] = 32;
        __syncwarp();
        WIDE_T(0);

        // ---- generation, queue phase: the lanes walk the to-do entries to the end of the year
        {
            int k = lane;
            bool busy = k < list_n;
            uint32_t u = 0u, nb = 1u;
            float m_a = 1.f, m_b = 1.f;
            int d_a = 0;
            bool sdn = false;
            unsigned long long tm1 = parked;
            auto take = [&]() {
                const unsigned long long e = todo[k];
                const uint32_t hi = (uint32_t)(e >> 32);
                const int pos = (int)((hi >> 8) & 0xfffu);
                const uint4 rec = __ldg(&a.wide_tab[pos]);
                u = (uint32_t)__ldg(&a.order[pos]);
                sdn = (hi & 0x40u) != 0u;
                nb = hi >> 20;
                tm1 = ((unsigned long long)(hi & 0x3fu) << 32) | (uint32_t)e;
                m_a = __uint_as_float(sdn ? rec.y : rec.z);
                m_b = __uint_as_float(sdn ? rec.z : rec.y);
                d_a = sdn ? -(int)rec.x : (int)rec.x;
                n_jobs -= nb;
            };
            if (busy) take();
            while (__any_sync(0xffffffffu, busy)) {
                const uint32_t h4 = wide_block<kDisc>(a, c_lo, c_hi, u, nb, m_a, m_b, d_a, sdn, tm1, tl_o, dummy_s, Hu, ne, one_bits);
                nb++;
                if (busy && h4 >= Hu) {         // unit done: take the next entry of the warp's queue
                    n_jobs += nb;
                    // One plain atomic per finishing lane.  Left alone, ptxas turns an atomic add of a constant to a
                    // warp-uniform address into a warp-aggregated sequence (vote, leader election, popc, shuffle: ~25
                    // instructions) that every iteration of the loop would pay for; the address therefore gets a per-lane
                    // offset the compiler cannot see through (a shared-memory word per lane that always holds 0).
                    asm volatile("{\n .reg .b32 z;\n ld.volatile.shared.u32 z, [%2];\n add.u32 z, z, %1;\n"
                                 " atom.shared.add.u32 %0, [z], 1;\n}\n"
                                 : "=r"(k) : "r"(qh_s), "r"(zero_s) : "memory");
                    busy = k < list_n;
                    if (busy) take();
                    else { tm1 = parked; m_a = 1.f; m_b = 1.f; d_a = 0; nb = 1u; }
                }
            }
        }
        ev64 += ne;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cap_up += __shfl_xor_sync(0xffffffffu, cap_up, d);
        if (lane == 0 && cap_up) atomicAdd(&sh->capacity, cap_up);
        WIDE_T(1);
        __syncthreads();
        WIDE_T(2);

        // ---- word sums of the hour deltas: lane = word, 128-bit loads skewed by the lane (conflict-free); the warp scans the
        //      sums of its 32 words right away: wsum[w] = sum of the words in front of w inside its block of 32 words,
        //      bsum[block] = sum of the whole block
        for (int w0 = warp * 32; w0 < a.Wd; w0 += nwarps * 32) {
            const int w = w0 + lane;
            // the word's maximum load goes into the stored negative sum here: the global load (what is left of the L1 beside five
            // timelines does not keep the table) completes under the shared-memory work of this phase instead of sitting
            // on the critical path of the evaluation behind the barrier
            const int lmx = w < a.Wd ? __ldg(&s_lmax[w]) : 0;
            int s = 0, n = 0;
            if (w < a.Wd) {
                const int4 *row = reinterpret_cast<const int4 *>(tl + w * 32);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int4 v = row[(j + lane) & 7];
                    s += (v.x + v.y) + (v.z + v.w);
                    n += (min(v.x, 0) + min(v.y, 0)) + (min(v.z, 0) + min(v.w, 0));
                }
            }
            const int incl = warp_incl_scan(s, lane);
            if (w < a.Wd) { wsum[w] = incl - s; wneg[w] = n - lmx; }
            if (lane == 31) bsum[w0 >> 5] = incl;
        }
        WIDE_T(3);
        __syncthreads();
        WIDE_T(4);

        // ---- evaluation: lane = word again (the mapping of the word sums): the capacity entering the word (capacity at
        //      hour 0 + the sums of the blocks in front + the prefix inside the block), the conservative test "capacity + negative hour deltas < maximum load of the word", and the lane walks the
        //      32 hours of such a word itself (rare: ~3 words per year, neighbours share one divergent pass; PSA.jl:253
        //      strict compare, deficit entries per calnlc.m:22-34).  Every lane clears its word afterwards.
        {
            unsigned int lolh = 0, entries = 0;
            long long ens_lane = 0;
            const int nwords = a.Wd;
            const int cap0 = sh->capacity;
            for (int w0 = warp * 32; w0 < nwords; w0 += nwarps * 32) {
                const bool valid = w0 + lane < nwords;
                const int w = valid ? w0 + lane : nwords - 1;
                int c_in = cap0 + wsum[w];                          // capacity entering the word
                for (int j = 0; j < (w0 >> 5); j++) c_in += bsum[j];
                const bool need = valid && (c_in + wneg[w] < 0);        // capacity + negative hour deltas < maximum load of the word
                uint32_t nm = __ballot_sync(0xffffffffu, need);
                n_flag += __popc(nm);
                int4 *row = reinterpret_cast<int4 *>(tl + w * 32);
                // a flagged word is resolved by the whole warp, lane = hour: one load, one shuffle scan, two ballots
                // (a single lane walking its 32 hours costs ~8 times the warp-instructions and leaves the other warps
                // of the block waiting at the barrier behind it)
                while (nm) {
                    const int src = __ffs(nm) - 1;
                    nm &= nm - 1u;
                    const int wq = w0 + src;
                    const int cq = __shfl_sync(0xffffffffu, c_in, src);      // capacity entering the word
                    const int hy = wq * 32 + lane;
                    const int L = __ldg(&a.load[hy]);                        // zero beyond the year: never a loss there
                    const int Lprev = __ldg(&a.load[max(wq * 32 - 1, 0)]);   // issued with it: one L2 round trip, not two
                    const int c = cq + warp_incl_scan(tl[hy], lane);
                    const bool lol = c < L;
                    const uint32_t lm = __ballot_sync(0xffffffffu, lol);
                    if (lm) {
                        const uint32_t prev0 = (wq > 0 && cq < Lprev) ? 1u : 0u;                      // the hour before the word
                        if (lane == 0) {
                            lolh += (unsigned int)__popc(lm);
                            entries += (unsigned int)__popc(lm & ~((lm << 1) | prev0));                 // calnlc.m:22-34
                        }
                        if (lol) {
                            ens_lane += (long long)(L - c);
                            if (a.fail) atomicAdd(&a.fail[hy], 1u);
                        }
                    }
                }
                __syncwarp();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; j++) row[(j + lane) & 7] = make_int4(0, 0, 0, 0);
                }
            }
            if (__any_sync(0xffffffffu, lolh != 0u)) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    lolh += __shfl_xor_sync(0xffffffffu, lolh, d);
                    entries += __shfl_xor_sync(0xffffffffu, entries, d);
                }
                const long long ens = warp_sum_ll(ens_lane);
                if (lane == 0) {
                    atomicAdd(&sh->lolh, lolh);
                    atomicAdd(&sh->entries, entries);
                    atomicAdd(&sh->ens, (unsigned long long)ens);
                }
            }
        }
        WIDE_T(5);
        __syncthreads();
        WIDE_T(6);

        // ---- per-year indices: thread 0 writes year `cl` out and re-arms its scalars for the year after next while the
        //      other warps already generate the next year with the other set
        if (threadIdx.x == 0) {
            const unsigned int lolh = sh->lolh, entries = sh->entries;
            const long long ens = (long long)sh->ens;
            sh->capacity = 0; sh->lolh = 0u; sh->entries = 0u; sh->ens = 0ull;
            if (a.lol) a.lol[cl] = lolh;
            if (a.ens) a.ens[cl] = ens;
            if (a.ent) a.ent[cl] = entries;
            if (a.group_lol && lolh) atomicAdd(&a.group_lol[(cl + a.group_phase) / a.group], (unsigned long long)lolh);
            if (lolh) seq_hist_add(a, ens);
            if (lolh) {
                bacc[0] += lolh; bacc[1] += (unsigned long long)ens; bacc[2] += entries;
                bacc[3] += 1ull;
                bacc[4] += (unsigned long long)lolh * lolh;
                const unsigned long long e = (unsigned long long)ens;
                const unsigned long long plo = e * e, phi = __umul64hi(e, e);
                const unsigned long long olo = bacc[5], nlo = olo + plo;
                bacc[6] += phi + (nlo < olo ? 1ull : 0ull);
                bacc[5] = nlo;
            }
        }
        WIDE_T(7);
    }
#ifdef WIDE_PROFILE
    if (lane == 0)
        for (int i = 0; i < 8; i++) atomicAdd(&a.acc[16 + i], prof[i]);
#endif

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
            if (bacc[0]) atomicAdd(&a.acc[ACC_LOL], bacc[0]);
            if (bacc[1]) atomicAdd(&a.acc[ACC_ENS], bacc[1]);
            if (bacc[2]) atomicAdd(&a.acc[ACC_ENT], bacc[2]);
            if (bacc[3]) atomicAdd(&a.acc[ACC_YWL], bacc[3]);
            if (bacc[4]) atomicAdd(&a.acc[ACC_LOL2], bacc[4]);
            if (bacc[5] | bacc[6]) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], bacc[5], bacc[6]);
        }
        atomicAdd(&a.acc[ACC_FLAGGED], (unsigned long long)n_flag);      // words resolved hour by hour
    }
}

int seq_wide_max_warps() { return WIDE_MAX_WARPS; }

static const void *wide_kernel_ptr(bool disc, int threads)
{
    if (threads <= 128) return disc ? (const void *)seq_wide_kernel<true, 4> : (const void *)seq_wide_kernel<false, 4>;
    if (threads <= 192) return disc ? (const void *)seq_wide_kernel<true, 6> : (const void *)seq_wide_kernel<false, 6>;
    return disc ? (const void *)seq_wide_kernel<true, 8> : (const void *)seq_wide_kernel<false, 8>;
}

cudaError_t seq_wide_prepare(bool disc, size_t smem, int threads, int *blocks_per_sm)
{
    const void *k = wide_kernel_ptr(disc, threads);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem);
}

void seq_wide_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream)
{
    void *args[] = {(void *)&a};
    cudaLaunchKernel(wide_kernel_ptr(a.disc != 0, threads), dim3(grid), dim3(threads), args, smem, stream);
}

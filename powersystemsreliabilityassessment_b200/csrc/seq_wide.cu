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

#define WIDE_THREADS 192
#define WIDE_BLOCKS_PER_SM 5

struct WideShared {     // one per year parity
    int capacity;       // sum of the capacities of the units that start the year UP
    int queue_head;     // next position of the unit order that has not been handed out
    unsigned int lolh, entries;
    unsigned long long ens;
};

size_t seq_wide_smem_bytes(int Wd)
{
    size_t b = sizeof(int32_t) * ((size_t)Wd * 32 + 32);            // hour timeline + one dummy slot per lane
    b += 3 * sizeof(int32_t) * (size_t)((Wd + 3) & ~3);              // word sums, negative sums, word maxima of the load
    b += 2 * sizeof(WideShared) + 16;
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

template <bool kDisc>
__global__ void __launch_bounds__(WIDE_THREADS, WIDE_BLOCKS_PER_SM) seq_wide_kernel(const SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int Hpad = a.Wd * 32, Wd4 = (a.Wd + 3) & ~3;
    int32_t *tl = reinterpret_cast<int32_t *>(smem_raw);                 // [Hpad + 32]
    int32_t *wsum = tl + Hpad + 32;                                      // [Wd4]
    int32_t *wneg = wsum + Wd4;
    int32_t *s_lmax = wneg + Wd4;
    WideShared *sh_all = reinterpret_cast<WideShared *>(s_lmax + Wd4);   // [2], 8-byte aligned (Wd4 is a multiple of 4)
    const uint32_t tl_s = (uint32_t)__cvta_generic_to_shared(tl);
    uint32_t dummy_s = tl_s + 4u * (uint32_t)(Hpad + lane);
    // keep the two shared addresses in registers: left alone, ptxas rebuilds them (S2R SR_CgaCtaId, LEA, ...) in every
    // iteration of the generation loop
    uint32_t tl_o = tl_s;
    asm volatile("" : "+r"(tl_o), "+r"(dummy_s));

    for (int i = threadIdx.x; i < a.Wd; i += blockDim.x) s_lmax[i] = a.lmax[i];
    for (int i = threadIdx.x; i < Hpad + 32; i += blockDim.x) tl[i] = 0;
    __syncthreads();

    unsigned long long acc_lol = 0, acc_ent = 0, acc_ywl = 0, acc_lol2 = 0, acc_e2lo = 0, acc_e2hi = 0;
    long long acc_ens = 0;
    unsigned int n_events = 0, n_jobs = 0, n_flag = 0;
    const unsigned long long end_t = (unsigned long long)a.H << PSRA_TICK_SHIFT;
    const unsigned long long parked = 0x00800000ull << 32;             // event time of a lane without a unit: far beyond any year
    const bool stationary = a.init_mode == PSRA_INIT_STATIONARY;

    if (threadIdx.x < 2) {
        WideShared *z = sh_all + threadIdx.x;
        z->capacity = 0; z->queue_head = (int)blockDim.x; z->lolh = 0u; z->entries = 0u; z->ens = 0ull;
    }
    __syncthreads();

    int par = 0;
    for (long long cl = blockIdx.x; cl < a.nchains; cl += gridDim.x, par ^= 1) {
        const unsigned long long chain = (unsigned long long)(a.chain_base + cl);
        WideShared *sh = sh_all + par;

        // ---- generation: lane = unit, block after block; finished lanes pull the next unit from the queue
        int pos = threadIdx.x;              // position in the unit order
        bool busy = pos < a.U;
        int u = 0, cu = 0, cap_up = 0;
        float mup = 1.f, mdn = 1.f;
        uint32_t thr = 0u, nb = 0u;
        bool s0u = true;
        unsigned long long t = 0ull;
        auto take_unit = [&]() {
            const uint4 rec = __ldg(&a.wide_tab[pos]);      // one 16-byte record per queue position
            u = __ldg(&a.order[pos]);
            cu = (int)rec.x;
            mup = __uint_as_float(rec.y);
            mdn = __uint_as_float(rec.z);
            thr = rec.w;
            nb = 0u;
            // MATLAB discretisation: a unit that fails after d whole hours is DOWN from hour d+1 (seq_mcsampling.m:63)
            t = kDisc ? (1ull << PSRA_TICK_SHIFT) : 0ull;
        };
        if (busy) take_unit();
        while (__any_sync(0xffffffffu, busy)) {
            uint32_t x[4];
            philox4x32_10_rk((uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)u, nb, a.rk, x);
            const bool first = nb == 0u;
            if (first) {                    // draw 0 of a stream is the initial state
                s0u = !(stationary && x[0] < thr);
                if (busy && s0u) cap_up += cu;
            }
            const float m_a = s0u ? mdn : mup;      // draws 0, 2 of a block: state s0^1
            const float m_b = s0u ? mup : mdn;      // draws 1, 3: state s0
            const unsigned long long p1 = first ? 0ull : dur_ticks_disc(m_a, x[0], !s0u, kDisc);
            const unsigned long long p2 = p1 + dur_ticks_disc(m_b, x[1], s0u, kDisc);
            const unsigned long long p3 = p2 + dur_ticks_disc(m_a, x[2], !s0u, kDisc);
            const unsigned long long p4 = p3 + dur_ticks_disc(m_b, x[3], s0u, kDisc);
            // hour of an event at tick T: ceil(T / 2^24) - 1 = (T - 1) >> 24
            const unsigned long long bm1 = busy ? t - 1ull : parked;
            const int delta_a = s0u ? cu : -cu;     // draws 0, 2 toggle the unit back to s0
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned long long tm1 = bm1 + (q == 0 ? p1 : q == 1 ? p2 : q == 2 ? p3 : p4);
                if (q == 0 && first) tm1 = parked;
                const uint32_t hs = __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
                wide_scatter(tl_o, dummy_s, hs, (uint32_t)a.H, (q & 1) ? -delta_a : delta_a, n_events);
            }
            t += p4;
            nb++;
            n_jobs += busy ? 1u : 0u;
            if (busy && t > end_t) {        // unit done: take the next one of the block's queue
                pos = atomicAdd(&sh->queue_head, 1);
                busy = pos < a.U;
                if (busy) take_unit();
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cap_up += __shfl_xor_sync(0xffffffffu, cap_up, d);
        if (lane == 0 && cap_up) atomicAdd(&sh->capacity, cap_up);
        __syncthreads();

        // ---- word sums of the hour deltas: lane = word, hour index skewed by the lane (conflict-free)
        for (int w0 = warp * 32; w0 < a.Wd; w0 += nwarps * 32) {
            const int w = w0 + lane;
            if (w < a.Wd) {
                const int32_t *row = tl + w * 32;
                int s = 0, n = 0;
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const int d = row[(j + lane) & 31];
                    s += d;
                    n += min(d, 0);
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
            const int cs_lane = sh->capacity + incl - loc;          // capacity entering the lane's run
            const bool flagged = (lmin != INT_MAX) && (cs_lane + lmin < 0);
            const uint32_t fm = __ballot_sync(0xffffffffu, flagged);
            if (warp == 0) n_flag += __popc(fm);
            for (int src = warp; src < 32; src += nwarps) {         // runs owned by this warp
                if ((fm >> src) & 1u) {                             // rare: the run may contain loss of load
                    int c_in = __shfl_sync(0xffffffffu, cs_lane, src);
                    for (int k = 0; k < wpl; k++) {
                        const int wq = src * wpl + k;
                        if (wq >= nwords) break;
                        if (c_in + wneg[wq] < s_lmax[wq]) {         // resolve the word hour by hour, lane = hour
                            const int c = c_in + warp_incl_scan(tl[wq * 32 + lane], lane);
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
                // clear the run (wpl * 32 hours; the dummy slots behind the year may keep their garbage)
                int4 *t4 = reinterpret_cast<int4 *>(tl + src * wpl * 32);
                const int n4 = min(wpl, max(0, nwords - src * wpl)) * 8;
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
            sh->capacity = 0; sh->queue_head = (int)blockDim.x; sh->lolh = 0u; sh->entries = 0u; sh->ens = 0ull;
            if (a.lol) a.lol[cl] = lolh;
            if (a.ens) a.ens[cl] = ens;
            if (a.ent) a.ent[cl] = entries;
            if (a.group_lol && lolh) atomicAdd(&a.group_lol[cl / a.group], (unsigned long long)lolh);
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

    unsigned long long ev = n_events, jb = n_jobs;
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

cudaError_t seq_wide_prepare(size_t smem, int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(seq_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(seq_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, seq_wide_kernel<false>, WIDE_THREADS, smem);
}

void seq_wide_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream)
{
    if (a.disc) seq_wide_kernel<true><<<grid, threads, smem, stream>>>(a);
    else seq_wide_kernel<false><<<grid, threads, smem, stream>>>(a);
}

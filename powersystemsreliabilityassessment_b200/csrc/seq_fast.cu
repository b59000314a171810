// seq_fast.cu -- sampler-driven sequential chronological MC for systems of <= 32 units (RTS-79).
//
// Same model and same per-year integers as seq_mc.cu (run_sequential_mc,
// GeneratingAdequacy/PowerSystemAdequacy.jl:214-269; indices per Montecarlo_seq/seqMain.m:160-176,
// Montecarlo_seq/calnlc.m:22-34), reorganised so that lanes are busy:
//
//  * Time is kept in integer ticks of 2^-24 h.  A sampler duration is D = RN(max(mean*2^24*E, 1))
//    ticks, so every residual of the reference's `ttf -= 1.0` / `ttf += D` recurrence
//    (PSA.jl:239-246) is an exact FP64 number and the event times are plain prefix sums
//    T_k = D_0 + ... + D_k; event k toggles the unit in hour ceil(T_k / 2^24).
//  * Prefix sums parallelise: in a "wave" every unit that does not yet cover the current timeline
//    segment asks for 1..4 Philox blocks (4 draws each); the requested (unit, block) jobs are
//    spread over the 32 lanes (lane = job, not lane = unit), each lane turns its block into four
//    durations and a local prefix, and a segmented warp-shuffle scan chains the blocks of a unit.
//    Lane utilisation no longer depends on the 6x spread of the units' event rates.
//  * The warp's shared-memory timeline is a ring of two segments kept at two resolutions: per
//    32-hour word the sum of the integer MW deltas and the sum of the negative deltas (atomicAdd at
//    scatter time), and a compact event list (hour, unit, sign) per ring half in which the events
//    of one word are linked (atomicExch on the word's list head).  Hour resolution is rebuilt from
//    that list only for the few words that can contain loss of load.  The rare
//    events beyond the ring wait in a small pending list.  A wave is one round of <= 32 jobs:
//    blocks for the units that are short of the current segment first, the spare lanes
//    pre-generate blocks towards the end of the next segment.  For RTS-79 the whole year is one
//    segment (no ring switch, no pending traffic); that variant (template kTwo = false) first
//    generates the first blocks of every unit lane = unit without any scheduling (static phase),
//    keeps the two word sums packed in one int32 (kPack; values and list heads are two arrays, so
//    the atomics of a scatter step spread over all 32 banks) and pads the word table to 32 equal
//    runs of neutral records, so the evaluation scan is branch-free.
//  * Evaluation: lane = run of consecutive words.  One shuffle scan per segment over the word
//    sums gives the capacity entering each run; conservative flag
//    capacity + (negative deltas of the word) < max load of the word (table staged in shared
//    memory).  Flagged words (about 3 per RTS-79 year) are resolved hour by hour, lane = hour:
//    the word's deltas are gathered by walking its linked event list, a shuffle scan turns them
//    into the 32 capacities, which are compared with the load curve staged once per block in shared memory;
//    __ballot_sync/__popc give LOL hours and deficit entries, per-lane int64 accumulators the ENS.
#include <limits.h>

#include "psra_internal.cuh"
#include "seq_args.cuh"

#define FAST_NB_MAX 4                  // Philox blocks a unit may get per wave
#define FAST_PEND_CAP 256              // far-future events (beyond the two-segment ring)
// single-segment variants run 32 warps per SM (<= 64 registers); the ring variants need a few more registers
#define FAST_MAX_THREADS(two) ((two) ? 896 : 1024)

struct FastWarpSmem {                  // per-warp scratch that precedes the event lists
    unsigned long long t_run[32];      // time (ticks) of the last generated event of each unit
    unsigned long long acc[8];         // per-warp accumulators (lane 0): lol, ens, entries, years with loss, lol^2, ens^2 lo/hi, state transitions
    unsigned int diag[8];              // waves, jobs, ahead jobs, flagged runs, max list depth
    unsigned char jobmap[32];
};

// word records a warp keeps: both ring halves, or -- single segment -- the words of the segment padded to 32 equal
// runs (one run per lane in the evaluation scan; the padding records stay neutral, so the scan needs no bounds checks)
__host__ __device__ inline int fast_words_alloc(int seg_words, bool two_halves)
{
    return two_halves ? ((2 * seg_words + 3) & ~3) : 32 * ((seg_words + 31) / 32);
}

// two_halves: the ring needs its second half (several segments per year, or multi-year chains)
// pack: the word's sum and negative sum share one 32-bit integer (single-segment mode only)
__host__ __device__ inline size_t fast_warp_bytes(int seg_words, int ev_cap, bool two_halves, bool pack)
{
    const int halves = two_halves ? 2 : 1;
    size_t b = sizeof(FastWarpSmem) + (two_halves ? sizeof(uint32_t) * FAST_PEND_CAP : 0) +
               sizeof(uint32_t) * (size_t)halves * ev_cap +                          // event lists
               (pack ? 2 : 3) * sizeof(int32_t) * (size_t)fast_words_alloc(seg_words, two_halves);   // per word: {sum, negative sum, list head}
    return (b + 15) & ~(size_t)15;
}

__host__ __device__ inline int fast_lmax_alloc(int Wd) { return 32 * ((Wd + 31) / 32); }

__host__ __device__ inline size_t fast_block_bytes(int Wd, bool load16)
{
    size_t b = (load16 ? 2 : 4) * (size_t)Wd * 32 + sizeof(int32_t) * (size_t)fast_lmax_alloc(Wd);   // load curve + word maxima
    b += 32 * (sizeof(int32_t) + 3 * sizeof(float) + sizeof(uint32_t)) + 64 * sizeof(int32_t);        // unit tables
    return (b + 15) & ~(size_t)15;
}

size_t seq_fast_smem_bytes(int Wd, int seg_words, int warps_per_block, int ev_cap, bool two_halves, bool load16, bool pack)
{
    return fast_block_bytes(Wd, load16) + (size_t)warps_per_block * fast_warp_bytes(seg_words, ev_cap, two_halves, pack);
}

int seq_fast_max_threads(bool two_halves) { return FAST_MAX_THREADS(two_halves); }

// predicated shared-memory updates (inline PTX keeps them branch-free in SASS)
__device__ __forceinline__ void red_add_shared_if(uint32_t saddr, int v, bool p)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p red.shared.add.s32 [%0], %1;\n}\n"
                 :: "r"(saddr), "r"(v), "r"((int)p) : "memory");
}

// push event `ent` (list link filled in) onto the list of its word: head <- idx1, entry.next <- old head
__device__ __forceinline__ void list_push_shared_if(uint32_t head_saddr, uint32_t ev_saddr, uint32_t idx1, uint32_t ent, bool p)
{
    asm volatile("{\n .reg .pred p;\n .reg .b32 nx;\n setp.ne.b32 p, %4, 0;\n @p atom.shared.exch.b32 nx, [%0], %2;\n"
                 " @p shl.b32 nx, nx, 20;\n @p or.b32 nx, nx, %3;\n @p st.shared.b32 [%1], nx;\n}\n"
                 :: "r"(head_saddr), "r"(ev_saddr), "r"(idx1), "r"(ent), "r"((int)p) : "memory");
}


// Single-segment scatter of one event (the whole chain is one timeline segment).  If hour `hs` lies in the
// year: add `delta` to the word's sum (and to its negative sum when the event takes the unit down, i.e. when
// s0i == qodd), push the event onto the word's list and count it.  `wa` = shared address of the word record
// {sum, negative sum, head}; the list slot at `slot_m8 + 8` is private to this lane.  ptxas turns predicated
// shared atomics into branches, so the atomics are unconditional instead: an out-of-year event adds 0 to the
// record of word `lane` (address `wlane`; only ever touched by atomics while a segment is being filled) and
// exchanges with its own, never linked, list slot.
__device__ __forceinline__ void scatter_event_single(uint32_t hs, uint32_t H, uint32_t wa, uint32_t wlane, int delta, uint32_t s0i,
                                                     uint32_t qodd, uint32_t idx1, uint32_t slot_m8, uint32_t ent, unsigned int &n_events)
{
    asm volatile("{\n .reg .pred p, pn;\n .reg .b32 nx, d0, d1, hr, hx;\n"
                 " setp.lt.u32 p, %1, %2;\n"
                 " setp.eq.and.u32 pn, %6, %7, p;\n"
                 " selp.b32 d0, %5, 0, p;\n"
                 " selp.b32 d1, %5, 0, pn;\n"
                 " selp.b32 hr, %3, %4, p;\n"
                 " selp.b32 hx, %3, %9, p;\n"
                 " red.shared.add.s32 [hr], d0;\n"
                 " red.shared.add.s32 [hr+4], d1;\n"
                 " atom.shared.exch.b32 nx, [hx+8], %8;\n"
                 " mad.lo.u32 nx, nx, 1048576, %10;\n"
                 " st.shared.b32 [%9+8], nx;\n"
                 " @p add.u32 %0, %0, 1;\n}\n"
                 : "+r"(n_events)
                 : "r"(hs), "r"(H), "r"(wa), "r"(wlane), "r"(delta), "r"(s0i), "r"(qodd), "r"(idx1), "r"(slot_m8), "r"(ent)
                 : "memory");
}

// Packed variant: the record is {sum + 2^K * negative sum, head}; `delta` already carries both fields
// (c for an up event, -c * (1 + 2^K) for a down event), so one add serves both sums.  The packed layout keeps
// two arrays, values and list heads (`wval`, `whead` = their shared addresses; 4-byte stride, so the 32 lanes of an
// atomic spread over all 32 banks): the word of hour `hs` is element hs >> 5 (one shift, two multiply-adds on the FMA
// pipe).  An out-of-year event does both atomics on the lane's own list slot (`slot`,
// never linked and overwritten by the store below), so the delta needs no select.
__device__ __forceinline__ void scatter_event_packed(uint32_t hs, uint32_t H, uint32_t wval, uint32_t whead, int delta,
                                                     uint32_t idx1, uint32_t slot, uint32_t ent, unsigned int &n_events)
{
    asm volatile("{\n .reg .pred p;\n .reg .b32 nx, hr, hx, w;\n"
                 " setp.lt.u32 p, %1, %2;\n"
                 " shr.u32 w, %1, 5;\n"
                 " mad.lo.u32 hr, w, 4, %3;\n"
                 " mad.lo.u32 hx, w, 4, %8;\n"
                 " selp.b32 hr, hr, %6, p;\n"
                 " selp.b32 hx, hx, %6, p;\n"
                 " red.shared.add.s32 [hr], %4;\n"
                 " atom.shared.exch.b32 nx, [hx], %5;\n"
                 " mad.lo.u32 nx, nx, 1048576, %7;\n"
                 " st.shared.b32 [%6], nx;\n"
                 " @p add.u32 %0, %0, 1;\n}\n"
                 : "+r"(n_events)
                 : "r"(hs), "r"(H), "r"(wval), "r"(delta), "r"(idx1), "r"(slot), "r"(ent), "r"(whead)
                 : "memory");
}

template <bool kDisc, bool kTwo, bool kPack>
__global__ void __launch_bounds__(FAST_MAX_THREADS(kTwo), 1) seq_fast_kernel(const SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int Hpad = a.Wd * 32;
    const bool load16 = a.load16 != 0;
    int32_t *s_load32 = reinterpret_cast<int32_t *>(smem_raw);
    short *s_load16 = reinterpret_cast<short *>(smem_raw);
    int32_t *s_lmax = reinterpret_cast<int32_t *>(smem_raw + (load16 ? 2 : 4) * (size_t)Hpad);
    int32_t *s_cap = s_lmax + fast_lmax_alloc(a.Wd);
    float *s_mup = reinterpret_cast<float *>(s_cap + 32);
    float *s_mdn = s_mup + 32;
    float *s_ispan = s_mdn + 32;         // Philox blocks a unit needs per hour of horizon (a block = 4 durations = 2 up/down cycles)
    uint32_t *s_thr = reinterpret_cast<uint32_t *>(s_ispan + 32);
    int32_t *s_dcap = reinterpret_cast<int32_t *>(s_thr + 32);     // [(unit << 1) | sign]: signed capacity step of an event-list entry
    constexpr bool two_halves = kTwo;     // the ring needs its second half (several segments per year or multi-year chains)
    const int halves = two_halves ? 2 : 1;
    const int ev_cap = a.ev_cap;
    unsigned char *wbase = smem_raw + fast_block_bytes(a.Wd, load16) + (size_t)warp * fast_warp_bytes(a.seg_words, ev_cap, two_halves, kPack);
    FastWarpSmem *ws = reinterpret_cast<FastWarpSmem *>(wbase);
    uint32_t *pend = reinterpret_cast<uint32_t *>(wbase + sizeof(FastWarpSmem));  // (hour << 6) | (unit << 1) | (delta > 0)
    uint32_t *evl = pend + (two_halves ? FAST_PEND_CAP : 0);                      // [halves][ev_cap]: (next << 20) | (hour in segment << 6) | (unit << 1) | sign; the events of a
                                                                                  // 32-hour word form a linked list (next = 1 + index, 0 = end)
    const int seg_slots = a.seg_words * 32;
    const int ring_words = fast_words_alloc(a.seg_words, two_halves);
    // one record per 32-hour word: sum of deltas, sum of negative deltas, 1 + index of its newest event (0 = none).
    // kPack: the two sums share one integer v = (sum + 2^(K-1)) + 2^K * neg: |sum| < 2^(K-1), so the biased low field
    // stays in [0, 2^K) and both fields decode with one mask / one shift; the host proves neg > -2^(31-K) for any
    // list that fits ev_cap.
    static_assert(!(kPack && kTwo), "packed word records are a single-segment feature");
    constexpr int RS = kPack ? 2 : 3;
    const int pk = a.pack_shift;
    const int pk_bias = kPack ? (1 << (pk - 1)) : 0, pk_mask = (1 << pk) - 1;
    int32_t *wtab = reinterpret_cast<int32_t *>(evl + (size_t)halves * ev_cap);
    // kPack: structure of arrays -- packed values [ring_words], then list heads [ring_words]; otherwise records of three
    int32_t *whd = wtab + ring_words;
    auto word_sums = [&](int i, int &sm, int &ng) {
        if constexpr (kPack) { const int v = wtab[i]; sm = (v & pk_mask) - pk_bias; ng = v >> pk; }
        else { sm = wtab[RS * i]; ng = wtab[RS * i + 1]; }
    };
#define WSUM(i) wtab[3 * (i)]
#define WNEG(i) wtab[3 * (i) + 1]
#define WHEAD(i) (kPack ? reinterpret_cast<uint32_t *>(whd)[i] : reinterpret_cast<uint32_t *>(wtab)[3 * (i) + 2])

    for (int i = threadIdx.x; i < Hpad; i += blockDim.x) {
        if (load16) s_load16[i] = (short)a.load[i]; else s_load32[i] = a.load[i];
    }
    if constexpr (kTwo) {
        for (int i = threadIdx.x; i < a.Wd; i += blockDim.x) s_lmax[i] = a.lmax[i];
    } else {
        // single segment: word w is word w % wpl of its lane's run; the evaluation scan keeps a running sum of the
        // *biased* packed sums, so the bias of the words before w in the run is folded into the table here.  Padding
        // words can never raise a flag.
        const int wpl0 = (a.Wd + 31) >> 5;
        for (int i = threadIdx.x; i < fast_lmax_alloc(a.Wd); i += blockDim.x)
            s_lmax[i] = i < a.Wd ? a.lmax[i] + pk_bias * (i % wpl0) : -(1 << 30);
    }
    if (threadIdx.x < 32) {
        const bool v = threadIdx.x < a.U;
        s_cap[threadIdx.x] = v ? a.cap[threadIdx.x] : 0;
        s_dcap[2 * threadIdx.x] = v ? -a.cap[threadIdx.x] : 0;
        s_dcap[2 * threadIdx.x + 1] = v ? a.cap[threadIdx.x] : 0;
        s_mup[threadIdx.x] = v ? __fmul_rn(a.mttf[threadIdx.x], 16777216.0f) : 1.0f;
        s_mdn[threadIdx.x] = v ? __fmul_rn(a.mttr[threadIdx.x], 16777216.0f) : 1.0f;
        s_thr[threadIdx.x] = v ? a.for_thr[threadIdx.x] : 0u;
        s_ispan[threadIdx.x] = v ? __fdividef(0.5f, a.mttf[threadIdx.x] + a.mttr[threadIdx.x]) : 0.f;
    }
    for (int i = lane; i < RS * ring_words; i += 32) wtab[i] = (kPack && i < ring_words) ? pk_bias : 0;
    const uint32_t wtab_s = (uint32_t)__cvta_generic_to_shared(wtab);
    const uint32_t whd_s = wtab_s + 4u * (uint32_t)ring_words;
    const uint32_t jobmap_s = (uint32_t)__cvta_generic_to_shared(ws->jobmap);
    const uint32_t wlane_s = wtab_s + 4u * RS * (uint32_t)min(lane, a.seg_words - 1);     // where this lane's out-of-year events add 0
    __syncthreads();
    auto load_at = [&](int i) -> int { return load16 ? (int)s_load16[i] : s_load32[i]; };

    // long-lived per-warp sums live in shared memory (registers are what limits the resident warps)
    if (lane < 8) { ws->acc[lane] = 0ull; ws->diag[lane] = 0u; }
    __syncwarp();
    unsigned int n_events = 0;

    const bool unit_valid = lane < a.U;

    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    // single-segment variant: one year per chain, one segment per year -- the loop bounds and the ring geometry below
    // are compile-time facts there
    const int n_ypc = kTwo ? a.ypc : 1, n_seg = kTwo ? a.nseg : 1;
    const int chain_end_h = n_ypc * a.H;

    for (long long cl = gw; cl < a.nchains; cl += nw) {
        const unsigned long long chain = (unsigned long long)(a.chain_base + cl);
        uint32_t nb = 0;              // next Philox block of unit `lane`
        uint32_t s0mask = 0;          // initial states (bit u = UP)
        int pend_cnt = 0;
        int capacity = 0;
        int ring = 0;                 // ring half that holds the current segment
        int ev_cnt0 = 0, ev_cnt1 = 0; // events in the list of ring half 0 / 1 (warp-uniform)
        // MATLAB discretisation: a unit that fails after d whole hours is DOWN from hour d+1 (seq_mcsampling.m:63)
        ws->t_run[lane] = kDisc ? (1ull << PSRA_TICK_SHIFT) : 0ull;
        __syncwarp();
        bool init_wave = kTwo;          // single-segment mode: block 0 of every unit belongs to the static phase

        for (int y = 0; y < n_ypc; y++) {
            unsigned int lolh = 0, entries = 0;
            long long ens_lane = 0;
            bool year_bad = false;        // warp-uniform
            for (int seg = 0; seg < n_seg; seg++, ring ^= 1) {
                const int seg_h0 = kTwo ? seg * seg_slots : 0;
                const int seg_h1 = kTwo ? min(a.H, seg_h0 + seg_slots) : a.H;
                const int abs0 = y * a.H + seg_h0, abs1 = y * a.H + seg_h1;   // chain-relative hours
                // the segment after this one (same year or first of the next year) lives in the other half
                const int nxt_h0 = (seg + 1 < n_seg) ? seg_h0 + seg_slots : 0;
                const int abs2 = kTwo ? min(chain_end_h, abs1 + min(seg_slots, a.H - nxt_h0)) : a.H;
                const unsigned long long seg_end_t = (unsigned long long)abs1 << PSRA_TICK_SHIFT;
                const unsigned long long nxt_end_t = (unsigned long long)abs2 << PSRA_TICK_SHIFT;
                // ring geometry: an event `rel` hours after abs0 belongs to the current half when rel < len_cur,
                // otherwise to the other half (which holds the next segment from its hour 0)
                const uint32_t len_cur = (uint32_t)(seg_h1 - seg_h0);
                const uint32_t ring_len = (uint32_t)(abs2 - abs0);
                const int ringc = kTwo ? ring : 0;
                const int wbase_cur = ringc * a.seg_words, wbase_nxt = (ringc ^ 1) * a.seg_words;
                uint32_t *ev_cur = evl + (size_t)ringc * ev_cap, *ev_nxt = kTwo ? evl + (size_t)(ringc ^ 1) * ev_cap : evl;
                int cnt_cur = kTwo ? (ring ? ev_cnt1 : ev_cnt0) : 0, cnt_nxt = kTwo ? (ring ? ev_cnt0 : ev_cnt1) : 0;
                const uint32_t evcur_s = (uint32_t)__cvta_generic_to_shared(ev_cur), evnxt_s = (uint32_t)__cvta_generic_to_shared(ev_nxt);

                // ---- far-future events that now fall into the next segment's half
                if (two_halves && pend_cnt) {
                    int outc = 0;
                    for (int base = 0; base < pend_cnt; base += 32) {
                        const int i = base + lane;
                        const bool v = i < pend_cnt;
                        const uint32_t e = v ? pend[i] : 0u;
                        const int hs = (int)(e >> 6);
                        const bool take = v && hs < abs2;
                        const uint32_t tm = __ballot_sync(0xffffffffu, take);
                        if (take) {                       // pending events are beyond abs1: next half
                            const int c = s_cap[(e >> 1) & 31];
                            const uint32_t hseg = (uint32_t)(hs - abs1);
                            atomicAdd(&WSUM(wbase_nxt + (hseg >> 5)), (e & 1u) ? c : -c);
                            if (!(e & 1u)) atomicAdd(&WNEG(wbase_nxt + (hseg >> 5)), -c);
                            const int pos = cnt_nxt + __popc(tm & lt_mask);
                            if (pos < ev_cap) {
                                const uint32_t nx = atomicExch(&WHEAD(wbase_nxt + (hseg >> 5)), (uint32_t)(pos + 1));
                                ev_nxt[pos] = (nx << 20) | (hseg << 6) | (e & 63u);
                            }
                        }
                        cnt_nxt += __popc(tm);
                        const bool keep = v && !take;
                        const uint32_t km = __ballot_sync(0xffffffffu, keep);
                        __syncwarp();
                        if (keep) pend[outc + __popc(km & lt_mask)] = e;
                        outc += __popc(km);
                    }
                    pend_cnt = outc;
                }

                // ---- static phase (single segment): lane = unit, blocks 0 .. static_blocks-1 of every unit, no scheduling.
                //      Every unit needs these blocks anyway (the host picks the count from the expected demand), so
                //      the job mapping, the ballots and the segmented scan of the waves below are skipped for them.
                if constexpr (!kTwo) {
                    unsigned long long t = kDisc ? (1ull << PSRA_TICK_SHIFT) : 0ull;
                    bool s0u = true;
                    const float mup = s_mup[lane], mdn = s_mdn[lane];
                    const int capu = s_cap[lane];
                    const int pk_dn = -capu - (capu << pk);
                    // list slots as in a wave of J = U jobs: unit u owns cnt + q U + u, q = 0..3; a lane without a unit keeps
                    // one dummy slot behind them
                    const uint32_t Uu = (uint32_t)a.U, idx_step = unit_valid ? Uu : 0u;
                    for (int k = 0; k < a.static_blocks; k++) {
                        uint32_t x[4];
                        philox4x32_10_rk((uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)lane, (uint32_t)k, a.rk, x);
                        if (k == 0) s0u = !(a.init_mode == PSRA_INIT_STATIONARY && x[0] < s_thr[lane]);   // draw 0 = initial state
                        const float m_a = s0u ? mdn : mup;      // draws 0, 2 of a block: state s0^1
                        const float m_b = s0u ? mup : mdn;      // draws 1, 3: state s0
                        const unsigned long long p1 = (k == 0) ? 0ull : dur_ticks_disc(m_a, x[0], !s0u, kDisc);
                        const unsigned long long p2 = p1 + dur_ticks_disc(m_b, x[1], s0u, kDisc);
                        const unsigned long long p3 = p2 + dur_ticks_disc(m_a, x[2], !s0u, kDisc);
                        const unsigned long long p4 = p3 + dur_ticks_disc(m_b, x[3], s0u, kDisc);
                        const bool room = cnt_cur + 3 * a.U + 32 <= ev_cap;          // warp-uniform
                        const unsigned long long bm1 = (unit_valid && room) ? t - 1ull : (0x00800000ull << 32);
                        uint32_t idx1 = (uint32_t)(room ? cnt_cur : 0) + (uint32_t)lane + 1u + (unit_valid ? 0u : 3u * Uu);
                        const uint32_t s0i = s0u ? 1u : 0u;
                        const int delta_a = s0u ? capu : -capu;
                        const int pk_a = s0u ? capu : pk_dn, pk_b = s0u ? pk_dn : capu;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            unsigned long long tm1 = bm1 + (q == 0 ? p1 : q == 1 ? p2 : q == 2 ? p3 : p4);
                            if (q == 0 && k == 0) tm1 = 0x00800000ull << 32;
                            const uint32_t hs = __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
                            const uint32_t ent = (hs << 6) + (((uint32_t)lane << 1) | (s0i ^ (uint32_t)(q & 1)));
                            if constexpr (kPack)
                                scatter_event_packed(hs, (uint32_t)a.H, wtab_s, whd_s, (q & 1) ? pk_b : pk_a, idx1,
                                                     evcur_s + 4u * idx1 - 4u, ent, n_events);
                            else
                                scatter_event_single(hs, (uint32_t)a.H, wtab_s + 12u * (hs >> 5), wlane_s, (q & 1) ? -delta_a : delta_a, s0i,
                                                     (uint32_t)(q & 1), idx1, evcur_s + 4u * idx1 - 12u, ent, n_events);
                            idx1 += idx_step;
                        }
                        t += p4;
                        if (room) cnt_cur += 4 * a.U; else cnt_cur = ev_cap + 1;    // reported as PSRA_E_OVERFLOW below
                        __syncwarp();                            // the dummy slots of unit-less lanes are the next block's slots
                    }
                    ws->t_run[lane] = t;
                    nb = (uint32_t)a.static_blocks;
                    s0mask = __ballot_sync(0xffffffffu, unit_valid && s0u);
                    int cp = (unit_valid && s0u) ? capu : 0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) cp += __shfl_xor_sync(0xffffffffu, cp, d);
                    capacity = cp;
                    init_wave = false;
                    if (lane == 0) { ws->diag[1] += (unsigned int)(a.U * a.static_blocks); ws->diag[0] += (unsigned int)a.static_blocks; }
                    __syncwarp();
                }

                // ---- waves: one round of <= 32 (unit, block) jobs each, until every unit covers the segment
                while (true) {
                    const unsigned long long tlast = ws->t_run[lane];
                    const bool is_short = unit_valid && tlast <= seg_end_t;
                    if (!__any_sync(0xffffffffu, is_short)) break;
                    int want = 0;                      // blocks towards the end of the NEXT segment
                    if (kTwo && init_wave) want = unit_valid ? 1 : 0;
                    else if (kTwo ? (unit_valid && tlast <= nxt_end_t) : is_short) {    // single segment: next end == this end
                        const float rem_h = (float)(int)(((kTwo ? nxt_end_t : seg_end_t) - tlast) >> PSRA_TICK_SHIFT);
                        want = min(FAST_NB_MAX, 1 + (int)(rem_h * s_ispan[lane]));
                    }
                    // exclusive prefix / total of a per-lane count in 0..7 from three ballots
                    auto count_scan = [&](int n, int &excl, int &total) {
                        const uint32_t b0 = __ballot_sync(0xffffffffu, n & 1), b1 = __ballot_sync(0xffffffffu, n & 2),
                                       b2 = __ballot_sync(0xffffffffu, n & 4);
                        excl = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
                        total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                    };
                    int n_m = is_short ? want : 0;                   // mandatory: units short of this segment
                    int off_m, J1;
                    count_scan(n_m, off_m, J1);
                    if (J1 < 32 && !init_wave) {       // spare lanes: one more block of slack for the short units
                        const bool extra = is_short && want < FAST_NB_MAX;
                        const uint32_t bx = __ballot_sync(0xffffffffu, extra);
                        const int Jx = J1 + __popc(bx);
                        if (Jx <= 32) { n_m += extra ? 1 : 0; off_m += __popc(bx & lt_mask); J1 = Jx; }
                    }
                    int n_u, off, J;
                    if (J1 >= 32 || init_wave || !two_halves) {     // truncate; the remaining demand is served by the next wave
                        off = off_m;
                        n_u = max(0, min(n_m, 32 - off));
                        J = min(J1, 32);
                    } else {                           // spare lanes pre-generate for the next segment
                        const int n_o = (!is_short && pend_cnt <= FAST_PEND_CAP / 2) ? want : 0;
                        int off_o, J2;
                        count_scan(n_o, off_o, J2);
                        off_o += J1;
                        off = is_short ? off_m : off_o;
                        n_u = is_short ? n_m : max(0, min(n_o, 32 - off_o));
                        J = min(32, J1 + J2);
                        if (lane == 0) ws->diag[2] += (unsigned int)(J - J1);
                    }
                    if (lane == 0) { ws->diag[1] += (unsigned int)J; ws->diag[0]++; }
                    // job map: entries off .. off + n_u - 1 name this unit (off + n_u <= 32).  Plain 32-bit shared addresses
                    // and predicated byte stores: through the generic pointer ptxas rebuilds the shared window base
                    // (S2R SR_CgaCtaId, LEA) for every store
                    {
                        const uint32_t jm_w = jobmap_s + (uint32_t)off;
                        asm volatile("{\n .reg .pred p0, p1, p2, p3;\n"
                                     " setp.gt.s32 p0, %2, 0;\n setp.gt.s32 p1, %2, 1;\n setp.gt.s32 p2, %2, 2;\n setp.gt.s32 p3, %2, 3;\n"
                                     " @p0 st.shared.u8 [%0], %1;\n @p1 st.shared.u8 [%0+1], %1;\n"
                                     " @p2 st.shared.u8 [%0+2], %1;\n @p3 st.shared.u8 [%0+3], %1;\n}\n"
                                     :: "r"(jm_w), "r"(lane), "r"(n_u) : "memory");
                    }
                    __syncwarp();
                    {
                        const bool act = lane < J;               // lane = job
                        int u;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(u) : "r"(jobmap_s + (uint32_t)lane) : "memory");
                        u = act ? u : 0;
                        const int offu = __shfl_sync(0xffffffffu, off, u);
                        const int nu = __shfl_sync(0xffffffffu, n_u, u);
                        const uint32_t b = __shfl_sync(0xffffffffu, nb, u) + (uint32_t)(lane - offu);
                        const bool is_last = act && (lane - offu) == nu - 1;

                        uint32_t x[4];
                        philox4x32_10_rk((uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)u, b, a.rk, x);
                        const bool blk0 = kTwo && b == 0u;       // block 0 of a stream: draw 0 is the initial state
                        bool s0u;
                        if (blk0) s0u = !(a.init_mode == PSRA_INIT_STATIONARY && x[0] < s_thr[u]);
                        else s0u = (s0mask >> u) & 1u;
                        const float mup = s_mup[u], mdn = s_mdn[u];
                        const float m_a = s0u ? mdn : mup;      // draws 0, 2 of a block: state s0^1
                        const float m_b = s0u ? mup : mdn;      // draws 1, 3: state s0
                        const unsigned long long p1 = blk0 ? 0ull : dur_ticks_disc(m_a, x[0], !s0u, kDisc);
                        const unsigned long long p2 = p1 + dur_ticks_disc(m_b, x[1], s0u, kDisc);
                        const unsigned long long p3 = p2 + dur_ticks_disc(m_a, x[2], !s0u, kDisc);
                        const unsigned long long p4 = p3 + dur_ticks_disc(m_b, x[3], s0u, kDisc);
                        const unsigned long long tot = act ? p4 : 0ull;
                        unsigned long long inc2 = tot;
#pragma unroll
                        for (int d = 1; d <= 2; d <<= 1) {       // runs are <= FAST_NB_MAX = 4 lanes long
                            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc2, d);
                            if (lane - d >= offu) inc2 += o;
                        }
                        const unsigned long long base_t = ws->t_run[u] + (inc2 - tot);
                        __syncwarp();
                        if (is_last) ws->t_run[u] = base_t + tot;

                        const int cu = s_cap[u];
                        const int delta_a = s0u ? cu : -cu;      // draws 0, 2 toggle the unit back to s0
                        if (blk0 && act) s0mask = s0u ? 1u : 0u;   // init wave (job lane == unit lane), ballot below
                        // hour of an event at tick T: ceil(T / 2^24) - 1 = (T - 1) >> 24 (fits 32 bits: T < 2^56)
                        if constexpr (!kTwo) {
                            // whole chain = one segment: hour in segment = hour in year, ring = year.  Job lane j of a wave
                            // of J jobs owns the list slots cnt + q*J + j, q = 0..3 (holes of out-of-year events are never linked)
                            // lanes without a job park their dummy exchange in "their" slot too, i.e. up to slot cnt + 3 J + 31
                            const bool room = cnt_cur + 3 * J + 32 <= ev_cap;         // warp-uniform; implies cnt + 4 J <= ev_cap
                            // lanes without a job (and a full list) are parked far beyond the year
                            const unsigned long long bm1 = (act && room) ? base_t - 1ull : (0x00800000ull << 32);
                            const uint32_t s0i = s0u ? 1u : 0u;
                            // a lane without a job keeps one dummy slot behind the wave's 4 J slots for all four events
                            uint32_t idx1 = (uint32_t)(room ? cnt_cur : 0) + (uint32_t)lane + 1u + (act ? 0u : 3u * (uint32_t)J);   // full list: stay in bounds (ev_cap >= 128)
                            const uint32_t idx_step = act ? (uint32_t)J : 0u;
                            const int dn_pk = -cu - (cu << pk);                       // kPack: a down event in both fields
                            const int pk_a = s0u ? cu : dn_pk, pk_b = s0u ? dn_pk : cu;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                unsigned long long tm1 = bm1 + (q == 0 ? p1 : q == 1 ? p2 : q == 2 ? p3 : p4);
                                const uint32_t hs = __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
                                const int delta = (q & 1) ? -delta_a : delta_a;
                                // down events: q even when the stream starts DOWN (s0i == 0), q odd when it starts UP
                                const uint32_t ent = (hs << 6) + (((uint32_t)u << 1) | (s0i ^ (uint32_t)(q & 1)));
                                if constexpr (kPack)
                                    scatter_event_packed(hs, (uint32_t)a.H, wtab_s, whd_s, (q & 1) ? pk_b : pk_a, idx1,
                                                         evcur_s + 4u * idx1 - 4u, ent, n_events);
                                else
                                    scatter_event_single(hs, (uint32_t)a.H, wtab_s + 12u * (hs >> 5), wlane_s, delta, s0i, (uint32_t)(q & 1), idx1,
                                                         evcur_s + 4u * idx1 - 12u, ent, n_events);
                                idx1 += idx_step;
                            }
                            cnt_cur += 4 * J;
                            if (!room) cnt_cur = ev_cap + 1;                          // reported as PSRA_E_OVERFLOW below
                        } else {
                            const unsigned long long bm1 = base_t - 1ull;
    #pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const unsigned long long tm1 = bm1 + (q == 0 ? p1 : q == 1 ? p2 : q == 2 ? p3 : p4);
                                const uint32_t hs = __funnelshift_r((uint32_t)tm1, (uint32_t)(tm1 >> 32), PSRA_TICK_SHIFT);
                                const bool valid = act && !(blk0 && q == 0);
                                const uint32_t rel = hs - (uint32_t)abs0;          // >= 0: events are never generated backwards
                                const int delta = (q & 1) ? -delta_a : delta_a;
                                const bool in_ring = valid && rel < ring_len;      // ring_len stops at the chain end
                                const bool in_cur = in_ring && rel < len_cur;
                                const bool in_nxt = in_ring && !in_cur;
                                const uint32_t hseg = in_cur ? rel : rel - len_cur;  // hour within its segment
                                const uint32_t wa = wtab_s + 12u * ((in_cur ? (uint32_t)wbase_cur : (uint32_t)wbase_nxt) + (hseg >> 5));
                                red_add_shared_if(wa, delta, in_ring);
                                red_add_shared_if(wa + 4u, delta, in_ring && delta < 0);
                                const uint32_t ent = (hseg << 6) | ((uint32_t)u << 1) | (delta > 0 ? 1u : 0u);
                                const uint32_t mc = __ballot_sync(0xffffffffu, in_cur);
                                {
                                    const uint32_t pos = (uint32_t)cnt_cur + __popc(mc & lt_mask);
                                    list_push_shared_if(wa + 8u, evcur_s + 4u * pos, pos + 1u, ent, in_cur && pos < (uint32_t)ev_cap);
                                }
                                cnt_cur += __popc(mc);
                                if (two_halves) {
                                    const uint32_t mn = __ballot_sync(0xffffffffu, in_nxt);
                                    const uint32_t pos = (uint32_t)cnt_nxt + __popc(mn & lt_mask);
                                    list_push_shared_if(wa + 8u, evnxt_s + 4u * pos, pos + 1u, ent, in_nxt && pos < (uint32_t)ev_cap);
                                    cnt_nxt += __popc(mn);
                                }
                                const bool inhor = valid && hs < (uint32_t)chain_end_h;
                                if (two_halves) {                              // events beyond the ring (a single segment has none)
                                    const bool pnd = inhor && !in_ring;
                                    const uint32_t pm = __ballot_sync(0xffffffffu, pnd);
                                    if (pm) {
                                        const int pos = pend_cnt + __popc(pm & lt_mask);
                                        if (pnd && pos < FAST_PEND_CAP)
                                            pend[pos] = (hs << 6) | ((uint32_t)u << 1) | (delta > 0 ? 1u : 0u);
                                        pend_cnt += __popc(pm);
                                    }
                                }
                                n_events += inhor ? 1u : 0u;
                            }
                        }
                        __syncwarp();
                    }
                    nb += (uint32_t)n_u;
                    if (init_wave) {
                        s0mask = __ballot_sync(0xffffffffu, unit_valid && (s0mask & 1u));
                        int cp = (unit_valid && ((s0mask >> lane) & 1u)) ? s_cap[lane] : 0;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) cp += __shfl_xor_sync(0xffffffffu, cp, d);
                        capacity = cp;
                        init_wave = false;
                    }
                }
                if (lane == 0) ws->diag[4] = max(ws->diag[4], (unsigned int)max(cnt_cur, cnt_nxt));
                bool list_ok = true;
                if (cnt_cur > ev_cap || cnt_nxt > ev_cap) {
                    // the event list is full.  Single segment (a chain = one year): the year goes to the redo list and the
                    // host replays it with the generic kernel; ring variant: the host repeats the whole call with the
                    // generic kernel (the earlier years of the chain are already in the sums)
                    if constexpr (kTwo) { if (lane == 0) atomicExch(&a.acc[ACC_OVERFLOW], 3ull); }
                    else year_bad = true;
                    cnt_cur = min(cnt_cur, ev_cap); cnt_nxt = min(cnt_nxt, ev_cap);
                    list_ok = false;                             // the links may be garbage: do not walk the lists
                }
                if (pend_cnt > FAST_PEND_CAP) {                  // reported as PSRA_E_OVERFLOW (never seen in practice)
                    if (lane == 0) atomicExch(&a.acc[ACC_OVERFLOW], 2ull);
                    pend_cnt = FAST_PEND_CAP;
                }
                __syncwarp();

                // ---- evaluation of the current half: lane = run of `wpl` consecutive words
                const int nwords = kTwo ? (seg_h1 - seg_h0 + 31) >> 5 : a.Wd;
                const int wpl = (nwords + 31) >> 5;
                const int wb = lane * wpl;
                int loc = 0, lmin = INT_MAX;
                if constexpr (kTwo) {
                    for (int k = 0; k < wpl; k++) {
                        const int w = wb + k;
                        if (w < nwords) {
                            int sm, ng;
                            word_sums(wbase_cur + w, sm, ng);
                            lmin = min(lmin, loc + ng - s_lmax[seg * a.seg_words + w]);
                            loc += sm;
                        }
                    }
                } else {
                    // every lane owns wpl records (the table is padded with neutral ones): no bounds checks, and the
                    // running sum keeps the packing bias (s_lmax carries the same bias, see above)
                    const int32_t *wp = wtab + (kPack ? 1 : RS) * wb;
                    const int32_t *lp = s_lmax + wb;
                    for (int k = 0; k < wpl; k++) {
                        int smb, ng;
                        if constexpr (kPack) { const int v = wp[k]; smb = v & pk_mask; ng = v >> pk; }
                        else { smb = wp[RS * k]; ng = wp[RS * k + 1]; }
                        lmin = min(lmin, loc + ng - lp[k]);
                        loc += smb;
                    }
                    loc -= wpl * pk_bias;
                }
                const int incl = warp_incl_scan(loc, lane);
                const int cs_lane = capacity + incl - loc;       // capacity entering the lane's run
                const bool flagged = (lmin != INT_MAX) && (cs_lane + lmin < 0);
                uint32_t fm = list_ok ? __ballot_sync(0xffffffffu, flagged) : 0u;
                if (lane == 0) ws->diag[3] += (unsigned int)__popc(fm);
                while (fm) {                                     // rare: a run that may contain loss of load
                    const int src = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const int c_run = __shfl_sync(0xffffffffu, cs_lane, src);
                    // lane k < wpl looks at word k of the run: capacity entering it and the per-word flag
                    int c_word = c_run;
                    bool wflag = false;
                    {
                        const int wq = src * wpl + lane;
                        const bool mine = lane < wpl && wq < nwords;
                        int ws_l = 0, wn_l = 0;
                        if (mine) word_sums(wbase_cur + wq, ws_l, wn_l);
                        int ws_i = ws_l;                                  // a run has <= 16 words (seg_words <= 512)
#pragma unroll
                        for (int d = 1; d < 16; d <<= 1) {
                            const int o = __shfl_up_sync(0xffffffffu, ws_i, d);
                            if (lane >= d) ws_i += o;
                        }
                        c_word += ws_i - ws_l;
                        if (mine) wflag = c_word + wn_l < (kTwo ? s_lmax[seg * a.seg_words + wq] : s_lmax[wq] - pk_bias * lane);
                    }
                    uint32_t wm = __ballot_sync(0xffffffffu, wflag);
                    while (wm) {                                 // resolve the word hour by hour, lane = hour
                        const int k = __ffs(wm) - 1;
                        wm &= wm - 1;
                        const int wq = src * wpl + k;
                        const int c_in = __shfl_sync(0xffffffffu, c_word, k);
                        int d = 0;                               // delta of hour `lane` of the word: walk its event list
                        for (uint32_t i = WHEAD(wbase_cur + wq); i; ) {
                            const uint32_t e = ev_cur[i - 1];
                            if (((e >> 6) & 31u) == (uint32_t)lane) d += s_dcap[e & 63u];
                            i = e >> 20;
                        }
                        const int c = c_in + warp_incl_scan(d, lane);
                        const int hy0 = seg_h0 + wq * 32;
                        const int L = load_at(hy0 + lane);
                        const bool lol = c < L;                  // PSA.jl:253 strict
                        const uint32_t mask = __ballot_sync(0xffffffffu, lol);
                        if (mask) {
                            const uint32_t prev = (hy0 > 0 && c_in < load_at(hy0 - 1)) ? 1u : 0u;
                            lolh += __popc(mask);
                            entries += __popc(mask & ~((mask << 1) | prev));   // calnlc.m:22-34
                            if (lol) {
                                ens_lane += (long long)(L - c);
                                if (a.fail) atomicAdd(&a.fail[hy0 + lane], 1u);
                            }
                        }
                    }
                }
                capacity += __shfl_sync(0xffffffffu, incl, 31);
                __syncwarp();
                // clear the evaluated half: word sums to zero, event list empty
                if constexpr (!kTwo) {      // the table is 16-byte aligned and padded to a multiple of four records
                    uint4 *t4 = reinterpret_cast<uint4 *>(wtab);
                    if constexpr (kPack) {          // values back to the bias, heads to "empty" (both arrays are 16-byte aligned)
                        uint4 *h4 = reinterpret_cast<uint4 *>(whd);
                        const int n4 = (nwords + 3) >> 2;
                        const uint4 zb = make_uint4((uint32_t)pk_bias, (uint32_t)pk_bias, (uint32_t)pk_bias, (uint32_t)pk_bias);
                        for (int i = lane; i < n4; i += 32) { t4[i] = zb; h4[i] = make_uint4(0u, 0u, 0u, 0u); }
                    } else {
                        const int n4 = (RS * ((nwords + 3) & ~3)) >> 2;
                        for (int i = lane; i < n4; i += 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
                    }
                } else {
                    for (int i = lane; i < RS * nwords; i += 32) wtab[RS * wbase_cur + i] = 0;
                }
                if constexpr (kTwo) { if (ring) { ev_cnt1 = 0; ev_cnt0 = cnt_nxt; } else { ev_cnt0 = 0; ev_cnt1 = cnt_nxt; } }
                __syncwarp();
                if (!two_halves) ring ^= 1;                      // single half: undo the toggle of the loop header
            }

            // ---- per-year indices
            long long ens = 0;
            if (lolh) ens = warp_sum_ll(ens_lane);
            const long long yi = kTwo ? cl * a.ypc + y : cl;
            // state transitions of the year (one REDUX): a year that goes to the redo list is counted by its replay
            {
                const unsigned int ev_year = __reduce_add_sync(0xffffffffu, n_events);
                n_events = 0u;
                if (!year_bad && lane == 0) ws->acc[7] += ev_year;
            }
            if (year_bad) {
                if (lane == 0) seq_redo_push(a, (long long)chain);
                lolh = 0;
            } else if (lane == 0) {
                if (a.lol) a.lol[yi] = lolh;
                if (a.ens) a.ens[yi] = ens;
                if (a.ent) a.ent[yi] = entries;
                if (a.group_lol && lolh) atomicAdd(&a.group_lol[yi / a.group], (unsigned long long)lolh);
                if (lolh) seq_hist_add(a, ens);
            }
            if (lolh && lane == 0) {             // a year without loss of load adds nothing (ENS and entries are 0 as well)
                ws->acc[0] += lolh; ws->acc[1] += (unsigned long long)ens; ws->acc[2] += entries;
                ws->acc[3] += 1ull;
                ws->acc[4] += (unsigned long long)lolh * lolh;
                const unsigned long long e = (unsigned long long)ens;
                const unsigned long long plo = e * e, phi = __umul64hi(e, e);
                const unsigned long long nlo = ws->acc[5] + plo;
                ws->acc[6] += phi + (nlo < ws->acc[5] ? 1ull : 0ull);
                ws->acc[5] = nlo;
            }
        }
        __syncwarp();
    }

    if (lane == 0) {
        const unsigned long long ev = ws->acc[7];
        if (ws->acc[0]) atomicAdd(&a.acc[ACC_LOL], ws->acc[0]);
        if (ws->acc[1]) atomicAdd(&a.acc[ACC_ENS], ws->acc[1]);
        if (ws->acc[2]) atomicAdd(&a.acc[ACC_ENT], ws->acc[2]);
        if (ws->acc[3]) atomicAdd(&a.acc[ACC_YWL], ws->acc[3]);
        if (ws->acc[4]) atomicAdd(&a.acc[ACC_LOL2], ws->acc[4]);
        if (ws->acc[5] | ws->acc[6]) atomic_add_u128(&a.acc[ACC_ENS2_LO], &a.acc[ACC_ENS2_HI], ws->acc[5], ws->acc[6]);
        if (ev) atomicAdd(&a.acc[ACC_EVENTS], ev);
        atomicAdd(&a.acc[ACC_WAVES], (unsigned long long)ws->diag[0]);
        atomicAdd(&a.acc[ACC_JOBS], (unsigned long long)ws->diag[1]);
        atomicAdd(&a.acc[ACC_OPT_JOBS], (unsigned long long)ws->diag[2]);
        atomicAdd(&a.acc[ACC_FLAGGED], (unsigned long long)ws->diag[3]);
        atomicMax(&a.acc[ACC_PEND_MAX], (unsigned long long)ws->diag[4]);
    }
}

static const void *fast_kernel_ptr(bool disc, bool two, bool pack)
{
    if (two) return disc ? (const void *)seq_fast_kernel<true, true, false> : (const void *)seq_fast_kernel<false, true, false>;
    if (pack) return disc ? (const void *)seq_fast_kernel<true, false, true> : (const void *)seq_fast_kernel<false, false, true>;
    return disc ? (const void *)seq_fast_kernel<true, false, false> : (const void *)seq_fast_kernel<false, false, false>;
}

cudaError_t seq_fast_prepare(bool disc, bool two, bool pack, size_t smem, int threads, int *blocks_per_sm)
{
    const void *k = fast_kernel_ptr(disc, two, pack);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem);
}

void seq_fast_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream)
{
    void *args[] = {(void *)&a};
    cudaLaunchKernel(fast_kernel_ptr(a.disc != 0, a.two_halves != 0, a.pack_shift != 0), dim3(grid), dim3(threads), args, smem, stream);
}

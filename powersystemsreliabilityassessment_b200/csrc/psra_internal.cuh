// psra_internal.cuh -- shared declarations of libpsra_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "psra_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpsra_b200 is written for sm_100a (B200) only"
#endif

#define PSRA_VERSION 1001

#define ACC_COUNT_MAX 32
// largest MTTF / MTTR: event times are 64-bit ticks of 2^-24 h and hour indices 32 bits (a duration is < 2^56 ticks)
#define PSRA_MAX_MEAN_HOURS 1.0e8
#define PSRA_MAX_CHUNKS 8
struct psra_handle {
    int device = 0;
    int sm_count = 0;
    int sm_clock_khz = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;              // history scan + read-back, overlapped with the kernels of `stream` (highest priority)
    cudaStream_t stream3 = nullptr;              // every other launch of a chunked run: its blocks fill the SMs the launch before drains
    cudaEvent_t ev_join = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_chunk[PSRA_MAX_CHUNKS] = {nullptr};
    psra_config cfg{};
    char err[512] = {0};

    // system (device)
    int U = 0;
    int64_t total_cap = 0;
    int32_t max_unit_cap = 0;
    double events_per_hour = 0.0;    // sum over units of 2 / (MTTF + MTTR): expected state transitions per hour
    int32_t *d_cap = nullptr;        // [U]
    float *d_mttf = nullptr;         // [U] binary32 means used by the sampler
    float *d_mttr = nullptr;
    uint32_t *d_for_thr = nullptr;   // [U] floor(FOR * 2^32)
    int32_t *d_order = nullptr;      // [U padded to 32] unit indices, most transitions per hour first
    uint4 *d_wide_tab = nullptr;     // [U padded to 32] in that order: {capacity, bits of mttf * 2^24, bits of mttr * 2^24, FOR threshold} (seq_wide.cu)
    std::vector<double> cycle_sorted;   // [U] MTTF + MTTR in hours, in that order (host copy: static block counts of seq_wide.cu)
    double *d_for = nullptr;         // [U] FOR in FP64 (injected-uniform path, PSA.jl:183)
    // load (device)
    int H = 0;
    int Wd = 0;                      // ceil(H/32)
    int32_t max_load = 0;
    int32_t *d_load = nullptr;       // [Wd*32] zero padded
    int32_t *d_lmax = nullptr;       // [Wd] max load of each 32-hour word
    // non-sequential evaluation tables (built lazily at the first psra_nonseq_* call after psra_set_load)
    bool tab_valid = false;
    int32_t *d_load_sorted = nullptr;   // [H] load curve sorted ascending
    int64_t *d_load_suffix = nullptr;   // [H+1] suffix sums of the sorted curve
    uint16_t *d_lol_tab = nullptr;      // [total_cap+1] LOL hours of every integer capacity (nonseq fast path), or null
    int32_t *d_byte_tab = nullptr;      // [4][256] capacity of every byte pattern of the 32-unit state word
    // accumulators / scratch
    unsigned long long *d_acc = nullptr;  // [32]
    // per-run outputs kept on the device (grown on demand)
    uint32_t *d_lol = nullptr; int64_t *d_ens = nullptr; uint32_t *d_ent = nullptr;
    int64_t out_cap = 0;                  // capacity (elements) of the three vectors above
    int64_t kept_n = 0;                   // valid entries of d_ens/d_lol kept for psra_tail
    uint32_t *d_fail = nullptr; int fail_cap = 0;
    long long *d_group = nullptr; int64_t group_cap = 0;
    void *d_scratch = nullptr; size_t scratch_cap = 0;   // inputs of injected paths, tail keys
    void *d_scratch2 = nullptr; size_t scratch2_cap = 0;
    void *d_hist = nullptr; size_t hist_cap = 0;         // convergence history + its scan partials
    // pinned staging of the history read-back: device -> pinned at link speed, then a host memcpy into the caller's
    // (pageable) buffer; a pageable cudaMemcpyAsync is staged by the driver at a fraction of that and blocks the host
    void *h_pin = nullptr; size_t pin_cap = 0;
    struct PinCopy { void *dst; const void *src; size_t bytes; cudaEvent_t ev; };
    std::vector<PinCopy> pin_pending;
    cudaEvent_t ev_pin[PSRA_MAX_CHUNKS + 2] = {nullptr};
    unsigned long long *d_redo = nullptr;                // [1 + PSRA_REDO_CAP] redo list of the sequential sampler kernels
    // per-year ENS histogram kept for psra_tail (tail.cu): [tail_bins + 2] counts + beyond-range {count, sum}
    unsigned long long *d_tail_hist = nullptr; int64_t tail_bins = 0;
    int64_t hist_years = 0, hist_years_with_loss = 0;    // > 0: the histogram of the last psra_seq_mc call is valid
    void *d_tail_work = nullptr; size_t tail_work_cap = 0;
    // multi-GPU handle (psra_config.ngpus > 1, multi.cu): this handle is device 0 of the set, `peers` the others
    std::vector<psra_handle *> peers;
    void *nccl_comms = nullptr;                          // ncclComm_t[1 + peers.size()]
    unsigned long long *d_red = nullptr;                 // [32] small all-reduce buffer (accumulators as 32-bit-safe limbs)
    bool multi_defer = false;    // set on the devices of a multi-GPU call: leave the per-hour counts, the group sums and the
                                 // ENS histogram on the device (the driver all-reduces / scans them)
    long long hist_carry0 = 0, hist_idx0 = 0;            // running-mean scan: LOL hours / groups in front of this device's range
    unsigned long long last_acc[ACC_COUNT_MAX] = {0};   // accumulators of the last MC call (diagnostics)
};

int psra_fail(psra_handle *h, int code, const char *fmt, ...);

#define PSRA_CUDA(h, call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return psra_fail((h), PSRA_E_CUDA, "%s failed: %s (%s:%d)", #call,               \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                   \
    } while (0)

#define PSRA_REQUIRE(h, cond, msg)                                                           \
    do {                                                                                     \
        if (!(cond)) return psra_fail((h), PSRA_E_INVALID, "%s (%s)", msg, #cond);           \
    } while (0)

// grow-only device buffer helpers (host side)
int psra_reserve(psra_handle *h, void **p, size_t *cap, size_t bytes);
int psra_reserve_outputs(psra_handle *h, int64_t n);
// running means cumsum(group sums)[k] / (group * (k+1)), k < nfull, into the host buffer `history` (device staging in d_hist).
// The sequence is cut into blocks of psra_history_block(nfull) groups; psra_history_range scans the blocks
// [b0, b1) on `stream` (all earlier blocks must have been scanned on the same stream) and copies them to the host.
int64_t psra_history_block(const psra_handle *h, int64_t nfull);
int psra_history_prepare(psra_handle *h, int64_t nfull);
int psra_history_range(psra_handle *h, const long long *d_group, int64_t nfull, int group, int64_t b0, int64_t b1,
                       double *history, cudaStream_t stream);
int psra_history_to_host(psra_handle *h, const long long *d_group, int64_t nfull, int group, double *history);
// wait for the staged ranges of psra_history_range (in issue order) and copy them into the caller's buffer
int psra_history_drain(psra_handle *h);
// (re)allocate and zero the ENS histogram for the system currently set (tail.cu)
int psra_tail_hist_prepare(psra_handle *h);
// index of the last non-empty bin of the device histogram + 1 (0 = empty), synchronous (tail.cu)
int psra_tail_hist_used(psra_handle *h, int64_t *used);
// per-hour failure counts: (re)allocate and zero d_fail (seq_mc.cu)
int psra_seq_prepare_fail(psra_handle *h);
// the single-device engines behind psra_seq_mc / psra_nonseq_mc (seq_mc.cu, nonseq_mc.cu)
int psra_run_seq_range(psra_handle *h, long long chain_base, long long nchains, int ypc, int init_mode, uint64_t seed,
                       const psra_seq_outputs *out, psra_seq_summary *summary);
int psra_run_nonseq_range(psra_handle *h, long long i0, long long n, uint64_t seed, const psra_nonseq_outputs *out,
                          psra_nonseq_summary *summary);
// multi-GPU driver (multi.cu)
int psra_multi_create(psra_handle *h);
void psra_multi_destroy(psra_handle *h);
int psra_multi_seq_mc(psra_handle *h, int64_t year0, int64_t nyears, uint64_t seed, int32_t init_mode, int32_t ypc,
                      const psra_seq_outputs *out, psra_seq_summary *summary);
int psra_multi_nonseq_mc(psra_handle *h, int64_t sample0, int64_t n, uint64_t seed, const psra_nonseq_outputs *out,
                         psra_nonseq_summary *summary);

// ------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// Philox4x32-10 (Salmon et al. SC'11), key (k0,k1), counter (c0..c3) -> 4 words.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;   // IMAD.WIDE.U32
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;                    // one LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// The same generator with the ten round keys read from a table `rk` = {k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85}
// (r = 0..9) that the host put into the kernel arguments: the keys become constant-bank operands of the
// LOP3s instead of 20 registers (or 20 uniform adds per block).
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const uint32_t (&rk)[20], uint32_t (&o)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// E = -ln(u) of the draw x, u = (x | 1) / 2^32 -- sampler specification v2 (the CPU checker
// spells out the same sequence): a fixed sequence of integer operations and correctly rounded
// binary32 add / fma, reproducible bit for bit on any IEEE-754 machine.
//   w = x | 1,  lz = clz(w),  u = w / 2^32 = m * 2^-k  with  k = lz + 1 (1..32)  and  m in [1, 2) truncated to 24 bits;
//   t = m - 1.5 (exact),  R = P7(t) ~ -ln(1.5 + t) (Horner, seven fma; minimax on [-0.5, 0.5), |error| < 2.5e-7);
//   E = fma(k, LN2, R).   -1.2e-7 <= E <= 32 ln 2: for the 768 draws closest to 2^32 the polynomial's error makes E <= 0;
//   the duration clamp max(., 1 tick) of dur_ticks absorbs them.
// Here the normalisation is one conversion: the binary32 value of w rounded toward zero has the exponent field
// e = 158 - lz and the 23 truncated mantissa bits of m.  -k as a float without an integer -> float conversion: a funnel
// shift puts e under the exponent of 2^23 (bits 0x4B000000 + e = the float 2^23 + e), and (2^23 + e) - (2^23 + 159) =
// e - 159 = -k is exact; fma(-k, -LN2, R) rounds the same real number as fma(k, LN2, R).
// 14 instructions per draw (v1, the fdlibm-style reduction to [sqrt(1/2), sqrt(2)) with a degree-9 polynomial: 21).
#define PSRA_LN2_F 0x1.62e430p-1f
// `one_bits` = 0x3F800000 in a register the compiler cannot see through (asm volatile("" : "+r"(one_bits)) outside the
// hot loop): with it the mantissa of m is ONE three-input LOP3, (fw & 0x007FFFFF) | one_bits -- with two immediates
// ptxas emits two.  The overload without it is for code that is not register-bound / not hot.
// Round 2 note: I2FP here and the F2I.S64 of ticks_rn run on the XU pipe (16 lanes per SM).  Builds of seq_wide.cu that simply
// leave them out run 12 % / 20 % faster, but so does leaving out any other eight instructions of the dependent chain: a
// sampler that REPLACES them (23-bit draw, integer log2 through the float bit pattern, duration by IMAD.WIDE + funnel shifts;
// scripts/microbench/genloop.cu keeps it) has 13 instructions per block fewer, none on the XU pipe, four more on the ALU pipe --
// and the same speed: the ALU pipe (2 cycles per warp-instruction) bounds the loop, not the XU pipe.
// Computing both on the FP64 pipe instead (w as the double 2^52 + w minus 2^52; P widened to a double plus 2^52 + 2^51) is
// bit-exact -- verified over all 2^32 draws -- but slower on B200 (DADD is a low-rate instruction: seq_fast -4.5 %, seq_wide
// -9.5 %), and an integer-only RN_int64 needs ~10 instructions; both were dropped.
__device__ __forceinline__ float neglog_u32(uint32_t x, uint32_t one_bits)
{
    const uint32_t fw = __float_as_uint(__uint2float_rz(x | 1u));
    uint32_t mb;
    asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(mb) : "r"(fw), "r"(one_bits));
    const float nk = __fadd_rn(__uint_as_float(__funnelshift_r(fw, 0x00258000u, 23)), -8388767.0f);   // e - 159 = -k
    const float t = __fadd_rn(__uint_as_float(mb), -1.5f);
    float p = -0x1.578b02p-7f;
    p = __fmaf_rn(p, t, 0x1.1d506cp-6f);
    p = __fmaf_rn(p, t, -0x1.a7b9fep-6f);
    p = __fmaf_rn(p, t, 0x1.90d388p-5f);
    p = __fmaf_rn(p, t, -0x1.94b470p-4f);
    p = __fmaf_rn(p, t, 0x1.c72898p-3f);
    p = __fmaf_rn(p, t, -0x1.555536p-1f);
    p = __fmaf_rn(p, t, -0x1.9f324cp-2f);
    return __fmaf_rn(nk, -PSRA_LN2_F, p);
}
__device__ __forceinline__ float neglog_u32(uint32_t x) { return neglog_u32(x, 0x3F800000u); }

// duration of one draw in ticks of 2^-24 h: RN_int64(max(mean_ticks * E, 1)), mean_ticks = mean * 2^24
// RN_int64(P) of a tick duration (1 <= P < 2^56: the host bounds the means, psra_set_system)
__device__ __forceinline__ unsigned long long ticks_rn(float P)
{
    return (unsigned long long)__float2ll_rn(P);
}

__device__ __forceinline__ unsigned long long dur_ticks(float mean_ticks, uint32_t x)
{
    return ticks_rn(fmaxf(__fmul_rn(mean_ticks, neglog_u32(x)), 1.0f));
}

// MATLAB next-event discretisation (Montecarlo_seq/seq_mcsampling.m:52-60): time to failure rounded to
// the nearest hour (half up), time to repair rounded up; result again in ticks
__device__ __forceinline__ unsigned long long dur_ticks_disc(float mean_ticks, uint32_t x, bool up_state, int disc)
{
    unsigned long long t = dur_ticks(mean_ticks, x);
    if (disc) t = ((t + (up_state ? (1ull << 23) : ((1ull << 24) - 1ull))) >> 24) << 24;
    return t;
}

// inclusive warp prefix sum; the shuffle's own "source lane in range" predicate guards the add (no lane compares)
__device__ __forceinline__ int warp_incl_scan(int v, int /*lane*/)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n .reg .pred p;\n .reg .b32 o;\n shfl.sync.up.b32 o|p, %0, %1, 0, 0xffffffff;\n @p add.s32 %0, %0, o;\n}\n"
                     : "+r"(v) : "r"(d));
    return v;
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// warp-wide sum of 128-bit values held as (lo, hi) pairs; every lane ends with the total
__device__ __forceinline__ void warp_sum_u128(unsigned long long &lo, unsigned long long &hi)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long olo = __shfl_xor_sync(0xffffffffu, lo, d), ohi = __shfl_xor_sync(0xffffffffu, hi, d);
        lo += olo;
        hi += ohi + (lo < olo ? 1ull : 0ull);
    }
}

// 128-bit accumulate into (lo, hi) pair in global memory; order independent.
__device__ __forceinline__ void atomic_add_u128(unsigned long long *lo, unsigned long long *hi,
                                                unsigned long long vlo, unsigned long long vhi)
{
    const unsigned long long old = atomicAdd(lo, vlo);
    const unsigned long long carry = (old + vlo < old) ? 1ull : 0ull;
    if (vhi + carry) atomicAdd(hi, vhi + carry);
}

#endif  // __CUDACC__

// accumulator slots in d_acc
enum {
    ACC_LOL = 0, ACC_ENS, ACC_ENT, ACC_YWL, ACC_LOL2, ACC_ENS2_LO, ACC_ENS2_HI, ACC_EVENTS,
    ACC_OVERFLOW, ACC_WAVES, ACC_JOBS, ACC_OPT_JOBS, ACC_FLAGGED, ACC_PEND_MAX, ACC_REDO, ACC_COUNT = 32
};

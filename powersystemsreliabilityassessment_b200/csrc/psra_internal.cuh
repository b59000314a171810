// psra_internal.cuh -- shared declarations of libpsra_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "psra_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpsra_b200 is written for sm_100a (B200) only"
#endif

#define PSRA_VERSION 1001

#define ACC_COUNT_MAX 32
#define PSRA_MAX_CHUNKS 8
struct psra_handle {
    int device = 0;
    int sm_count = 0;
    int sm_clock_khz = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;              // history scan + read-back, overlapped with the kernels of `stream`
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
    int32_t *d_order = nullptr;      // [U] unit indices, most transitions per hour first
    uint4 *d_wide_tab = nullptr;     // [U] in that order: {capacity, bits of mttf * 2^24, bits of mttr * 2^24, FOR threshold} (seq_wide.cu)
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

// E = -ln(u) of the draw x, u = (x | 1) / 2^32: integer normalisation + binary32 fma polynomial,
// a fixed sequence of correctly rounded operations (DESIGN.md "Sampler"), reproducible bit for bit
// on any IEEE-754 machine.  0 < E <= 32 ln 2.
// Specification (what the CPU checker spells out): w = x | 1, lz = clz(w), X = w << lz,
// ix = (X >> 8) + 0x3F000000 + 0x004AFB0D (m = X / 2^31 truncated to 24 bits, fdlibm reduction to
// [sqrt(.5), sqrt(2))), k = lz + 1 - ((ix >> 23) - 127), m' = (ix & 0x7FFFFF) + 0x3F3504F3.
// Here the normalisation is one conversion: the binary32 value of w rounded toward zero has the
// exponent field 158 - lz and the same 23 truncated mantissa bits, so iy = bits(RZ(w)) + 0x004AFB0D
// equals ix + ((31 - lz) << 23): same mantissa field, k = 159 - (iy >> 23).  (float)k comes from the
// 2^23 trick (0 <= k <= 32), which keeps the conversion off the ALU pipe.  Checked exhaustively
// (all 2^32 draws) against the specification on the CPU.
__device__ __forceinline__ float neglog_u32(uint32_t x)
{
    // Pipe balance (the integer ALU pipe is what limits the sequential kernels): with e = (fw + 0x004AFB0D) >> 23 the
    // mantissa of m is fw + 0x3F800000 - (e << 23) (one IMAD; 0x004AFB0D + 0x3F3504F3 = 0x3F800000), and
    // 2k = 2 * (159 - e) is built as a float by an IMAD with the factor -2 (ptxas turns a factor -1 into an ALU
    // subtract); the two ln 2 constants below are halved instead -- exact scalings, so every fma rounds as specified.
    const uint32_t fw = __float_as_uint(__uint2float_rz(x | 1u));
    const uint32_t e = (fw + 0x004AFB0Du) >> 23;
    uint32_t mb, kb;
    asm("mad.lo.u32 %0, %1, 0xFF800000, %2;" : "=r"(mb) : "r"(e), "r"(fw + 0x3F800000u));
    asm("mad.lo.u32 %0, %1, 0xFFFFFFFE, %2;" : "=r"(kb) : "r"(e), "r"(0x4B00013Eu));
    const float m = __uint_as_float(mb);
    const float kf2 = __fadd_rn(__uint_as_float(kb), -8388608.0f);     // 2 k, 0 <= k <= 32
    const float t = __fadd_rn(m, -1.0f);
    float p = 0x1.65b9f8p-4f;
    p = __fmaf_rn(p, t, -0x1.27c4d6p-3f);
    p = __fmaf_rn(p, t, 0x1.32c6a8p-3f);
    p = __fmaf_rn(p, t, -0x1.52fdeep-3f);
    p = __fmaf_rn(p, t, 0x1.98a666p-3f);
    p = __fmaf_rn(p, t, -0x1.000688p-2f);
    p = __fmaf_rn(p, t, 0x1.5557acp-2f);
    p = __fmaf_rn(p, t, -0x1.fffff4p-2f);
    const float r = __fmaf_rn(__fmul_rn(t, t), p, t);       // ln m
    // k * 9.0580006145e-06 + (k * 6.9313812256e-01 - r), written with 2k and the halved constants
    return __fmaf_rn(kf2, 0x1.2fefa2p-18f, __fmaf_rn(kf2, 0x1.62e3p-2f, -r));
}

// duration of one draw in ticks of 2^-24 h: RN_int64(max(mean_ticks * E, 1)), mean_ticks = mean * 2^24
__device__ __forceinline__ unsigned long long dur_ticks(float mean_ticks, uint32_t x)
{
    return (unsigned long long)__float2ll_rn(fmaxf(__fmul_rn(mean_ticks, neglog_u32(x)), 1.0f));
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
    ACC_OVERFLOW, ACC_WAVES, ACC_JOBS, ACC_OPT_JOBS, ACC_FLAGGED, ACC_PEND_MAX, ACC_COUNT = 32
};

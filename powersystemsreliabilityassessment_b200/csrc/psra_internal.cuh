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

struct psra_handle {
    int device = 0;
    int sm_count = 0;
    int sm_clock_khz = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    psra_config cfg{};
    char err[512] = {0};

    // system (device)
    int U = 0;
    int64_t total_cap = 0;
    int32_t *d_cap = nullptr;        // [U]
    float *d_mttf = nullptr;         // [U] binary32 means used by the sampler
    float *d_mttr = nullptr;
    uint32_t *d_for_thr = nullptr;   // [U] floor(FOR * 2^32)
    double *d_for = nullptr;         // [U] FOR in FP64 (injected-uniform path, PSA.jl:183)
    // load (device)
    int H = 0;
    int Wd = 0;                      // ceil(H/32)
    int32_t max_load = 0;
    int32_t *d_load = nullptr;       // [Wd*32] zero padded
    int32_t *d_lmax = nullptr;       // [Wd] max load of each 32-hour word
    // non-sequential lookup tables over capacity c = 0..total_cap (built lazily)
    bool tab_valid = false;
    uint32_t *d_tab_lol = nullptr;   // #{h : load[h] > c}
    int64_t *d_tab_ens = nullptr;    // sum_h max(load[h]-c, 0)
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

// ------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// Philox4x32-10 (Salmon et al. SC'11), key (k0,k1), counter (c0..c3) -> 4 words.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&o)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// E = -ln((x + 0.5) / 2^32) in binary32 with a fixed sequence of correctly rounded
// operations (DESIGN.md "Sampler"): reproducible bit for bit on any IEEE-754 machine.
__device__ __forceinline__ float neglog_u32(uint32_t x)
{
    const unsigned long long n = 2ull * x + 1ull;
    const float f = __ull2float_rn(n);
    const uint32_t b = __float_as_uint(f);
    int e = (int)(b >> 23) - 127;
    float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) { m = __fmul_rn(m, 0.5f); e += 1; }
    const float t = __fadd_rn(m, -1.0f);
    const float s = __fdiv_rn(t, __fadd_rn(2.0f, t));
    const float z = __fmul_rn(s, s);
    float p = __fmaf_rn(z, 0.111111112f, 0.142857149f);
    p = __fmaf_rn(z, p, 0.2f);
    p = __fmaf_rn(z, p, 0.333333343f);
    p = __fmaf_rn(z, p, 1.0f);
    const float lnm = __fmul_rn(__fmul_rn(2.0f, s), p);
    const float k = (float)(33 - e);
    const float E = __fmaf_rn(k, 9.0580006145e-06f, __fmaf_rn(k, 6.9313812256e-01f, -lnm));
    return fmaxf(E, 9.31322575e-10f);
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
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
    ACC_OVERFLOW, ACC_COUNT = 32
};

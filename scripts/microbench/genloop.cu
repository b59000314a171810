// genloop.cu -- where does the time of the generation loop of seq_wide.cu go?  The loop body (Philox4x32-10 block -> four
// sampler durations -> running 64-bit event time -> four hours -> four shared-memory atomics) with pieces switched off,
// at 5 warps per SMSP (20 per SM, as the product kernel), in SMSP cycles per warp-level block.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../include -I../../powersystemsreliabilityassessment_b200/csrc -o genloop genloop.cu
#include <cstdio>
#include <cstdint>
#include "psra_internal.cuh"

#include <cmath>
#define ITER 512
#define PSRA_CEFF 0x1.62e42c57b31a2p-1
static uint32_t sampler_word(float mean_h)
{
    const double c = (double)mean_h * PSRA_CEFF;
    int e = 0; frexp(c, &e);
    int r = 31 - e; if (r < 0) r = 0; if (r > 31) r = 31;
    const double s = ldexp(c, r + 1);
    long long q = llrint((s - r) / 32.0);
    return (uint32_t)(q * 32 + r);
}
#define TL_WORDS 8736

// ---- experimental sampler "v3" (not in the product): no conversion instructions -- 23-bit draw, integer log2 by the float
// bit pattern + degree-7 correction polynomial, duration = (W * Ei) >> (W & 31) + 1 with a 32-bit mean word W.  Tried in
// round 2 to get the I2FP / F2I.S64 of the product's sampler off the XU pipe: 13 instructions per Philox block fewer, but four
// more on the ALU pipe, which is the pipe that bounds the loop -- no gain in seq_wide.cu (25.5 against 26.0 M years/s).
#define PSRA_V3_K 0x96400002u
__device__ __forceinline__ uint32_t exp_fix_u32(uint32_t x, uint32_t one_bits, uint32_t magic_bits)
{
    uint32_t fb, mb;
    asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(fb) : "r"(x), "r"(magic_bits));
    const uint32_t L = __float_as_uint(__fadd_rn(__uint_as_float(fb), -8388607.5f));
    asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(mb) : "r"(L), "r"(one_bits));
    const float t = __fadd_rn(__uint_as_float(mb), -1.5f);
    float p = 0x1.efa0b6p+16f;
    p = __fmaf_rn(p, t, -0x1.9b9f0ep+17f);
    p = __fmaf_rn(p, t, 0x1.31a772p+18f);
    p = __fmaf_rn(p, t, -0x1.212292p+19f);
    p = __fmaf_rn(p, t, 0x1.23eecep+20f);
    p = __fmaf_rn(p, t, -0x1.4853d8p+21f);
    p = __fmaf_rn(p, t, -0x1.38f8e6p+18f);
    p = __fmaf_rn(p, t, 0x1.95c01ep+23f);
    return PSRA_V3_K - L - __float_as_uint(p);
}
__device__ __forceinline__ unsigned long long dur_ticks3(uint32_t w, uint32_t ei)
{
    const unsigned long long p = (unsigned long long)w * ei;
    const uint32_t lo = __funnelshift_r((uint32_t)p, (uint32_t)(p >> 32), w);
    const uint32_t hi = __funnelshift_r((uint32_t)(p >> 32), 0u, w);
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d + 1ull;
}

// flags: 1 = Philox, 2 = sampler (log + duration + time), 4 = hour + address + predicate, 8 = the atomic itself,
// 16 = the experimental v3 sampler instead of the product's, 32 = v3 with event times in per-unit quanta (duration and accumulation
// are one IMAD.WIDE with a 64-bit addend, the hour one shift of the high word)
template <int F>
__global__ void __launch_bounds__(128, 5) gen(unsigned long long *out, long long *cyc, const uint32_t *rk_g, uint32_t wa, uint32_t wb)
{
    extern __shared__ int32_t tl[];
    __shared__ uint32_t rk_s[20];
    for (int i = threadIdx.x; i < TL_WORDS + 32; i += blockDim.x) tl[i] = 0;
    uint32_t rk[20];
#pragma unroll
    for (int i = 0; i < 20; i++) rk[i] = rk_g[i];
    (void)rk_s;
    const uint32_t tl_s = (uint32_t)__cvta_generic_to_shared(tl);
    uint32_t dummy_s = tl_s + 4u * (TL_WORDS + (threadIdx.x & 31));
    uint32_t one_bits = 0x3F800000u, magic_bits = 0x4B000000u, H = TL_WORDS, qshift = 5u;
    asm volatile("" : "+r"(one_bits), "+r"(magic_bits), "+r"(dummy_s), "+r"(H), "+r"(qshift));
    unsigned long long tm1 = ~0ull;
    unsigned int ne = 0;
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    const int d_a = 7;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (uint32_t b = 1; b <= ITER; b++) {
        uint32_t x[4];
        if (F & 1) philox4x32_10_rk(u, 0u, u ^ 0x5555u, b, rk, x);
        else { x[0] = u * b; x[1] = x[0] + 0x9E3779B9u; x[2] = x[1] + 0x9E3779B9u; x[3] = x[2] + 0x9E3779B9u; }
        unsigned long long t1, t2, t3, t4;
        if ((F & 2) && !(F & 16)) {
            const float ma = __uint_as_float(wa), mb = __uint_as_float(wb);
            t1 = tm1 + ticks_rn(fmaxf(__fmul_rn(ma, neglog_u32(x[0], one_bits)), 1.0f));
            t2 = t1 + ticks_rn(fmaxf(__fmul_rn(mb, neglog_u32(x[1], one_bits)), 1.0f));
            t3 = t2 + ticks_rn(fmaxf(__fmul_rn(ma, neglog_u32(x[2], one_bits)), 1.0f));
            t4 = t3 + ticks_rn(fmaxf(__fmul_rn(mb, neglog_u32(x[3], one_bits)), 1.0f));
        } else if ((F & 2) && (F & 32)) {
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t1) : "r"(wa), "r"(exp_fix_u32(x[0], one_bits, magic_bits)), "l"(tm1));
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t2) : "r"(wb), "r"(exp_fix_u32(x[1], one_bits, magic_bits)), "l"(t1));
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t3) : "r"(wa), "r"(exp_fix_u32(x[2], one_bits, magic_bits)), "l"(t2));
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t4) : "r"(wb), "r"(exp_fix_u32(x[3], one_bits, magic_bits)), "l"(t3));
        } else if (F & 2) {
            t1 = tm1 + dur_ticks3(wa, exp_fix_u32(x[0], one_bits, magic_bits));
            t2 = t1 + dur_ticks3(wb, exp_fix_u32(x[1], one_bits, magic_bits));
            t3 = t2 + dur_ticks3(wa, exp_fix_u32(x[2], one_bits, magic_bits));
            t4 = t3 + dur_ticks3(wb, exp_fix_u32(x[3], one_bits, magic_bits));
        } else {
            t1 = tm1 + ((unsigned long long)x[0] << 8); t2 = t1 + ((unsigned long long)x[1] << 8);
            t3 = t2 + ((unsigned long long)x[2] << 8); t4 = t3 + ((unsigned long long)x[3] << 8);
        }
        tm1 = t4;
        if (F & 32) { if (tm1 > (7000ull << 37)) tm1 -= (7000ull << 37); }
        else if (tm1 > (7000ull << 24)) tm1 -= (7000ull << 24);      // stay inside the year: the events land all over the timeline
        if (F & 4) {
            uint32_t h[4];
            if (F & 32) {        // quantum 2^-37 h: the hour is the high word shifted by a per-unit amount
                h[0] = (uint32_t)(t1 >> 32) >> qshift; h[1] = (uint32_t)(t2 >> 32) >> qshift;
                h[2] = (uint32_t)(t3 >> 32) >> qshift; h[3] = (uint32_t)(t4 >> 32) >> qshift;
            } else {
                h[0] = __funnelshift_r((uint32_t)t1, (uint32_t)(t1 >> 32), 24); h[1] = __funnelshift_r((uint32_t)t2, (uint32_t)(t2 >> 32), 24);
                h[2] = __funnelshift_r((uint32_t)t3, (uint32_t)(t3 >> 32), 24); h[3] = __funnelshift_r((uint32_t)t4, (uint32_t)(t4 >> 32), 24);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (F & 8) {
                    asm volatile("{\n .reg .pred p;\n .reg .b32 ad;\n setp.lt.u32 p, %1, %2;\n mad.lo.u32 ad, %1, 4, %3;\n selp.b32 ad, ad, %4, p;\n"
                                 " red.shared.add.s32 [ad], %5;\n @p add.u32 %0, %0, 1;\n}\n"
                                 : "+r"(ne) : "r"(h[k]), "r"(H), "r"(tl_s), "r"(dummy_s), "r"((k & 1) ? -d_a : d_a) : "memory");
                } else {
                    asm volatile("{\n .reg .pred p;\n .reg .b32 ad;\n setp.lt.u32 p, %1, %2;\n mad.lo.u32 ad, %1, 4, %3;\n selp.b32 ad, ad, %4, p;\n"
                                 " xor.b32 %5, %5, ad;\n @p add.u32 %0, %0, 1;\n}\n"
                                 : "+r"(ne), "+r"(acc) : "r"(h[k]), "r"(H), "r"(tl_s), "r"(dummy_s));
                }
            }
        } else acc ^= (uint32_t)t1 ^ (uint32_t)t2 ^ (uint32_t)t3 ^ (uint32_t)(t4 >> 7);
    }
    const long long t1c = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = tm1 + ne + acc + tl[threadIdx.x];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1c - t0;
}

template <int F>
void run(const char *name, int bps, int threads)
{
    unsigned long long *out; long long *cyc; uint32_t *rk;
    const int grid = 148 * bps;
    cudaMalloc(&out, sizeof(unsigned long long) * grid * threads);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaMalloc(&rk, 80);
    uint32_t h_rk[20];
    for (int r = 0; r < 10; r++) { h_rk[2 * r] = 123u + r * 0x9E3779B9u; h_rk[2 * r + 1] = 456u + r * 0xBB67AE85u; }
    cudaMemcpy(rk, h_rk, 80, cudaMemcpyHostToDevice);
    const size_t smem = 4 * (TL_WORDS + 32);
    cudaFuncSetAttribute(gen<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint32_t wa = sampler_word(2940.0f), wb = sampler_word(60.0f);
    if (F & 32) { wa = (uint32_t)llrint(2940.0 * PSRA_CEFF * 8192.0); wb = (uint32_t)llrint(60.0 * PSRA_CEFF * 8192.0); }   // mean * C * 2^(37 - 24)
    if (!(F & 16)) { const float fa = 2940.0f * 16777216.0f, fb = 60.0f * 16777216.0f; memcpy(&wa, &fa, 4); memcpy(&wb, &fb, 4); }
    for (int rep = 0; rep < 2; rep++) gen<F><<<grid, threads, smem>>>(out, cyc, rk, wa, wb);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < grid; i++) s += (double)h[i];
    s /= grid;
    const double warps_per_smsp = bps * threads / 32 / 4.0;
    printf("%-52s %d blocks/SM x %d warps: %7.1f SMSP cycles per warp-level Philox block   (%s)\n", name, bps, threads / 32,
           s / (ITER * warps_per_smsp), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); cudaFree(rk);
}

int main()
{
    run<15>("full loop body", 5, 128);
    run<7>("no atomic (address computed, xor-ed)", 5, 128);
    run<3>("Philox + sampler, no hour / scatter", 5, 128);
    run<1>("Philox only", 5, 128);
    run<2>("sampler only", 5, 128);
    run<14>("sampler + scatter, no Philox", 5, 128);
    run<13>("Philox + scatter, no sampler", 5, 128);
    run<15>("full loop body", 2, 128);
    run<15>("full loop body", 1, 128);
    run<31>("full loop body, experimental v3 sampler", 5, 128);
    run<18>("v3 sampler only", 5, 128);
    run<63>("full loop body, v3 sampler, per-unit quanta", 5, 128);
    run<50>("v3 sampler in per-unit quanta only", 5, 128);
    return 0;
}

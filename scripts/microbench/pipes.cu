// pipes.cu -- reciprocal throughput of the instructions the sampler kernels are made of, per SM sub-partition (SMSP),
// on B200: IMAD.WIDE.U32, LOP3, IADD3, SHF, FFMA, the Philox4x32-10 round, at 1..8 warps per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int OP>
__global__ void bench(unsigned long long *out, long long *cyc, uint32_t seed)
{
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3u + 1u, a2 = a0 * 5u + 2u, a3 = a0 * 7u + 3u;
    uint32_t a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
    unsigned long long w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;
    float f0 = a0, f1 = a1, f2 = a2, f3 = a3, f4 = a4, f5 = a5, f6 = a6, f7 = a7;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
        if (OP == 0) {          // 8 independent IMAD.WIDE.U32 (64-bit accumulate)
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w0) : "r"(a0));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w1) : "r"(a1));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w2) : "r"(a2));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w3) : "r"(a3));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w4) : "r"(a4));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w5) : "r"(a5));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w6) : "r"(a6));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w7) : "r"(a7));
        } else if (OP == 1) {   // 8 independent mul.wide.u32 (no accumulate), result feeds the next multiplicand
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a0));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a1));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a2));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a3));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a4));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a5));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a6));
            asm volatile("{.reg .b32 lo, hi; .reg .b64 p; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; xor.b32 %0, lo, hi;}" : "+r"(a7));
        } else if (OP == 2) {   // 8 independent LOP3 (three register inputs)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(a1), "r"(a2));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(a2), "r"(a3));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(a3), "r"(a4));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(a4), "r"(a5));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a4) : "r"(a5), "r"(a6));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a5) : "r"(a6), "r"(a7));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a6) : "r"(a7), "r"(a0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a7) : "r"(a0), "r"(a1));
        } else if (OP == 3) {   // 8 independent FFMA with an immediate addend
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f0) : "f"(f1));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f1) : "f"(f2));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f2) : "f"(f3));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f3) : "f"(f4));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f4) : "f"(f5));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f5) : "f"(f6));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f6) : "f"(f7));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f7) : "f"(f0));
        } else if (OP == 4) {   // 8 independent 32-bit IMAD (mad.lo)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(a1), "r"(a2));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(a2), "r"(a3));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(a3), "r"(a4));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(a4), "r"(a5));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a4) : "r"(a5), "r"(a6));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a5) : "r"(a6), "r"(a7));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a6) : "r"(a7), "r"(a0));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a7) : "r"(a0), "r"(a1));
        } else if (OP == 5) {   // two interleaved Philox4x32 rounds x 4 (8 IMAD.WIDE + 8 LOP3, the real dependency pattern)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const unsigned long long p0 = (unsigned long long)0xD2511F53u * a0, p1 = (unsigned long long)0xCD9E8D57u * a2;
                const unsigned long long q0 = (unsigned long long)0xD2511F53u * a4, q1 = (unsigned long long)0xCD9E8D57u * a6;
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ a1 ^ seed, n2 = (uint32_t)(p0 >> 32) ^ a3 ^ seed;
                const uint32_t m0 = (uint32_t)(q1 >> 32) ^ a5 ^ seed, m2 = (uint32_t)(q0 >> 32) ^ a7 ^ seed;
                a0 = n0; a1 = (uint32_t)p1; a2 = n2; a3 = (uint32_t)p0;
                a4 = m0; a5 = (uint32_t)q1; a6 = m2; a7 = (uint32_t)q0;
            }
        } else if (OP == 6) {   // 8 independent mul.hi.u32
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a0) : "r"(a1));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a1) : "r"(a2));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a2) : "r"(a3));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a3) : "r"(a4));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a4) : "r"(a5));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a5) : "r"(a6));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a6) : "r"(a7));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a7) : "r"(a0));
        } else if (OP == 7) {   // 4 IMAD.WIDE + 4 LOP3 + 8 FFMA interleaved (can the pipes overlap?)
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w0) : "r"(a0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a4) : "r"(a5), "r"(a6));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f0) : "f"(f1));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f1) : "f"(f2));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w1) : "r"(a1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a5) : "r"(a6), "r"(a7));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f2) : "f"(f3));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f3) : "f"(f4));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w2) : "r"(a2));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a6) : "r"(a7), "r"(a4));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f4) : "f"(f5));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f5) : "f"(f6));
            asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w3) : "r"(a3));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a7) : "r"(a4), "r"(a5));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f6) : "f"(f7));
            asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f7) : "f"(f0));
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7 + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 +
                                                 (unsigned long long)(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int inst_per_iter)
{
    unsigned long long *out; long long *cyc;
    cudaMalloc(&out, sizeof(unsigned long long) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    printf("%-44s", name);
    for (int wps = 1; wps <= 8; wps *= 2) {     // warps per SMSP; one block per SM
        bench<OP><<<148, wps * 128>>>(out, cyc, 12345u);
        bench<OP><<<148, wps * 128>>>(out, cyc, 12345u);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; i++) s += (double)h[i];
        s /= 148.0;
        // SMSP cycles per warp-instruction: the block's wps warps of an SMSP issue wps * ITER * inst_per_iter instructions in s cycles
        printf("  %dw: %6.3f", wps, s / ((double)wps * ITER * inst_per_iter));
    }
    printf("   cycles per warp-instruction per SMSP\n");
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("IMAD.WIDE.U32 (64-bit accumulate) x8", 8);
    run<1>("mul.wide.u32 + xor of halves x8", 16);
    run<2>("LOP3 (3 registers) x8", 8);
    run<3>("FFMA (immediate addend) x8", 8);
    run<4>("IMAD (32-bit) x8", 8);
    run<5>("2 x Philox round (8 IMAD.WIDE + 8 LOP3) x4", 64);
    run<6>("IMAD.HI.U32 x8", 8);
    run<7>("4 IMAD.WIDE + 4 LOP3 + 8 FFMA interleaved", 16);
    return 0;
}

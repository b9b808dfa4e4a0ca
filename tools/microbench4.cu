// Pipe model of the integer multiplier on B200, measured with loops whose SASS was checked instruction by instruction
// (round 1's `imad_wide` loop had loop-invariant products: ptxas hoisted them and the loop measured IADD3 pairs, so its
// "59 IMAD.WIDE/clk/SM" was wrong). Every kernel runs ITERS iterations of a fixed instruction mix on 16 independent
// accumulators per thread; operands change every iteration so nothing can be hoisted, and the 64-bit accumulate is
// written as a (mad.lo.cc, madc.hi) pair, which ptxas keeps as one IMAD.WIDE (it re-associates `mad.wide` chains).
// Output: clocks per warp-iteration per SM sub-partition for each mix -> cost per instruction and which mixes overlap.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench4 tools/microbench4.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 16

__device__ __forceinline__ void wide(uint64_t& c, uint32_t a, uint32_t b) {
    uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
    asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
    c = ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void add64(uint64_t& c, uint32_t a, uint32_t b) {
    uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
    asm("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
    c = ((uint64_t)hi << 32) | lo;
}

// MODE bits: 1 = 16 wide MACs, 2 = 16 64-bit adds (IADD3 + IADD3.X), 4 = 32 plain ALU ops (LOP3/IADD3 without carry),
//            8 = 16 IMAD.lo, 16 = 16 DFMA, 32 = 16 wide MACs on one carry chain (IMAD.WIDE.X form), 64 = 16 SHF funnel shifts
template <int MODE>
__global__ void k_mix(uint32_t* out, const uint32_t* in, uint32_t seed) {
    uint64_t w[NACC], s64[NACC];
    uint32_t a[NACC], p[NACC], q[NACC], xl[NACC], xh[NACC];
    double d[NACC];
    const double dm = 1.0000001 + seed * 1e-9, da = 1e-9;
    for (int k = 0; k < NACC; k++) {
        w[k] = threadIdx.x + k; s64[k] = seed + k; a[k] = in[k] + threadIdx.x; p[k] = k + seed + threadIdx.x; q[k] = 3 * k + seed + threadIdx.x;
        xl[k] = in[k + 16]; xh[k] = in[k + 32];
        d[k] = 1.0 + k + threadIdx.x;
    }
    uint32_t b = seed | 1u;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
        if (MODE & 1) {
#pragma unroll
            for (int k = 0; k < NACC; k++) wide(w[k], a[k], b);
        }
        if (MODE & 32) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(xl[0]), "+r"(xh[0]) : "r"(a[0]), "r"(b));
#pragma unroll
            for (int k = 1; k < NACC; k++)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(xl[k]), "+r"(xh[k]) : "r"(a[k]), "r"(b));
        }
        if (MODE & 2) {
#pragma unroll
            for (int k = 0; k < NACC; k++) add64(s64[k], a[k], b);
        }
        if (MODE & 4) {
#pragma unroll
            for (int k = 0; k < NACC; k++) {
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(p[k]) : "r"(a[k]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(q[k]) : "r"(a[k]));
            }
        }
        if (MODE & 8) {
#pragma unroll
            for (int k = 0; k < NACC; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(q[k]) : "r"(a[k]), "r"(b));
        }
        if (MODE & 16) {
#pragma unroll
            for (int k = 0; k < NACC; k++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(dm), "d"(da));
        }
        if (MODE & 64) {
#pragma unroll
            for (int k = 0; k < NACC; k++) asm volatile("shf.r.clamp.b32 %0, %0, %1, 29;" : "+r"(p[k]) : "r"(a[k]));
        }
        b += 2;     // operands change every iteration
    }
    uint32_t s = 0;
    for (int k = 0; k < NACC; k++)
        s += (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32) ^ (uint32_t)s64[k] ^ (uint32_t)(s64[k] >> 32) ^ p[k] ^ q[k] ^ xl[k] ^ xh[k] ^
             (uint32_t)__double2loint(d[k]) ^ (uint32_t)__double2hiint(d[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 2048 * 4);
    uint32_t* in;
    cudaMalloc(&in, 64 * 4);
    { uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x9E3779B9u * (i + 1); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice); }
    const int threads = 256, bps = 4;           // 8 warps per sub-partition
    const int blocks = sms * bps;
    const double warps_per_smsp = threads * bps / 32.0 / 4.0;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"warps_per_smsp\": %.0f,\n", prop.name, sms, khz, warps_per_smsp);
    struct { const char* name; float ms; } r[32];
    int nr = 0;
#define RUN(MODE, NAME) r[nr++] = {NAME, timeit([&] { k_mix<MODE><<<blocks, threads>>>(out, in, 12345u); })};
    RUN(1, "16 IMAD.WIDE (64-bit accumulate, no flags)")
    RUN(32, "16 IMAD.WIDE.X (carry in and out)")
    RUN(8, "16 IMAD.lo")
    RUN(2, "16 x (IADD3 + IADD3.X) 64-bit adds")
    RUN(4, "32 plain ALU ops (LOP3 + IADD3)")
    RUN(64, "16 SHF")
    RUN(16, "16 DFMA")
    RUN(1 | 2, "16 IMAD.WIDE + 16 64-bit adds")
    RUN(1 | 4, "16 IMAD.WIDE + 32 plain ALU ops")
    RUN(1 | 8, "16 IMAD.WIDE + 16 IMAD.lo")
    RUN(1 | 16, "16 IMAD.WIDE + 16 DFMA")
    RUN(32 | 2, "16 IMAD.WIDE.X + 16 64-bit adds")
    RUN(32 | 4, "16 IMAD.WIDE.X + 32 plain ALU ops")
    RUN(1 | 2 | 4, "16 IMAD.WIDE + 16 64-bit adds + 32 plain ALU ops")
    RUN(16 | 2, "16 DFMA + 16 64-bit adds")
    RUN(16 | 4, "16 DFMA + 32 plain ALU ops")
    RUN(1 | 16 | 2, "16 IMAD.WIDE + 16 DFMA + 16 64-bit adds")
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    for (int i = 0; i < nr; i++) {
        // clocks per warp-iteration per sub-partition = time * clock / (iterations * warps per sub-partition)
        double clk = r[i].ms * 1e-3 * khz * 1e3 / ((double)ITERS * warps_per_smsp);
        printf(" \"%s\": {\"ms\": %.4f, \"clk_per_warp_iteration_per_smsp\": %.1f}%s\n", r[i].name, r[i].ms, clk, i + 1 < nr ? "," : "");
    }
    printf("}\n");
    return 0;
}

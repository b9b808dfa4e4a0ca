// Feasibility of a 52-bit-limb field multiplier on the FP64 pipe: the inner step of Emmart-style
// exact products is, per 52x52-bit limb product,
//     hi = fma.rz(a, b, 2^104);  s = (2^104 + 2^52) - hi;  lo = fma.rz(a, b, s);       (3 FP64-pipe instructions)
//     acc_hi += bits(hi);  acc_lo += bits(lo);                                        (2 64-bit integer adds)
// This loop runs exactly that mix on NPROD independent products per iteration and reports clocks per product per SM
// sub-partition, next to the integer multiplier's cost for the same amount of work: a 256-bit Montgomery product is
// 2 x 25 such products (5 limbs) against 136 IMAD.WIDE (4 clk each) = 544 clk.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench6 tools/microbench6.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define NPROD 10

// MODE 0: all five instructions; 1: FP64 part only; 2: integer adds only; 3: all five, the carry half of each add on the
// multiplier pipe (mad with a carry-in: IMAD.X) instead of the ALU
template <int MODE>
__global__ void k_dfma_prod(uint64_t* out, const double* in, double c1, double c2) {
    double a[NPROD], b[NPROD];
    uint64_t acc_hi[NPROD / 2], acc_lo[NPROD / 2];
    for (int k = 0; k < NPROD; k++) { a[k] = in[k] + threadIdx.x; b[k] = in[16 + k] + threadIdx.x; }
    for (int k = 0; k < NPROD / 2; k++) { acc_hi[k] = k; acc_lo[k] = k + 7; }
    double bump = in[40];
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < NPROD; k++) {
            double hi = 0, lo = 0;
            if (MODE != 2) {
                double s;
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(a[k]), "d"(b[k]), "d"(c1));
                asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(s) : "d"(c2), "d"(hi));
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(a[k]), "d"(b[k]), "d"(s));
            } else {
                hi = a[k]; lo = b[k];
            }
            if (MODE != 1) {
                uint64_t h = (uint64_t)__double_as_longlong(hi), l = (uint64_t)__double_as_longlong(lo);
                if (MODE == 3) {
                    uint32_t x0 = (uint32_t)acc_hi[k / 2], x1 = (uint32_t)(acc_hi[k / 2] >> 32), y0 = (uint32_t)acc_lo[k / 2], y1 = (uint32_t)(acc_lo[k / 2] >> 32);
                    asm volatile("add.cc.u32 %0, %0, %2; madc.lo.u32 %1, %3, 1, %1;" : "+r"(x0), "+r"(x1) : "r"((uint32_t)h), "r"((uint32_t)(h >> 32)));
                    asm volatile("add.cc.u32 %0, %0, %2; madc.lo.u32 %1, %3, 1, %1;" : "+r"(y0), "+r"(y1) : "r"((uint32_t)l), "r"((uint32_t)(l >> 32)));
                    acc_hi[k / 2] = ((uint64_t)x1 << 32) | x0;
                    acc_lo[k / 2] = ((uint64_t)y1 << 32) | y0;
                } else {
                    acc_hi[k / 2] += h;
                    acc_lo[k / 2] += l;
                }
            } else {
                a[k] = lo;      // keep the FP chain alive
            }
        }
        if (MODE != 1) {
#pragma unroll
            for (int k = 0; k < NPROD; k++) a[k] += bump;    // operands change every iteration (1 DADD per product: counted below)
        }
    }
    uint64_t s = 0;
    for (int k = 0; k < NPROD / 2; k++) s += acc_hi[k] ^ acc_lo[k];
    for (int k = 0; k < NPROD; k++) s += (uint64_t)__double_as_longlong(a[k]);
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
    int sms = prop.multiProcessorCount, khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint64_t* out; cudaMalloc(&out, (size_t)sms * 2048 * 8);
    double* in; cudaMalloc(&in, 64 * 8);
    { double h[64]; for (int i = 0; i < 64; i++) h[i] = 4503599627370496.0 / 3 + i * 1048577.0; h[40] = 3.0; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice); }
    const double c1 = 20282409603651670423947251286016.0;            // 2^104
    const double c2 = c1 + 4503599627370496.0;                       // 2^104 + 2^52
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", prop.name, sms);
    const char* names[4] = {"dfma+dadd+dfma + 2 x 64-bit add (+1 dadd operand bump)", "fp64 part only (3 per product)", "2 x 64-bit add only (+1 dadd bump)",
                            "all, carry half of the adds as IMAD.X (+1 dadd bump)"};
    for (int wps : {4, 8}) {
        const int threads = 256, bps = wps / 2;
        const int blocks = sms * bps;
        float ms[4];
        ms[0] = timeit([&] { k_dfma_prod<0><<<blocks, threads>>>(out, in, c1, c2); });
        ms[1] = timeit([&] { k_dfma_prod<1><<<blocks, threads>>>(out, in, c1, c2); });
        ms[2] = timeit([&] { k_dfma_prod<2><<<blocks, threads>>>(out, in, c1, c2); });
        ms[3] = timeit([&] { k_dfma_prod<3><<<blocks, threads>>>(out, in, c1, c2); });
        for (int m = 0; m < 4; m++) {
            double clk = ms[m] * 1e-3 * khz * 1e3 / ((double)ITERS * NPROD * wps);
            printf(" \"%s, %d warps/SMSP\": {\"ms\": %.4f, \"clk_per_product_per_smsp\": %.2f},\n", names[m], wps, ms[m], clk);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf(" \"status\": \"%s\", \"reference\": \"8x32 CIOS: 136 IMAD.WIDE x 4 clk = 544 clk per modmul; 5x52: 50 products + 5 quotient products\"}\n", cudaGetErrorString(e));
    return 0;
}

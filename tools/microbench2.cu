// Which instruction mixes overlap on sm_100a? (follow-up to the sparse-modulus experiment)
// Each kernel runs ITERS x a fixed body per thread on 16 warps per SMSP; "ops" = instructions of the named kinds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define ITERS 2048

#define WMADC(lo, hi, a, b) asm volatile("madc.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b))
#define ADDC(x, y) asm volatile("addc.cc.u32 %0,%0,%1;" : "+r"(x) : "r"(y))
#define ADDP(x, y) asm volatile("add.u32 %0,%0,%1;" : "+r"(x) : "r"(y))
#define WMAD(x64, a, b) asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(x64) : "r"(a), "r"(b))

// 16 carry-chained wide mads per iteration
__global__ void k_wide_x(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) { WMADC(x[0], x[1], a, b); WMADC(x[2], x[3], a + 1, b); WMADC(x[4], x[5], a + 2, b); WMADC(x[6], x[7], a + 3, b); }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 32 carry-chained adds per iteration
__global__ void k_addc(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) ADDC(x[k], a + k);
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 16 carry-chained wide mads + 32 carry-chained adds (separate registers) per iteration
__global__ void k_wide_x_addc(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8], y[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x * 3 + k; }
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            WMADC(x[0], x[1], a, b); WMADC(x[2], x[3], a + 1, b); WMADC(x[4], x[5], a + 2, b); WMADC(x[6], x[7], a + 3, b);
#pragma unroll
            for (int k = 0; k < 8; k++) ADDC(y[k], a + k);
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k] + y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 16 carry-chained wide mads + 32 plain adds per iteration
__global__ void k_wide_x_add(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8], y[8];
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x * 3 + k; }
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            WMADC(x[0], x[1], a, b); WMADC(x[2], x[3], a + 1, b); WMADC(x[4], x[5], a + 2, b); WMADC(x[6], x[7], a + 3, b);
#pragma unroll
            for (int k = 0; k < 8; k++) ADDP(y[k], x[k] ^ (a + k));
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k] + y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 16 carry-free wide mads (64-bit accumulate) + 32 carry-chained adds per iteration
__global__ void k_wide_addc(uint32_t* out, uint32_t a, uint32_t b) {
    uint64_t x[4];
    uint32_t y[8];
    for (int k = 0; k < 4; k++) x[k] = threadIdx.x + k;
    for (int k = 0; k < 8; k++) y[k] = threadIdx.x * 3 + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            WMAD(x[0], a, b); WMAD(x[1], a + 1, b); WMAD(x[2], a + 2, b); WMAD(x[3], a + 3, b);
#pragma unroll
            for (int k = 0; k < 8; k++) ADDC(y[k], a + k);
        }
    }
    uint64_t s = 0;
    for (int k = 0; k < 4; k++) s += x[k];
    for (int k = 0; k < 8; k++) s += y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}
// 16 carry-free wide mads only
__global__ void k_wide(uint32_t* out, uint32_t a, uint32_t b) {
    uint64_t x[4];
    for (int k = 0; k < 4; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) { WMAD(x[0], a, b); WMAD(x[1], a + 1, b); WMAD(x[2], a + 2, b); WMAD(x[3], a + 3, b); }
    }
    uint64_t s = 0;
    for (int k = 0; k < 4; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}

template <class F> float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
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
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    int blocks = sms * 8, threads = 256;
    double nt = (double)blocks * threads;
    struct R { const char* name; float ms; double wide, adds; } r[8];
    int n = 0;
    r[n++] = {"wide_x (16 IMAD.WIDE.X)", timeit([&] { k_wide_x<<<blocks, threads>>>(out, 3, 5); }), 16, 0};
    r[n++] = {"addc (32 IADD3.X)", timeit([&] { k_addc<<<blocks, threads>>>(out, 3, 5); }), 0, 32};
    r[n++] = {"wide_x + addc (16 + 32)", timeit([&] { k_wide_x_addc<<<blocks, threads>>>(out, 3, 5); }), 16, 32};
    r[n++] = {"wide_x + plain add (16 + 32)", timeit([&] { k_wide_x_add<<<blocks, threads>>>(out, 3, 5); }), 16, 32};
    r[n++] = {"wide (16 IMAD.WIDE no carry)", timeit([&] { k_wide<<<blocks, threads>>>(out, 3, 5); }), 16, 0};
    r[n++] = {"wide + addc (16 + 32)", timeit([&] { k_wide_addc<<<blocks, threads>>>(out, 3, 5); }), 16, 32};
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", prop.name, sms);
    for (int i = 0; i < n; i++) {
        double clk_per_iter = r[i].ms * 1e-3 * prop.clockRate * 1e3 / ITERS / 4;   // per group of (4 wide [+ 8 adds]) ... per thread
        double per_warp_smsp = r[i].ms * 1e-3 * prop.clockRate * 1e3 / ITERS / 16.0; // 16 warps per SMSP run concurrently: clk per warp-iteration
        printf(" \"%s\": {\"ms\": %.4f, \"clk_per_warp_iteration_per_smsp\": %.1f, \"wide_per_clk_sm\": %.1f, \"adds_per_clk_sm\": %.1f}%s\n", r[i].name, r[i].ms,
               per_warp_smsp, r[i].wide * nt * ITERS / (r[i].ms * 1e-3) / sms / (prop.clockRate * 1e3), r[i].adds * nt * ITERS / (r[i].ms * 1e-3) / sms / (prop.clockRate * 1e3),
               i + 1 < n ? "," : "");
        (void)clk_per_iter;
    }
    printf("}\n");
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}

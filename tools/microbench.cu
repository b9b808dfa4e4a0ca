// Integer-pipe micro-benchmarks for the roofline denominators ("nominal peak
// must be replaced by a measured IMAD micro-benchmark") and the field/curve primitives.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../ark_bulletproofs_b200/csrc/ec.cuh"
using namespace bp;

#define ITERS 4096

__global__ void k_imad_lo(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_hi(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint32_t* out, uint32_t a, uint32_t b) {
    uint64_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(a + k), "r"(b));
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}
// carry-chained wide pairs as used by Fp::mul
__global__ void k_imad_wide_cc(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
        wmad_cc(x[0], x[1], a, b);
        wmadc_cc(x[2], x[3], a + 1, b);
        wmadc_cc(x[4], x[5], a + 2, b);
        wmadc_cc(x[6], x[7], a + 3, b);
        wmad_cc(x[1], x[2], a, b + 1);
        wmadc_cc(x[3], x[4], a + 1, b + 1);
        wmadc_cc(x[5], x[6], a + 2, b + 1);
        wmadc_cc(x[7], x[0], a + 3, b + 1);
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
#define MITERS 256
template <class M>
__global__ void k_fpmul(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32);
    for (int i = 0; i < MITERS; i++) {
        x = Fp<M>::mul(x, y);
        y = Fp<M>::mul(y, x);
    }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, Fp<M>::add(x, y));
}
template <class M>
__global__ void k_fpsqr(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32);
    for (int i = 0; i < MITERS; i++) {
        x = Fp<M>::sqr(x);
        y = Fp<M>::sqr(y);
    }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, Fp<M>::add(x, y));
}
template <class M>
__global__ void k_fpadd(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32);
    for (int i = 0; i < MITERS * 8; i++) {
        x = Fp<M>::add(x, y);
        y = Fp<M>::sub(y, x);
    }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, Fp<M>::add(x, y));
}
__global__ void k_madd(xyzz* out, const affine* in) {
    using E = SW<Secq256k1>;
    affine p = ld_affine(in + threadIdx.x % 32);
    xyzz acc = E::dbl_affine(ld_affine(in + 32 + threadIdx.x % 32));
    for (int i = 0; i < MITERS; i++) E::madd(acc, p);
    st_xyzz(out + blockIdx.x * blockDim.x + threadIdx.x, acc);
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
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, prop.clockRate);
    void* out; cudaMalloc(&out, (size_t)sms * 16 * 1024 * 128);
    fe* in; cudaMalloc(&in, 64 * 64);
    uint32_t hin[64 * 16];
    for (int i = 0; i < 64 * 16; i++) hin[i] = 0x12345 * (i + 1) + 77;
    for (int i = 0; i < 64; i++) hin[i * 8 + 7] &= 0x7FFFFFFF;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    // a valid curve point for madd: use generator in both slots
    affine g;
    for (int k = 0; k < 8; k++) { g.x.v[k] = Secq256k1::gx(k); g.y.v[k] = Secq256k1::gy(k); }
    affine* pin; cudaMalloc(&pin, 64 * sizeof(affine));
    { affine h[64]; for (int i = 0; i < 64; i++) h[i] = g; cudaMemcpy(pin, h, sizeof(h), cudaMemcpyHostToDevice); }
    int blocks = sms * 8, threads = 256;
    double nthreads = (double)blocks * threads;
    struct { const char* name; float ms; double ops; } r[24];
    int nr = 0;
    r[nr++] = {"imad_lo", timeit([&] { k_imad_lo<<<blocks, threads>>>((uint32_t*)out, 3, 5); }), nthreads * ITERS * 8};
    r[nr++] = {"imad_hi", timeit([&] { k_imad_hi<<<blocks, threads>>>((uint32_t*)out, 3, 5); }), nthreads * ITERS * 8};
    r[nr++] = {"imad_wide", timeit([&] { k_imad_wide<<<blocks, threads>>>((uint32_t*)out, 3, 5); }), nthreads * ITERS * 8};
    r[nr++] = {"imad_wide_cc", timeit([&] { k_imad_wide_cc<<<blocks, threads>>>((uint32_t*)out, 3, 5); }), nthreads * ITERS * 8};
    r[nr++] = {"iadd", timeit([&] { k_iadd3<<<blocks, threads>>>((uint32_t*)out, 3, 5); }), nthreads * ITERS * 8};
    for (int t : {128, 256, 512}) {
        int bl = sms * (t == 512 ? 2 : t == 256 ? 4 : 8);
        double nt = (double)bl * t;
        static char names[8][32];
        snprintf(names[nr % 8], 32, "fpmul_secqfq_t%d", t);
        r[nr] = {names[nr % 8], timeit([&] { k_fpmul<SecqFq><<<bl, t>>>((fe*)out, in); }), nt * MITERS * 2};
        nr++;
    }
    r[nr++] = {"fpsqr_secqfq", timeit([&] { k_fpsqr<SecqFq><<<blocks, threads>>>((fe*)out, in); }), nthreads * MITERS * 2};
    r[nr++] = {"fpmul_zorrofq", timeit([&] { k_fpmul<ZorroFq><<<blocks, threads>>>((fe*)out, in); }), nthreads * MITERS * 2};
    r[nr++] = {"fpmul_secqfr", timeit([&] { k_fpmul<SecqFr><<<blocks, threads>>>((fe*)out, in); }), nthreads * MITERS * 2};
    r[nr++] = {"fpaddsub_secqfq", timeit([&] { k_fpadd<SecqFq><<<blocks, threads>>>((fe*)out, in); }), nthreads * MITERS * 16};
    r[nr++] = {"madd_secq", timeit([&] { k_madd<<<sms * 8, 128>>>((xyzz*)out, pin); }), (double)sms * 8 * 128 * MITERS};
    for (int i = 0; i < nr; i++)
        printf(" \"%s\": {\"ms\": %.4f, \"gops\": %.2f, \"per_sm_per_clk_at_%dMHz\": %.2f}%s\n", r[i].name, r[i].ms,
               r[i].ops / r[i].ms / 1e6, prop.clockRate / 1000, r[i].ops / (r[i].ms * 1e-3) / sms / (prop.clockRate * 1e3),
               i + 1 < nr ? "," : "");
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

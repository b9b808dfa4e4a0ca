// Round-2 micro-benchmark: dedicated squaring (Fp::sqr_sos) and the fused two-product operation (Fp::mul2, one Montgomery
// reduction for a*b - c*d) against the plain CIOS product, as dependent chains, plus the XYZZ mixed addition / full
// addition / doubling built on them. Build twice to compare the group law with and without the new field operations:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/microbench7 tools/microbench7.cu
//   nvcc ... -DBP_NO_SQR -DBP_NO_MUL2 -o tools/microbench7_base tools/microbench7.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../ark_bulletproofs_b200/csrc/ec.cuh"
using namespace bp;

#define MITERS 256
using F = Fp<SecqFq>;
using E = SW<Secq256k1>;

__global__ void k_mul(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), x2 = ld_fe(in + (threadIdx.x + 1) % 32), y = ld_fe(in + 32 + threadIdx.x % 32);
    for (int i = 0; i < MITERS; i++) { x = F::mul(x, y); x2 = F::mul(x2, y); }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, F::add(x, x2));
}
__global__ void k_sqr_mul(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32);
    for (int i = 0; i < MITERS; i++) { x = F::mul(x, x); y = F::mul(y, y); }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, F::add(x, y));
}
__global__ void k_sqr_sos(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32);
    for (int i = 0; i < MITERS; i++) { x = F::sqr_sos(x); y = F::sqr_sos(y); }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, F::add(x, y));
}
// a*b - c*d: fused (one reduction) and unfused (two products and a subtraction); each iteration counts as 2 modmul
__global__ void k_mul2_fused(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32), z = ld_fe(in + (threadIdx.x + 5) % 32);
    fe x2 = ld_fe(in + (threadIdx.x + 9) % 32);
    for (int i = 0; i < MITERS; i++) { x = F::mul2<true>(x, y, z, x); x2 = F::mul2<true>(x2, z, y, x2); }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, F::add(x, x2));
}
__global__ void k_mul2_plain(fe* out, const fe* in) {
    fe x = ld_fe(in + threadIdx.x % 32), y = ld_fe(in + 32 + threadIdx.x % 32), z = ld_fe(in + (threadIdx.x + 5) % 32);
    fe x2 = ld_fe(in + (threadIdx.x + 9) % 32);
    for (int i = 0; i < MITERS; i++) { x = F::sub(F::mul(x, y), F::mul(z, x)); x2 = F::sub(F::mul(x2, z), F::mul(y, x2)); }
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, F::add(x, x2));
}
__global__ void k_madd(xyzz* out, const affine* in, int iters) {
    affine p = ld_affine(in + threadIdx.x % 32);
    xyzz acc = E::dbl_affine(ld_affine(in + 32 + threadIdx.x % 32));
    for (int i = 0; i < iters; i++) E::madd(acc, p);
    st_xyzz(out + blockIdx.x * blockDim.x + threadIdx.x, acc);
}
__global__ void k_add(xyzz* out, const affine* in, int iters) {
    xyzz q = E::dbl(E::dbl_affine(ld_affine(in + threadIdx.x % 32)));
    xyzz acc = E::dbl_affine(ld_affine(in + 32 + threadIdx.x % 32));
    for (int i = 0; i < iters; i++) E::add(acc, q);
    st_xyzz(out + blockIdx.x * blockDim.x + threadIdx.x, acc);
}
__global__ void k_dbl(xyzz* out, const affine* in, int iters) {
    xyzz acc = E::dbl_affine(ld_affine(in + 32 + threadIdx.x % 32));
    for (int i = 0; i < iters; i++) acc = E::dbl(acc);
    st_xyzz(out + blockIdx.x * blockDim.x + threadIdx.x, acc);
}
__global__ void k_to_affine(affine* out, const xyzz* in, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine a = E::to_affine(ld_xyzz(in + i));
    st_fe(&out[i].x, a.x);
    st_fe(&out[i].y, a.y);
}

template <class Fn>
static float timeit(Fn f) {
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

static void dump(const char* name, const void* dptr, size_t bytes) {
    static unsigned char h[512 * 128];
    cudaMemcpy(h, dptr, bytes, cudaMemcpyDeviceToHost);
    unsigned long long acc = 1469598103934665603ull;
    for (size_t i = 0; i < bytes; i++) { acc ^= h[i]; acc *= 1099511628211ull; }
    printf(" \"check_%s\": \"%016llx\",\n", name, acc);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
#if defined(BP_NO_SQR)
    const char* variant = "base (sqr = mul, two reductions)";
#else
    const char* variant = "sqr_sos + mul2";
#endif
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"variant\": \"%s\",\n", prop.name, sms, variant);
    const size_t maxthreads = (size_t)sms * 2048;
    void* out;
    cudaMalloc(&out, maxthreads * 128);
    fe* in; cudaMalloc(&in, 64 * 32);
    uint32_t hin[64 * 8];
    for (int i = 0; i < 64 * 8; i++) hin[i] = 0x12345u * (i + 1) + 77u;
    for (int i = 0; i < 64; i++) hin[i * 8 + 7] &= 0x7FFFFFFFu;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    affine g;
    for (int k = 0; k < 8; k++) { g.x.v[k] = Secq256k1::gx(k); g.y.v[k] = Secq256k1::gy(k); }
    affine* pin; cudaMalloc(&pin, 64 * sizeof(affine));
    { affine h[64]; for (int i = 0; i < 64; i++) h[i] = g; cudaMemcpy(pin, h, sizeof(h), cudaMemcpyHostToDevice); }

    // ---- value checks: identical digests in both builds (affine results are representation-free) ----
    {
        k_sqr_mul<<<4, 128>>>((fe*)out, in); dump("sqr_by_mul", out, 512 * 32);
        k_sqr_sos<<<4, 128>>>((fe*)out, in); dump("sqr_sos", out, 512 * 32);
        k_mul2_fused<<<4, 128>>>((fe*)out, in); dump("mul2_fused", out, 512 * 32);
        k_mul2_plain<<<4, 128>>>((fe*)out, in); dump("mul2_plain", out, 512 * 32);
        affine* d1; cudaMalloc(&d1, 512 * sizeof(affine));
        k_madd<<<4, 128>>>((xyzz*)out, pin, 37); k_to_affine<<<4, 128>>>(d1, (xyzz*)out, 512); dump("madd37_affine", d1, 512 * 64);
        k_add<<<4, 128>>>((xyzz*)out, pin, 37); k_to_affine<<<4, 128>>>(d1, (xyzz*)out, 512); dump("add37_affine", d1, 512 * 64);
        k_dbl<<<4, 128>>>((xyzz*)out, pin, 37); k_to_affine<<<4, 128>>>(d1, (xyzz*)out, 512); dump("dbl37_affine", d1, 512 * 64);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    }

    struct Row { char name[64]; float ms; double ops; };
    static Row r[96];
    int nr = 0;
    auto put = [&](const char* base, int t, int bps, float ms, double ops) {
        snprintf(r[nr].name, 64, "%s_t%d_b%d", base, t, bps);
        r[nr].ms = ms; r[nr].ops = ops; nr++;
    };
    for (int bps : {2, 4, 8}) {
        const int t = 128, blocks = sms * bps;
        const double nth = (double)blocks * t;
        put("mul", t, bps, timeit([&] { k_mul<<<blocks, t>>>((fe*)out, in); }), nth * MITERS * 2);
        put("sqr_by_mul", t, bps, timeit([&] { k_sqr_mul<<<blocks, t>>>((fe*)out, in); }), nth * MITERS * 2);
        put("sqr_sos", t, bps, timeit([&] { k_sqr_sos<<<blocks, t>>>((fe*)out, in); }), nth * MITERS * 2);
        put("mul2_fused(2 modmul)", t, bps, timeit([&] { k_mul2_fused<<<blocks, t>>>((fe*)out, in); }), nth * MITERS * 4);
        put("mul2_plain(2 modmul)", t, bps, timeit([&] { k_mul2_plain<<<blocks, t>>>((fe*)out, in); }), nth * MITERS * 4);
    }
    for (int bps : {2, 3, 4, 6}) {
        const int t = 128, blocks = sms * bps;
        const double nth = (double)blocks * t;
        put("madd", t, bps, timeit([&] { k_madd<<<blocks, t>>>((xyzz*)out, pin, 256); }), nth * 256);
        put("add", t, bps, timeit([&] { k_add<<<blocks, t>>>((xyzz*)out, pin, 256); }), nth * 256);
        put("dbl", t, bps, timeit([&] { k_dbl<<<blocks, t>>>((xyzz*)out, pin, 256); }), nth * 256);
    }
    for (int i = 0; i < nr; i++)
        printf(" \"%s\": {\"ms\": %.4f, \"gops\": %.2f}%s\n", r[i].name, r[i].ms, r[i].ops / r[i].ms / 1e6, i + 1 < nr ? "," : "");
    printf("}\n");
    return 0;
}

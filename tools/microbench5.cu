// Variants of the balanced 29-bit multiplication (csrc/fp29.cuh), measured as dependent chains like tools/microbench3.cu:
//   V0  as shipped: carries moved with IMAD.WIDE (column += carry * 1)
//   V1  carries moved with 64-bit adds (IADD3 + IADD3.X / LEA.HI.X on the ALU pipe)
//   V2  3-way Karatsuba on 3-limb blocks (54 instead of 81 products, ~45 64-bit add/subs), carries as in V0
//   V3  V2 with the carries of V1
// All four are checked against Fp29::mul on the device before timing.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/microbench5 tools/microbench5.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../ark_bulletproofs_b200/csrc/ec.cuh"
#include "../ark_bulletproofs_b200/csrc/experimental/fp29.cuh"
using namespace bp;
using M = SecqFq;
using F = Fp29<M>;
static constexpr uint32_t MASK = 0x1FFFFFFFu;
static constexpr int32_t HALF = 1 << 28;

template <bool ALU_CARRY>
__device__ __forceinline__ void carry_into(int64_t& c, int32_t carry) {
    if (ALU_CARRY) c += carry;
    else F::addcarry(c, carry);
}
template <bool ALU_CARRY>
__device__ __forceinline__ fl reduce_v(int64_t (&c)[18]) {
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int32_t q = F::sx29((uint32_t)c[i] * M::NINV29);
#pragma unroll
        for (int j = 0; j < 9; j++)
            if (M::mb(j) != 0) F::mac(c[i + j], q, M::mb(j));
        carry_into<ALU_CARRY>(c[i + 1], (int32_t)(c[i] >> 29));
    }
    fl r;
#pragma unroll
    for (int k = 9; k < 17; k++) {
        r.v[k - 9] = (int32_t)((uint32_t)c[k] & MASK) - HALF;
        carry_into<ALU_CARRY>(c[k + 1], (int32_t)(c[k] >> 29));
    }
    r.v[8] = (int32_t)c[17];
    return r;
}
template <int V>
__device__ __forceinline__ fl mul_v(const fl& a, const fl& b) {
    int64_t c[18];
#pragma unroll
    for (int k = 0; k < 18; k++) c[k] = (k >= 9 && k < 17) ? (int64_t)HALF : 0;
    if (V < 2) {
#pragma unroll
        for (int i = 0; i < 9; i++)
#pragma unroll
            for (int j = 0; j < 9; j++) F::mac(c[i + j], a.v[i], b.v[j]);
    } else {
        // blocks X0 = limbs 0..2, X1 = 3..5, X2 = 6..8
        int64_t t[3][5];
#pragma unroll
        for (int blk = 0; blk < 3; blk++) {
#pragma unroll
            for (int k = 0; k < 5; k++) t[blk][k] = 0;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) F::mac(t[blk][i + j], a.v[3 * blk + i], b.v[3 * blk + j]);
        }
        // cross blocks accumulate straight into the columns: (Xi + Xj)(Yi + Yj) at column 3(i + j)
        int32_t sa[3], sb[3];
#pragma unroll
        for (int pr = 0; pr < 3; pr++) {
            const int i0 = pr == 2 ? 1 : 0, i1 = pr == 0 ? 1 : 2;
#pragma unroll
            for (int i = 0; i < 3; i++) { sa[i] = a.v[3 * i0 + i] + a.v[3 * i1 + i]; sb[i] = b.v[3 * i0 + i] + b.v[3 * i1 + i]; }
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) F::mac(c[3 * (i0 + i1) + i + j], sa[i], sb[j]);
        }
        // Pii: + at 6i, - at 3(i + j) for both partners j != i
#pragma unroll
        for (int k = 0; k < 5; k++) {
            c[k] += t[0][k];      c[3 + k] -= t[0][k];  c[6 + k] -= t[0][k];
            c[6 + k] += t[1][k];  c[3 + k] -= t[1][k];  c[9 + k] -= t[1][k];
            c[12 + k] += t[2][k]; c[6 + k] -= t[2][k];  c[9 + k] -= t[2][k];
        }
    }
    return reduce_v<(V & 1) != 0>(c);
}

#define MITERS 256
template <int V, int CHAINS>
__global__ void k_mulv(fe* out, const fe* in) {
    fl x[CHAINS], y = F::from_storage(ld_fe(in + 32 + threadIdx.x % 32));
    for (int c = 0; c < CHAINS; c++) x[c] = F::from_storage(ld_fe(in + (threadIdx.x + c) % 32));
    for (int i = 0; i < MITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = (V < 0) ? F::mul(x[c], y) : mul_v<(V < 0 ? 0 : V)>(x[c], y);
    }
    fl s = x[0];
    for (int c = 1; c < CHAINS; c++) s = F::add(s, x[c]);
    st_fe(out + blockIdx.x * blockDim.x + threadIdx.x, F::to_storage(s));
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

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", prop.name, sms);
    fe *out, *ref, *in;
    cudaMalloc(&out, (size_t)sms * 2048 * 32);
    cudaMalloc(&ref, 512 * 32);
    cudaMalloc(&in, 64 * 32);
    uint32_t hin[64 * 8];
    for (int i = 0; i < 64 * 8; i++) hin[i] = 0x12345u * (i + 1) + 77u;
    for (int i = 0; i < 64; i++) hin[i * 8 + 7] &= 0x7FFFFFFFu;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    static uint32_t h0[512 * 8], h1[512 * 8];
    k_mulv<-1, 1><<<4, 128>>>(ref, in);
    cudaMemcpy(h0, ref, sizeof(h0), cudaMemcpyDeviceToHost);
    int bad = 0;
    auto check = [&](const char* name) {
        cudaMemcpy(h1, out, sizeof(h1), cudaMemcpyDeviceToHost);
        int ok = memcmp(h0, h1, sizeof(h0)) == 0;
        printf(" \"%s_equals_fp29_mul\": %s,\n", name, ok ? "true" : "false");
        bad += !ok;
    };
    k_mulv<0, 1><<<4, 128>>>(out, in); check("v0");
    k_mulv<1, 1><<<4, 128>>>(out, in); check("v1");
    k_mulv<2, 1><<<4, 128>>>(out, in); check("v2");
    k_mulv<3, 1><<<4, 128>>>(out, in); check("v3");
    const int t = 256, blocks = sms * 4;
    const double ops = (double)blocks * t * MITERS * 2;
    struct { const char* name; float ms; } r[8];
    int nr = 0;
    r[nr++] = {"v0_carry_imad", timeit([&] { k_mulv<0, 2><<<blocks, t>>>(out, in); })};
    r[nr++] = {"v1_carry_alu", timeit([&] { k_mulv<1, 2><<<blocks, t>>>(out, in); })};
    r[nr++] = {"v2_karatsuba_carry_imad", timeit([&] { k_mulv<2, 2><<<blocks, t>>>(out, in); })};
    r[nr++] = {"v3_karatsuba_carry_alu", timeit([&] { k_mulv<3, 2><<<blocks, t>>>(out, in); })};
    for (int i = 0; i < nr; i++)
        printf(" \"%s\": {\"ms\": %.4f, \"g_modmul_per_s\": %.2f}%s\n", r[i].name, r[i].ms, ops / r[i].ms / 1e6, i + 1 < nr ? "," : "");
    printf("}\n");
    return bad != 0;
}

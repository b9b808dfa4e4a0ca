// flattened_constraints on the device.
//
// Replaces the serial sparse scatter of src/r1cs/prover.rs:354-397 / src/r1cs/verifier.rs:304-349:
//   for constraint q, term (var, coeff):  w_{kind(var)}[index(var)] +-= z^(q+1) * coeff
// as   contrib[t] = coeff_t * z^(q_t+1)            (one thread per constraint, powers from a 2^k table; +-1 coefficients,
//                                                   the bulk of real circuits, are not stored: 8 bytes per term cross PCIe)
//      sort terms by (kind, index)                  (cub::DeviceRadixSort)
//      sum runs of equal key                        (cub::DeviceReduce::ReduceByKey with the field add)
//      scatter the per-variable sums into wL, wR, wO, wV, wc.
// Results are field elements, so the order of summation is irrelevant: bit-identical to the host loop.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include "vec_kernels.cuh"

namespace bp {

template <class C>
struct FeAddOp {
    __device__ __forceinline__ fe operator()(const fe& a, const fe& b) const { return Fp<typename C::Fr>::add(a, b); }
};

// The sort keys (kind << 29 | index, index < 2^29, kind in 0..4) are built on the host as the constraints are recorded
// (ConstraintStore, r1cs.cuh) and uploaded as they are. cref: 0 = coefficient +1, 1 = -1, k + 2 = coeff_ex[k].
template <class C>
__global__ void __launch_bounds__(256) flatten_contrib_kernel(const uint32_t* __restrict__ cref, const fe* __restrict__ coeff_ex,
                                                              const uint32_t* __restrict__ start, size_t ncons,
                                                              const __grid_constant__ PowTable zt, uint32_t* __restrict__ perm,
                                                              fe* __restrict__ contrib) {
    using F = Fp<typename C::Fr>;
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ncons) return;
    const fe zp = pow_from_table<F>(zt, (uint32_t)q + 1u);     // exp_z for constraint q is z^(q+1)
    const fe zn = F::neg(zp);
    for (uint32_t t = start[q]; t < start[q + 1]; t++) {
        const uint32_t c = cref[t];
        perm[t] = t;
        st_fe(contrib + t, c == 0 ? zp : c == 1 ? zn : F::mul(zp, ld_fe(coeff_ex + (c - 2u))));
    }
}

static __global__ void __launch_bounds__(256) flatten_gather_kernel(const fe* __restrict__ contrib, const uint32_t* __restrict__ perm, size_t nterms,
                                                             fe* __restrict__ sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nterms) st_fe(sorted + i, ld_fe_rw(contrib + perm[i]));
}

// kinds: 0 committed (wV -= ), 1 left, 2 right, 3 out (+=), 4 one (wc -=)
template <class C>
__global__ void __launch_bounds__(256) flatten_scatter_kernel(const uint32_t* __restrict__ ukeys, const fe* __restrict__ sums,
                                                              const int* __restrict__ nruns, fe* __restrict__ wL, fe* __restrict__ wR,
                                                              fe* __restrict__ wO, fe* __restrict__ wV, fe* __restrict__ wc, uint32_t n, uint32_t m) {
    using F = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)*nruns) return;
    uint32_t k = ukeys[i], kd = k >> 29, ix = k & 0x1FFFFFFFu;
    fe s = ld_fe_rw(sums + i);
    // defence in depth: the host rejects out-of-range variables when they are recorded (check_terms, r1cs.cuh)
    if (kd >= 1 && kd <= 3 ? ix >= n : kd == 0 ? ix >= m : false) return;
    switch (kd) {
        case 0: st_fe(wV + ix, F::neg(s)); break;
        case 1: st_fe(wL + ix, s); break;
        case 2: st_fe(wR + ix, s); break;
        case 3: st_fe(wO + ix, s); break;
        case 4: st_fe(wc, F::neg(s)); break;
    }
}

}  // namespace bp

// Host-side serial tails (compiled by g++, no device code). See host/fp_host.hpp.
#include <cstring>
#include "../../include/bp_b200.h"
#include "ec.cuh"
#include "host/fp_host.hpp"

namespace bp {

template <class C>
static void combine_t(const xyzz* win, int W, int c, uint8_t out_xy[64], int* out_is_identity) {
    using E = GroupLaw<C, HostFp<typename C::Fq>>;
    xyzz acc = win[W - 1];
    for (int w = W - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = E::dbl(acc);
        E::add(acc, win[w]);
    }
    affine a = E::to_affine(acc);
    memcpy(out_xy, &a, 64);
    if (out_is_identity) *out_is_identity = E::is_identity(acc) ? 1 : 0;
}

// Horner over the per-window sums of the Pippenger MSM: sum_w 2^(c*w) * win[w]  -> affine
int host_combine(int curve, const void* win, int W, int c, uint8_t out_xy[64], int* out_is_identity) {
    switch (curve) {
        case BP_CURVE_SECQ256K1: combine_t<Secq256k1>((const xyzz*)win, W, c, out_xy, out_is_identity); return BP_OK;
        case BP_CURVE_ZORRO: combine_t<Zorro>((const xyzz*)win, W, c, out_xy, out_is_identity); return BP_OK;
        case BP_CURVE_CURVE25519: combine_t<Curve25519>((const xyzz*)win, W, c, out_xy, out_is_identity); return BP_OK;
    }
    return BP_ERR_UNSUPPORTED;
}

template <class C>
static void sum_t(const affine* pts, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    using E = GroupLaw<C, HostFp<typename C::Fq>>;
    xyzz acc = E::identity();
    for (size_t i = 0; i < n; i++) E::madd(acc, pts[i]);
    affine a = E::to_affine(acc);
    memcpy(out_xy, &a, 64);
    if (out_is_identity) *out_is_identity = E::is_identity(acc) ? 1 : 0;
}

// sum of a handful of affine points (per-GPU partial MSM results after the all-gather)
int host_points_sum(int curve, const uint8_t* pts_xy, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    switch (curve) {
        case BP_CURVE_SECQ256K1: sum_t<Secq256k1>((const affine*)pts_xy, n, out_xy, out_is_identity); return BP_OK;
        case BP_CURVE_ZORRO: sum_t<Zorro>((const affine*)pts_xy, n, out_xy, out_is_identity); return BP_OK;
        case BP_CURVE_CURVE25519: sum_t<Curve25519>((const affine*)pts_xy, n, out_xy, out_is_identity); return BP_OK;
    }
    return BP_ERR_UNSUPPORTED;
}

}  // namespace bp

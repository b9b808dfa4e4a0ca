// Curve dispatch for the MSM entry points (kernels: msm_kernels.cuh).
#include "ctx.cuh"

namespace bp {
template <class C> int msm_run(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);
template <class C> int synth_points_run(bp_ctx*, void*, size_t, uint64_t);
extern template int msm_run<Secq256k1>(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);
extern template int msm_run<Zorro>(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);
extern template int msm_run<Curve25519>(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);
extern template int synth_points_run<Curve25519>(bp_ctx*, void*, size_t, uint64_t);
extern template int synth_points_run<Secq256k1>(bp_ctx*, void*, size_t, uint64_t);
extern template int synth_points_run<Zorro>(bp_ctx*, void*, size_t, uint64_t);

int msm_dispatch(bp_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    switch (ctx->curve) {
        case BP_CURVE_SECQ256K1:
            return msm_run<Secq256k1>(ctx, (const affine*)d_bases, (const fe*)d_scalars, n, out_xy, out_is_identity);
        case BP_CURVE_ZORRO:
            return msm_run<Zorro>(ctx, (const affine*)d_bases, (const fe*)d_scalars, n, out_xy, out_is_identity);
        case BP_CURVE_CURVE25519:
            return msm_run<Curve25519>(ctx, (const affine*)d_bases, (const fe*)d_scalars, n, out_xy, out_is_identity);
        default:
            ctx->err = "curve not supported by the CUDA MSM yet";
            return BP_ERR_UNSUPPORTED;
    }
}

int synth_points_dispatch(bp_ctx* ctx, void* d_out, size_t n, uint64_t start) {
    switch (ctx->curve) {
        case BP_CURVE_SECQ256K1: return synth_points_run<Secq256k1>(ctx, d_out, n, start);
        case BP_CURVE_ZORRO: return synth_points_run<Zorro>(ctx, d_out, n, start);
        case BP_CURVE_CURVE25519: return synth_points_run<Curve25519>(ctx, d_out, n, start);
        default: ctx->err = "curve not supported"; return BP_ERR_UNSUPPORTED;
    }
}

}  // namespace bp

// Curve dispatch for the MSM entry points (kernels: msm_kernels.cuh).
#include <vector>
#include "ctx.cuh"

namespace bp {
template <class C> int msm_run(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);
template <class C> int synth_points_run(bp_ctx*, void*, size_t, uint64_t);
template <class C> int msm_run_streamed(bp_ctx*, const uint8_t*, const uint8_t*, size_t, const std::vector<size_t>&, const std::vector<size_t>&,
                                        uint8_t*, int*, const affine*);
#define BP_STREAMED_EXTERN(C) \
    extern template int msm_run_streamed<C>(bp_ctx*, const uint8_t*, const uint8_t*, size_t, const std::vector<size_t>&, \
                                            const std::vector<size_t>&, uint8_t*, int*, const affine*);
BP_STREAMED_EXTERN(Secq256k1)
BP_STREAMED_EXTERN(Zorro)
BP_STREAMED_EXTERN(Curve25519)
#undef BP_STREAMED_EXTERN
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

// bp_msm over host buffers, chunks [lo_of[k], lo_of[k] + cnt_of[k]) streamed into one bucket array (msm_run_streamed)
int msm_streamed_dispatch(bp_ctx* ctx, const uint8_t* h_bases, const uint8_t* h_scalars, size_t n, const std::vector<size_t>& lo_of,
                          const std::vector<size_t>& cnt_of, uint8_t out_xy[64], int* out_is_identity, const void* d_bases) {
    const affine* db = (const affine*)d_bases;
    switch (ctx->curve) {
        case BP_CURVE_SECQ256K1: return msm_run_streamed<Secq256k1>(ctx, h_bases, h_scalars, n, lo_of, cnt_of, out_xy, out_is_identity, db);
        case BP_CURVE_ZORRO: return msm_run_streamed<Zorro>(ctx, h_bases, h_scalars, n, lo_of, cnt_of, out_xy, out_is_identity, db);
        case BP_CURVE_CURVE25519: return msm_run_streamed<Curve25519>(ctx, h_bases, h_scalars, n, lo_of, cnt_of, out_xy, out_is_identity, db);
        default: ctx->err = "curve not supported by the CUDA MSM yet"; return BP_ERR_UNSUPPORTED;
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

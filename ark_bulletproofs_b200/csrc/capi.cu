// extern "C" surface of libbp_b200.so (declared in include/bp_b200.h).
#include <cstring>
#include "ctx.cuh"

namespace bp {
int msm_dispatch(bp_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, uint8_t out_xy[64], int* out_is_identity);
int host_points_sum(int curve, const uint8_t* pts_xy, size_t n, uint8_t out_xy[64], int* out_is_identity);
int synth_points_dispatch(bp_ctx* ctx, void* d_out, size_t n, uint64_t start);
}  // namespace bp

extern "C" {

int bp_ctx_create(int curve, int device, bp_ctx** out) {
    if (!out) return BP_ERR_ARG;
    *out = nullptr;
    if (curve < BP_CURVE_SECQ256K1 || curve > BP_CURVE_CURVE25519) return BP_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BP_ERR_NOGPU;   // no CPU fallback, by design
    if (device < 0 || device >= ndev) return BP_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BP_ERR_CUDA;
    bp_ctx* ctx = new bp_ctx();
    ctx->curve = curve;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&ctx->h_result, 16384) != cudaSuccess) {
        delete ctx;
        return BP_ERR_CUDA;
    }
    *out = ctx;
    return BP_OK;
}

void bp_ctx_destroy(bp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bp::DevBuf* bufs[] = {&ctx->keys_a, &ctx->keys_b, &ctx->vals_a, &ctx->vals_b, &ctx->cub_tmp, &ctx->buckets, &ctx->part_keys,
                          &ctx->part_pts, &ctx->seg_out, &ctx->win_out, &ctx->result, &ctx->stage_bases, &ctx->stage_scalars};
    for (auto* b : bufs) b->release();
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    for (int i = 0; i < 8; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* bp_last_error(const bp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* bp_ctx_stream(bp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t bp_ctx_launch_count(const bp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bp_ctx_sync(bp_ctx* ctx) {
    if (!ctx) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

int bp_ctx_set_timing(bp_ctx* ctx, int enable) {
    if (!ctx) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (enable && !ctx->ev[0])
        for (int i = 0; i < 8; i++) BP_CUDA_TRY(ctx, cudaEventCreate(&ctx->ev[i]));
    ctx->timing = enable != 0;
    return BP_OK;
}

int bp_msm_last_phases(const bp_ctx* ctx, float phase_ms[8], int* c, int* windows, uint64_t* entries) {
    if (!ctx || !phase_ms) return BP_ERR_ARG;
    for (int i = 0; i < 8; i++) phase_ms[i] = ctx->phase_ms[i];
    if (c) *c = ctx->last_c;
    if (windows) *windows = ctx->last_W;
    if (entries) *entries = ctx->last_entries;
    return BP_OK;
}

int bp_msm_set_window(bp_ctx* ctx, int c) {
    if (!ctx || c < 0 || c > 20 || c == 1 || c == 2) return BP_ERR_ARG;
    ctx->force_c = c;
    return BP_OK;
}

int bp_msm_device(bp_ctx* ctx, const void* d_bases_xy, const void* d_scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    if (!ctx || !out_xy || (n && (!d_bases_xy || !d_scalars))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return bp::msm_dispatch(ctx, d_bases_xy, d_scalars, n, out_xy, out_is_identity);
}

int bp_msm(bp_ctx* ctx, const uint8_t* bases_xy, const uint8_t* scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    if (!ctx || !out_xy || (n && (!bases_xy || !scalars))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n) {
        BP_CUDA_TRY(ctx, ctx->stage_bases.reserve(n * 64));
        BP_CUDA_TRY(ctx, ctx->stage_scalars.reserve(n * 32));
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage_bases.p, bases_xy, n * 64, cudaMemcpyHostToDevice, ctx->stream));
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage_scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    return bp::msm_dispatch(ctx, ctx->stage_bases.p, ctx->stage_scalars.p, n, out_xy, out_is_identity);
}

int bp_points_sum(bp_ctx* ctx, const uint8_t* points_xy, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    if (!ctx || !out_xy || (n && !points_xy)) return BP_ERR_ARG;
    // a handful of points (one per GPU): serial adds + one inversion, done on the host (host_tail.cpp)
    return bp::host_points_sum(ctx->curve, points_xy, n, out_xy, out_is_identity);
}

int bp_synth_points_device(bp_ctx* ctx, void* d_out_xy, size_t n, uint64_t start) {
    if (!ctx || (n && !d_out_xy)) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return bp::synth_points_dispatch(ctx, d_out_xy, n, start);
}

}  // extern "C"

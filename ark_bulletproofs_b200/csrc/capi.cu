// extern "C" surface of libbp_b200.so (declared in include/bp_b200.h).
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ctx.cuh"
#include "host/nccl_dyn.hpp"
#include <new>
#include <memory>

// No C++ exception unwinds across the C ABI: every multi-statement entry point is a function-try-block.
#define BP_ABI_CATCH catch (const std::bad_alloc&) { return BP_ERR_INTERNAL; } catch (...) { return BP_ERR_INTERNAL; }

namespace bp {
int msm_dispatch(bp_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, uint8_t out_xy[64], int* out_is_identity);
int host_points_sum(int curve, const uint8_t* pts_xy, size_t n, uint8_t out_xy[64], int* out_is_identity);
int synth_points_dispatch(bp_ctx* ctx, void* d_out, size_t n, uint64_t start);
int msm_streamed_dispatch(bp_ctx* ctx, const uint8_t* h_bases, const uint8_t* h_scalars, size_t n, const std::vector<size_t>& lo_of,
                          const std::vector<size_t>& cnt_of, uint8_t out_xy[64], int* out_is_identity, const void* d_bases);
}  // namespace bp

extern "C" {

// BP_DEBUG_BACKTRACE=1: print a native backtrace on SIGFPE / SIGSEGV / SIGABRT (debugging aid for the ctypes harness)
static void bp_debug_signal(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    fprintf(stderr, "libbp_b200: signal %d, native backtrace:\n", sig);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int bp_ctx_create(int curve, int device, bp_ctx** out) try {
    if (!out) return BP_ERR_ARG;
    if (getenv("BP_DEBUG_BACKTRACE")) { signal(SIGFPE, bp_debug_signal); signal(SIGSEGV, bp_debug_signal); signal(SIGABRT, bp_debug_signal); }
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
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->copy_ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->copy_ev[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost(&ctx->h_result, BP_HOST_RESULT_BYTES) != cudaSuccess) {
        delete ctx;
        return BP_ERR_CUDA;
    }
    *out = ctx;
    return BP_OK;
} BP_ABI_CATCH

void bp_ctx_destroy(bp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->for_each_buf([](bp::DevBuf* b) { b->release(); });
    delete ctx->workers;
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->h_coll) cudaFreeHost(ctx->h_coll);
    if (ctx->nccl_comm && bp::nccl_api().ok()) bp::nccl_api().CommDestroy((ncclComm_t)ctx->nccl_comm);
    for (int i = 0; i < 8; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; i++) if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
    delete ctx;
}

const char* bp_last_error(const bp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* bp_ctx_stream(bp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t bp_ctx_launch_count(const bp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bp_ctx_sync(bp_ctx* ctx) try {
    if (!ctx) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BP_OK;
} BP_ABI_CATCH

int bp_ctx_set_timing(bp_ctx* ctx, int enable) try {
    if (!ctx) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (enable && !ctx->ev[0])
        for (int i = 0; i < 8; i++) BP_CUDA_TRY(ctx, cudaEventCreate(&ctx->ev[i]));
    ctx->timing = enable != 0;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_last_phases(const bp_ctx* ctx, float phase_ms[8], int* c, int* windows, uint64_t* entries) try {
    if (!ctx || !phase_ms) return BP_ERR_ARG;
    for (int i = 0; i < 8; i++) phase_ms[i] = ctx->phase_ms[i];
    if (c) *c = ctx->last_c;
    if (windows) *windows = ctx->last_W;
    if (entries) *entries = ctx->last_entries;
    return BP_OK;
} BP_ABI_CATCH

int bp_ctx_last_stage_ms(const bp_ctx* ctx, double out[16]) try {
    if (!ctx || !out) return BP_ERR_ARG;
    for (int i = 0; i < 16; i++) out[i] = ctx->stage_ms[i];
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_set_chunk(bp_ctx* ctx, size_t points) try {
    if (!ctx || points == 0) return BP_ERR_ARG;
    ctx->msm_chunk = points;
    return BP_OK;
} BP_ABI_CATCH

int bp_ipa_set_nofold_threshold(bp_ctx* ctx, size_t n) try {
    if (!ctx) return BP_ERR_ARG;
    ctx->ipa_nofold_n = n;
    return BP_OK;
} BP_ABI_CATCH

int bp_ctx_set_collective(bp_ctx* ctx, int rank, int world, bp_allgather_fn fn, void* user) try {
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world & (world - 1)) || (world > 1 && !fn)) return BP_ERR_ARG;
    if (ctx->nccl_comm && bp::nccl_api().ok()) { bp::nccl_api().CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ctx->rank = rank;
    ctx->world = world;
    ctx->coll = fn;
    ctx->coll_user = user;
    return BP_OK;
} BP_ABI_CATCH

// ---- library-owned collective: NCCL ----
int bp_nccl_unique_id(uint8_t out[128]) try {
    if (!out) return BP_ERR_ARG;
    const bp::NcclApi& n = bp::nccl_api();
    if (!n.ok()) return BP_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) return BP_ERR_CUDA;
    memcpy(out, &id, 128);
    return BP_OK;
} BP_ABI_CATCH

int bp_ctx_init_nccl(bp_ctx* ctx, int rank, int world, const uint8_t unique_id[128]) try {
    if (!ctx || !unique_id || world < 1 || world > 64 || rank < 0 || rank >= world || (world & (world - 1))) return BP_ERR_ARG;
    const bp::NcclApi& n = bp::nccl_api();
    if (!n.ok()) { ctx->err = "libnccl.so.2 not found"; return BP_ERR_UNSUPPORTED; }
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) { n.CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = n.CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) { ctx->err = std::string("ncclCommInitRank: ") + n.GetErrorString(r); return BP_ERR_CUDA; }
    if (!ctx->h_coll) BP_CUDA_TRY(ctx, cudaMallocHost(&ctx->h_coll, BP_HOST_COLL_BYTES));
    BP_CUDA_TRY(ctx, ctx->coll_send.reserve(64 * 8));
    BP_CUDA_TRY(ctx, ctx->coll_recv.reserve(BP_HOST_COLL_BYTES));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    ctx->coll = nullptr;
    ctx->coll_user = nullptr;
    return BP_OK;
} BP_ABI_CATCH

int bp_gens_set_device_generation(bp_ctx* ctx, int enable) try {
    if (!ctx) return BP_ERR_ARG;
    ctx->gens_on_device = enable != 0;
    return BP_OK;
} BP_ABI_CATCH

int bp_pedersen_set_table(bp_ctx* ctx, int enable) try {
    if (!ctx) return BP_ERR_ARG;
    ctx->pedersen_table = enable != 0;
    return BP_OK;
} BP_ABI_CATCH

int bp_ipa_set_glv(bp_ctx* ctx, int enable) try {
    if (!ctx) return BP_ERR_ARG;
    ctx->ipa_glv = enable != 0;
    ctx->ipa_jsf = enable != 2;
    return BP_OK;
} BP_ABI_CATCH

int bp_ipa_set_geometric(bp_ctx* ctx, int enable) try {
    if (!ctx) return BP_ERR_ARG;
    ctx->ipa_geo = enable != 0;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_set_affine_rounds(bp_ctx* ctx, int rounds, size_t min_entries) try {
    if (!ctx || rounds < 0 || rounds > 6) return BP_ERR_ARG;
    ctx->msm_affine_rounds = rounds;
    if (min_entries) ctx->msm_affine_min_entries = min_entries;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_set_sort(bp_ctx* ctx, int mode, size_t min_entries) try {
    if (!ctx || mode < 0 || mode > 1) return BP_ERR_ARG;
    ctx->msm_sort_mode = mode;
    if (min_entries) ctx->msm_sort_min_entries = min_entries;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_set_two_level_reduce(bp_ctx* ctx, int enable) try {
    if (!ctx) return BP_ERR_ARG;
    ctx->msm_pair_reduce = enable != 0;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_set_tiny(bp_ctx* ctx, int max_terms) try {
    if (!ctx || max_terms < 0 || max_terms > 4096) return BP_ERR_ARG;
    ctx->msm_tiny_max = max_terms;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_set_window(bp_ctx* ctx, int c) try {
    if (!ctx || c < 0 || c > 20 || c == 1 || c == 2) return BP_ERR_ARG;
    ctx->force_c = c;
    return BP_OK;
} BP_ABI_CATCH

int bp_msm_device(bp_ctx* ctx, const void* d_bases_xy, const void* d_scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) try {
    if (!ctx || !out_xy || (n && (!d_bases_xy || !d_scalars))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return bp::msm_dispatch(ctx, d_bases_xy, d_scalars, n, out_xy, out_is_identity);
} BP_ABI_CATCH

// bases_xy: host bases (d_bases == nullptr) or ignored when the bases are already on the device
static int msm_host_scalars(bp_ctx* ctx, const uint8_t* bases_xy, const void* d_bases, const uint8_t* scalars, size_t n, uint8_t out_xy[64],
                            int* out_is_identity) {
    // Large host-resident inputs are streamed in chunks: the H2D copy of chunk k+1 (copy stream) overlaps the kernels
    // of chunk k (compute stream), and every chunk adds into the one bucket array of the whole MSM
    // (bp::msm_run_streamed, msm_kernels.cuh). An MSM is a sum of independent terms, so chunking does not change the value.
    const size_t CHUNK = ctx->msm_chunk;
    if (n <= CHUNK + CHUNK / 2) {
        if (n) {
            if (!d_bases) BP_CUDA_TRY(ctx, ctx->stage_bases.reserve(n * 64));
            BP_CUDA_TRY(ctx, ctx->stage_scalars.reserve(n * 32));
            if (!d_bases) BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage_bases.p, bases_xy, n * 64, cudaMemcpyHostToDevice, ctx->stream));
            BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage_scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        }
        return bp::msm_dispatch(ctx, d_bases ? d_bases : ctx->stage_bases.p, ctx->stage_scalars.p, n, out_xy, out_is_identity);
    }
    // Chunk schedule: the copy stream runs back to back (55 GB/s: 1.7 ms per 2^20 points, 0.6 ms when only the scalars
    // move), the kernels follow at ~2.4 ms per 2^20 points. The first copy cannot be hidden, so the first chunk is small
    // (CHUNK/8; CHUNK/4 when only scalars move); every further chunk costs ~0.6 ms (a partly filled last wave of the
    // accumulate kernel, its slot levels, and every run starting from its bucket's value instead of from its first point),
    // so chunks should be few, and chunk k+1 has arrived when chunk k is done as long as it is at most ~1.3x as large (~4x
    // with resident bases) -- so the chunks grow by 1.5x (3x), up to 2 CHUNK (4 CHUNK): waiting ~2 ms for copies in total is
    // cheaper than the four extra chunks a growth of 1.3x needs. Swept at 2^24 with the bucket prefetch of the accumulate
    // kernel (tools/msm_chunk_matrix.py, profiles/r2_msm_chunk_matrix.jsonl): first chunk 2^19 / 2^18 / 2^17 -> 46.4 /
    // 45.1 / 45.4 ms; growth 1.2 / 1.25 / 1.3 / 1.4 / 1.5 -> 48.2 / 47.2 / 47.8 / 45.0 / 45.1 ms; CHUNK 2^20 / 2^22 -> 45.9 /
    // 46.3 ms; against 39.6 ms device-resident. Resident bases: 42.7-43.3 ms for every setting tried.
    // 2^24 points with the default CHUNK = 2^21: 0.26M, 0.39M, 0.59M, 0.88M, 1.3M, 2.0M, 3.0M, 4M, 4.3M.
    // BP_MSM_FIRST_CHUNK / BP_MSM_GROWTH_PCT / BP_MSM_CHUNK_CAP (points, percent, points) override the schedule for sweeps.
    std::vector<size_t> lo_of, cnt_of;
    {
        size_t a = d_bases ? CHUNK / 4 : CHUNK / 8, rem = n, lo = 0;
        if (a == 0) a = 1;
        if (const char* e = getenv("BP_MSM_FIRST_CHUNK")) { size_t v = strtoull(e, nullptr, 10); if (v) a = v; }
        size_t cap = d_bases ? 4 * CHUNK : 2 * CHUNK;
        if (const char* e = getenv("BP_MSM_CHUNK_CAP")) { size_t v = strtoull(e, nullptr, 10); if (v) cap = v; }
        while (rem > 0) {
            size_t take = a < rem ? a : rem;
            if (rem - take < a / 2) take = rem;
            lo_of.push_back(lo); cnt_of.push_back(take);
            lo += take; rem -= take;
            size_t g = d_bases ? 300 : 150;
            if (const char* e = getenv("BP_MSM_GROWTH_PCT")) { size_t v = strtoull(e, nullptr, 10); if (v > 100) g = v; }
            a = a * g / 100 + 1;
            if (a > cap) a = cap;
        }
    }
    return bp::msm_streamed_dispatch(ctx, bases_xy, scalars, n, lo_of, cnt_of, out_xy, out_is_identity, d_bases);
}

int bp_msm(bp_ctx* ctx, const uint8_t* bases_xy, const uint8_t* scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) try {
    if (!ctx || !out_xy || (n && (!bases_xy || !scalars))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return msm_host_scalars(ctx, bases_xy, nullptr, scalars, n, out_xy, out_is_identity);
} BP_ABI_CATCH

// ---- bases resident on the device (every large MSM of the protocol is over fixed generators) ----
struct bp_bases { bp_ctx* ctx; bp::DevBuf buf; size_t n; };
int bp_bases_upload(bp_ctx* ctx, const uint8_t* bases_xy, size_t n, bp_bases** out) try {
    if (!ctx || !out || (n && !bases_xy)) return BP_ERR_ARG;
    *out = nullptr;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<bp_bases> b(new bp_bases{ctx, {}, n});
    if (n) {
        if (b->buf.reserve(n * 64) != cudaSuccess) { ctx->err = "bp_bases_upload: out of device memory"; return BP_ERR_CUDA; }
        cudaError_t e = cudaMemcpyAsync(b->buf.p, bases_xy, n * 64, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { b->buf.release(); ctx->err = cudaGetErrorString(e); return BP_ERR_CUDA; }
    }
    *out = b.release();
    return BP_OK;
} BP_ABI_CATCH
void bp_bases_free(bp_bases* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    b->buf.release();
    delete b;
}
const void* bp_bases_device_ptr(const bp_bases* b) { return b ? b->buf.p : nullptr; }
int bp_msm_bases(bp_ctx* ctx, const bp_bases* bases, size_t offset, const uint8_t* scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) try {
    if (!ctx || !bases || !out_xy || bases->ctx != ctx || (n && !scalars)) return BP_ERR_ARG;
    if (offset > bases->n || n > bases->n - offset) return BP_ERR_LEN;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return msm_host_scalars(ctx, nullptr, n ? (const uint8_t*)bases->buf.p + offset * 64 : nullptr, scalars, n, out_xy, out_is_identity);
} BP_ABI_CATCH

int bp_points_sum(bp_ctx* ctx, const uint8_t* points_xy, size_t n, uint8_t out_xy[64], int* out_is_identity) try {
    if (!ctx || !out_xy || (n && !points_xy)) return BP_ERR_ARG;
    // a handful of points (one per GPU): serial adds + one inversion, done on the host (host_tail.cpp)
    return bp::host_points_sum(ctx->curve, points_xy, n, out_xy, out_is_identity);
} BP_ABI_CATCH

int bp_points_sum_curve(int curve, const uint8_t* points_xy, size_t n, uint8_t out_xy[64], int* out_is_identity) try {
    if (!out_xy || (n && !points_xy)) return BP_ERR_ARG;
    return bp::host_points_sum(curve, points_xy, n, out_xy, out_is_identity);
} BP_ABI_CATCH

int bp_synth_points_device(bp_ctx* ctx, void* d_out_xy, size_t n, uint64_t start) try {
    if (!ctx || (n && !d_out_xy)) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return bp::synth_points_dispatch(ctx, d_out_xy, n, start);
} BP_ABI_CATCH


}  // extern "C"

// =====================================================================================================
// Handle-based API: transcript, RNG, generators, constraint systems, prover, verifier, proofs, IPA.
// =====================================================================================================
#include <memory>
#include <string>
#include "api_types.hpp"

#include "r1cs_types.hpp"

struct bp_transcript { bp::Transcript t; };
struct bp_rng { std::unique_ptr<bp::Rng> r; bp::ChaCha20Rng* chacha = nullptr; };
struct bp_gens { int curve; bp::GensDev* g; };
struct bp_cs { int curve; bp::ConstraintSystemBase* cs; };
struct bp_prover { int curve; bp_ctx* ctx; void* impl; bp_cs cs; };
struct bp_verifier { int curve; bp_ctx* ctx; void* impl; bp_cs cs; };
struct bp_proof { int curve; void* impl; };

namespace {
struct CallbackRng : bp::Rng {
    void* user;
    uint64_t (*f64)(void*);
    uint32_t (*f32)(void*);
    void (*ffill)(void*, uint8_t*, size_t);
    uint32_t next_u32() override { return f32(user); }
    uint64_t next_u64() override { return f64(user); }
    void fill_bytes(uint8_t* out, size_t n) override { ffill(user, out, n); }
};
inline std::string lbl(const uint8_t* l, size_t n) { return std::string(reinterpret_cast<const char*>(l), n); }
inline void split_terms(const bp_term* t, size_t n, std::vector<bp::Variable>& v, std::vector<bp::fe>& c) {
    v.resize(n);
    c.resize(n);
    for (size_t i = 0; i < n; i++) {
        v[i].kind = t[i].var.kind;
        v[i].idx = t[i].var.index;
        memcpy(c[i].v, t[i].coeff, 32);
    }
}
inline void put_var(bp_var* o, const bp::Variable& v) { o->kind = v.kind; o->reserved = 0; o->index = v.idx; }
}  // namespace

extern "C" {

// ---- transcript (merlin::Transcript) ----
bp_transcript* bp_transcript_new(const uint8_t* label, size_t len) {
    bp_transcript* t = new bp_transcript();
    t->t = bp::Transcript(label, len);
    return t;
}
bp_transcript* bp_transcript_clone(const bp_transcript* t) { return t ? new bp_transcript(*t) : nullptr; }
void bp_transcript_free(bp_transcript* t) { delete t; }
void bp_transcript_append_message(bp_transcript* t, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen) {
    t->t.append_message_l(label, llen, msg, mlen);
}
void bp_transcript_append_u64(bp_transcript* t, const uint8_t* label, size_t llen, uint64_t v) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
    t->t.append_message_l(label, llen, b, 8);
}
void bp_transcript_challenge_bytes(bp_transcript* t, const uint8_t* label, size_t llen, uint8_t* out, size_t n) {
    t->t.challenge_bytes_l(label, llen, out, n);
}
int bp_transcript_challenge_scalar(int curve, bp_transcript* t, const uint8_t* label, size_t llen, uint8_t out[32]) try {
    const bp::CurveApi* api = bp::curve_api(curve);
    if (!api || !t) return BP_ERR_ARG;
    return api->challenge_scalar(&t->t, lbl(label, llen).c_str(), out);
} BP_ABI_CATCH

// ---- RNG (rand_core::RngCore) ----
bp_rng* bp_rng_chacha20(const uint8_t seed[32]) {
    bp_rng* r = new bp_rng();
    r->chacha = new bp::ChaCha20Rng(seed);
    r->r.reset(r->chacha);
    return r;
}
bp_rng* bp_rng_from_callbacks(void* user, uint64_t (*next_u64)(void*), uint32_t (*next_u32)(void*), void (*fill_bytes)(void*, uint8_t*, size_t)) {
    CallbackRng* c = new CallbackRng();
    c->user = user; c->f64 = next_u64; c->f32 = next_u32; c->ffill = fill_bytes;
    bp_rng* r = new bp_rng();
    r->r.reset(c);
    return r;
}
void bp_rng_free(bp_rng* r) { delete r; }
uint64_t bp_rng_words_used(const bp_rng* r) { return r && r->chacha ? r->chacha->words_used : 0; }
int bp_rng_scalars(int curve, bp_rng* r, size_t n, uint8_t* out) try {
    const bp::CurveApi* api = bp::curve_api(curve);
    if (!api || !r || (n && !out)) return BP_ERR_ARG;
    return api->rng_scalars(r->r.get(), n, out);
} BP_ABI_CATCH
// merlin's TranscriptRngBuilder as the prover uses it (src/r1cs/prover.rs:483-494)
bp_rng* bp_transcript_build_rng(const bp_transcript* t, const uint8_t* label, size_t llen, const uint8_t* witnesses, size_t nwit, bp_rng* external) {
    if (!t || !external || (nwit && !witnesses)) return nullptr;
    std::vector<std::vector<uint8_t>> wit;
    for (size_t i = 0; i < nwit; i++) wit.emplace_back(witnesses + 32 * i, witnesses + 32 * i + 32);
    bp_rng* r = new bp_rng();
    r->r.reset(new bp::TranscriptRng(t->t.make_rng(lbl(label, llen).c_str(), wit, *external->r)));
    return r;
}
uint64_t bp_rng_next_u64(bp_rng* r) { return r ? r->r->next_u64() : 0; }
int bp_host_keccak_select(int which) { return bp::keccak_select(which); }
int bp_rng_scalar(int curve, bp_rng* r, uint8_t out[32]) try {
    const bp::CurveApi* api = bp::curve_api(curve);
    if (!api || !r) return BP_ERR_ARG;
    return api->rng_scalar(r->r.get(), out);
} BP_ABI_CATCH

// ---- serialisation helpers (ark-serialize) ----
int bp_scalar_to_bytes(int curve, const uint8_t mont[32], uint8_t out[32]) { auto a = bp::curve_api(curve); return a ? a->scalar_to_bytes(mont, out) : BP_ERR_ARG; }
int bp_scalar_from_bytes(int curve, const uint8_t in[32], uint8_t mont[32]) { auto a = bp::curve_api(curve); return a ? a->scalar_from_bytes(in, mont) : BP_ERR_ARG; }
int bp_point_compress(int curve, const uint8_t xy[64], uint8_t out[33]) { auto a = bp::curve_api(curve); return a ? a->point_compress(xy, out) : BP_ERR_ARG; }
int bp_point_serialize_uncompressed(int curve, const uint8_t xy[64], uint8_t out[65]) { auto a = bp::curve_api(curve); return a ? a->point_uncompressed(xy, out) : BP_ERR_ARG; }
int bp_point_decompress(int curve, const uint8_t in[33], uint8_t xy[64]) { auto a = bp::curve_api(curve); return a ? a->point_decompress(in, xy) : BP_ERR_ARG; }

// ---- generators ----
int bp_gens_generate_host(int curve, size_t capacity, uint8_t* G_xy, uint8_t* H_xy, uint8_t B[64], uint8_t B_blinding[64]) try {
    auto a = bp::curve_api(curve);
    return a ? a->gens_generate_host(capacity, G_xy, H_xy, B, B_blinding) : BP_ERR_ARG;
} BP_ABI_CATCH
int bp_gens_create(bp_ctx* ctx, size_t capacity, bp_gens** out) try {
    if (!ctx || !out) return BP_ERR_ARG;
    auto a = bp::curve_api(ctx->curve);
    if (!a) return BP_ERR_UNSUPPORTED;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    bp::GensDev* g = nullptr;
    int rc = a->gens_create(ctx, capacity, &g);
    if (rc) return rc;
    *out = new bp_gens{ctx->curve, g};
    return BP_OK;
} BP_ABI_CATCH
int bp_gens_from_points(bp_ctx* ctx, const uint8_t B[64], const uint8_t B_blinding[64], const uint8_t* G_xy, const uint8_t* H_xy, size_t capacity,
                        bp_gens** out) try {
    if (!ctx || !out || !B || !B_blinding || (capacity && (!G_xy || !H_xy))) return BP_ERR_ARG;
    auto a = bp::curve_api(ctx->curve);
    if (!a) return BP_ERR_UNSUPPORTED;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    bp::GensDev* g = nullptr;
    int rc = a->gens_from_points(ctx, B, B_blinding, G_xy, H_xy, capacity, &g);
    if (rc) return rc;
    *out = new bp_gens{ctx->curve, g};
    return BP_OK;
} BP_ABI_CATCH
void bp_gens_free(bp_gens* g) { if (g) { delete g->g; delete g; } }
size_t bp_gens_capacity(const bp_gens* g) { return g ? g->g->capacity : 0; }
int bp_gens_export(const bp_gens* g, int which, size_t offset, size_t count, uint8_t* out_xy) try {
    if (!g || !out_xy) return BP_ERR_ARG;
    bp_ctx* ctx = g->g->ctx;
    const bp::DevBuf& b = which == 0 ? g->g->G : which == 1 ? g->g->H : g->g->pc;
    size_t lim = 2;
    if (which != 2) { size_t lo; g->g->slice(0, g->g->capacity, lo, lim); }   // sharded contexts export their own shard
    if (offset + count > lim) return BP_ERR_LEN;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(out_xy, b.as<uint8_t>() + offset * 64, count * 64, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BP_OK;
} BP_ABI_CATCH
int bp_pedersen_commit(const bp_gens* g, const uint8_t value[32], const uint8_t blinding[32], uint8_t out_xy[64]) try {
    if (!g) return BP_ERR_ARG;
    return bp::curve_api(g->curve)->pedersen_commit(g->g, value, blinding, out_xy);
} BP_ABI_CATCH

// ---- constraint system (trait ConstraintSystem / RandomizableConstraintSystem) ----
int bp_cs_multiply(bp_cs* cs, const bp_term* left, size_t nl, const bp_term* right, size_t nr, bp_var out[3]) try {
    if (!cs || !out) return BP_ERR_ARG;
    std::vector<bp::Variable> lv, rv;
    std::vector<bp::fe> lc, rc;
    split_terms(left, nl, lv, lc);
    split_terms(right, nr, rv, rc);
    bp::Variable o[3];
    int r = cs->cs->multiply(lv.data(), lc.data(), nl, rv.data(), rc.data(), nr, o);
    for (int i = 0; i < 3; i++) put_var(&out[i], o[i]);
    return r;
} BP_ABI_CATCH
int bp_cs_allocate(bp_cs* cs, const uint8_t* assignment, bp_var* out) try {
    if (!cs || !out) return BP_ERR_ARG;
    bp::fe a;
    if (assignment) memcpy(a.v, assignment, 32);
    bp::Variable o{0, 0};
    int r = cs->cs->allocate(assignment ? &a : nullptr, &o);
    put_var(out, o);
    return r;
} BP_ABI_CATCH
int bp_cs_allocate_multiplier(bp_cs* cs, const uint8_t* left, const uint8_t* right, bp_var out[3]) try {
    if (!cs || !out) return BP_ERR_ARG;
    bp::fe l, r;
    if (left) memcpy(l.v, left, 32);
    if (right) memcpy(r.v, right, 32);
    bp::Variable o[3] = {{0, 0}, {0, 0}, {0, 0}};
    int rc = cs->cs->allocate_multiplier(left ? &l : nullptr, right ? &r : nullptr, o);
    for (int i = 0; i < 3; i++) put_var(&out[i], o[i]);
    return rc;
} BP_ABI_CATCH
int bp_cs_constrain(bp_cs* cs, const bp_term* terms, size_t n) try {
    if (!cs) return BP_ERR_ARG;
    std::vector<bp::Variable> v;
    std::vector<bp::fe> c;
    split_terms(terms, n, v, c);
    return cs->cs->constrain(v.data(), c.data(), n);
} BP_ABI_CATCH
size_t bp_cs_multipliers_len(const bp_cs* cs) { return cs ? cs->cs->multipliers_len() : 0; }
int bp_cs_specify_randomized_constraints(bp_cs* cs, bp_randomized_cb cb, void* user) try {
    if (!cs || !cb) return BP_ERR_ARG;
    int curve = cs->curve;
    return cs->cs->specify_randomized_constraints([cb, user, curve](bp::ConstraintSystemBase& inner) {
        bp_cs h{curve, &inner};
        return cb(&h, user);
    });
} BP_ABI_CATCH
int bp_cs_challenge_scalar(bp_cs* cs, const uint8_t* label, size_t llen, uint8_t out[32]) try {
    if (!cs || !out) return BP_ERR_ARG;
    bp::fe s;
    int rc = cs->cs->challenge_scalar(lbl(label, llen).c_str(), &s);
    if (rc == BP_OK) memcpy(out, s.v, 32);
    return rc;
} BP_ABI_CATCH

// Synthetic measurement circuit of SURVEY.md 8(d) config 2(i): the one-phase public-multiplier chain.
//   (L_i,R_i,O_i) = allocate_multiplier((x_i,k_i)); constrain(R_i - k_i); constrain(L_0 - V) / constrain(L_i - O_{i-1}).
// x0 == NULL builds the verifier's side. ks: n Montgomery scalars.
int bp_cs_chain_circuit(bp_cs* cs, const bp_var* v0, size_t n, const uint8_t* ks, const uint8_t* x0) try {
    if (!cs || !v0 || (n && !ks)) return BP_ERR_ARG;
    bp::Variable v{v0->kind, v0->index};
    return bp::curve_api(cs->curve)->chain_circuit(cs->cs, &v, n, ks, x0);
} BP_ABI_CATCH

// k-shuffle gadget of the reference's benches and tests, built natively (benches/r1cs_secq256k1.rs:35-75).
int bp_cs_shuffle_gadget(bp_cs* cs, const bp_var* x, const bp_var* y, size_t k) try {
    if (!cs || !x || !y || k == 0) return BP_ERR_ARG;
    std::vector<bp::Variable> xs(k), ys(k);
    for (size_t i = 0; i < k; i++) { xs[i] = {x[i].kind, x[i].index}; ys[i] = {y[i].kind, y[i].index}; }
    return bp::curve_api(cs->curve)->shuffle_gadget(cs->cs, xs.data(), ys.data(), k);
} BP_ABI_CATCH

// ---- prover ----
int bp_prover_new(bp_ctx* ctx, const bp_gens* pc_gens, bp_transcript* transcript, bp_prover** out) try {
    if (!ctx || !pc_gens || !transcript || !out) return BP_ERR_ARG;
    auto a = bp::curve_api(ctx->curve);
    if (!a || pc_gens->curve != ctx->curve) return BP_ERR_UNSUPPORTED;
    bp_prover* p = new bp_prover{ctx->curve, ctx, a->prover_new(ctx, pc_gens->g, &transcript->t), {ctx->curve, nullptr}};
    p->cs.cs = a->prover_cs(p->impl);
    *out = p;
    return BP_OK;
} BP_ABI_CATCH
void bp_prover_free(bp_prover* p) { if (p) { bp::curve_api(p->curve)->prover_free(p->impl); delete p; } }
bp_cs* bp_prover_cs(bp_prover* p) { return p ? &p->cs : nullptr; }
int bp_prover_commit(bp_prover* p, const uint8_t value[32], const uint8_t blinding[32], uint8_t out_commitment[64], bp_var* out_var) try {
    if (!p || !value || !blinding || !out_commitment || !out_var) return BP_ERR_ARG;
    bp::Variable v{0, 0};
    int rc = bp::curve_api(p->curve)->prover_commit(p->impl, value, blinding, out_commitment, &v);
    put_var(out_var, v);
    return rc;
} BP_ABI_CATCH
int bp_prover_commit_batch(bp_prover* p, const uint8_t* values, const uint8_t* blindings, size_t m, uint8_t* out_commitments, bp_var* out_vars) try {
    if (!p || (m && (!values || !blindings || !out_commitments || !out_vars))) return BP_ERR_ARG;
    if (((uintptr_t)values | (uintptr_t)blindings | (uintptr_t)out_commitments) & 15) return BP_ERR_ARG;   // 16-byte aligned arrays
    BP_CUDA_TRY(p->ctx, cudaSetDevice(p->ctx->device));
    std::vector<bp::Variable> vars(m);
    int rc = bp::curve_api(p->curve)->prover_commit_batch(p->impl, values, blindings, m, out_commitments, vars.data());
    if (rc) return rc;
    for (size_t i = 0; i < m; i++) put_var(&out_vars[i], vars[i]);
    return BP_OK;
} BP_ABI_CATCH
int bp_prover_prove(bp_prover* p, bp_rng* rng, bp_proof** out) try {
    if (!p || !rng || !out) return BP_ERR_ARG;
    BP_CUDA_TRY(p->ctx, cudaSetDevice(p->ctx->device));
    void* pr = nullptr;
    int rc = bp::curve_api(p->curve)->prover_prove(p->impl, rng->r.get(), &pr);
    if (rc) return rc;
    *out = new bp_proof{p->curve, pr};
    return BP_OK;
} BP_ABI_CATCH

// ---- verifier ----
int bp_verifier_new(bp_ctx* ctx, bp_transcript* transcript, bp_verifier** out) try {
    if (!ctx || !transcript || !out) return BP_ERR_ARG;
    auto a = bp::curve_api(ctx->curve);
    if (!a) return BP_ERR_UNSUPPORTED;
    bp_verifier* v = new bp_verifier{ctx->curve, ctx, a->verifier_new(ctx, &transcript->t), {ctx->curve, nullptr}};
    v->cs.cs = a->verifier_cs(v->impl);
    *out = v;
    return BP_OK;
} BP_ABI_CATCH
void bp_verifier_free(bp_verifier* v) { if (v) { bp::curve_api(v->curve)->verifier_free(v->impl); delete v; } }
bp_cs* bp_verifier_cs(bp_verifier* v) { return v ? &v->cs : nullptr; }
int bp_verifier_commit(bp_verifier* v, const uint8_t commitment[64], bp_var* out_var) try {
    if (!v || !commitment || !out_var) return BP_ERR_ARG;
    bp::Variable var{0, 0};
    int rc = bp::curve_api(v->curve)->verifier_commit(v->impl, commitment, &var);
    put_var(out_var, var);
    return rc;
} BP_ABI_CATCH
int bp_verifier_commit_batch(bp_verifier* v, const uint8_t* commitments, size_t m, bp_var* out_vars) try {
    if (!v || (m && (!commitments || !out_vars))) return BP_ERR_ARG;
    for (size_t i = 0; i < m; i++)
        if (int rc = bp_verifier_commit(v, commitments + 64 * i, out_vars + i)) return rc;
    return BP_OK;
} BP_ABI_CATCH
int bp_verifier_verify(bp_verifier* v, const bp_proof* proof, const bp_gens* gens) try {
    if (!v || !proof || !gens) return BP_ERR_ARG;
    if (proof->curve != v->curve || gens->curve != v->curve || gens->g->ctx != v->ctx) return BP_ERR_ARG;
    BP_CUDA_TRY(v->ctx, cudaSetDevice(v->ctx->device));
    return bp::curve_api(v->curve)->verifier_verify(v->impl, proof->impl, gens->g);
} BP_ABI_CATCH
int bp_batch_verify(bp_ctx* ctx, bp_rng* rng, bp_verifier* const* verifiers, const bp_proof* const* proofs, size_t n, const bp_gens* gens) try {
    if (!ctx || !rng || !gens || (n && (!verifiers || !proofs))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<void*> vs(n);
    std::vector<const void*> ps(n);
    for (size_t i = 0; i < n; i++) {
        if (!verifiers[i] || !proofs[i] || verifiers[i]->curve != ctx->curve || proofs[i]->curve != ctx->curve) return BP_ERR_ARG;
        // the verifiers' scalar vectors live in their context's buffers and are produced on its stream: one context per batch
        if (verifiers[i]->ctx != ctx) return BP_ERR_ARG;
        vs[i] = verifiers[i]->impl;
        ps[i] = proofs[i]->impl;
    }
    if (gens->curve != ctx->curve || gens->g->ctx != ctx) return BP_ERR_ARG;
    return bp::curve_api(ctx->curve)->batch_verify(ctx, rng->r.get(), vs.data(), ps.data(), n, gens->g);
} BP_ABI_CATCH

int bp_batch_verify_partial(bp_ctx* ctx, const uint8_t* alphas, bp_verifier* const* verifiers, const bp_proof* const* proofs, size_t n,
                            const bp_gens* gens, uint8_t out_xy[64], int* out_is_identity) try {
    if (!ctx || !gens || !out_xy || !out_is_identity || (n && (!verifiers || !proofs || !alphas))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<void*> vs(n);
    std::vector<const void*> ps(n);
    for (size_t i = 0; i < n; i++) {
        if (!verifiers[i] || !proofs[i] || verifiers[i]->curve != ctx->curve || proofs[i]->curve != ctx->curve) return BP_ERR_ARG;
        if (verifiers[i]->ctx != ctx) return BP_ERR_ARG;
        vs[i] = verifiers[i]->impl;
        ps[i] = proofs[i]->impl;
    }
    if (gens->curve != ctx->curve || gens->g->ctx != ctx) return BP_ERR_ARG;
    return bp::curve_api(ctx->curve)->batch_verify_partial(ctx, alphas, vs.data(), ps.data(), n, gens->g, out_xy, out_is_identity);
} BP_ABI_CATCH

int bp_batch_verify_set_device_transcript(bp_ctx* ctx, int min_proofs) try {
    if (!ctx || min_proofs < 0) return BP_ERR_ARG;
    ctx->dev_transcript_min = min_proofs;
    return BP_OK;
} BP_ABI_CATCH
int bp_transcript_ipa_challenges_device(bp_ctx* ctx, const bp_transcript* t, uint64_t padded_n, const uint8_t* L_xy, const uint8_t* R_xy, size_t lg_n,
                                        uint8_t* out_u, uint8_t* out_u_inv, uint8_t out_r[32], int* out_identity_seen) try {
    if (!ctx || !t || !out_u || !out_u_inv || !out_r || !out_identity_seen || (lg_n && (!L_xy || !R_xy))) return BP_ERR_ARG;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return bp::curve_api(ctx->curve)->ipa_challenges_device(ctx, &t->t, padded_n, L_xy, R_xy, lg_n, out_u, out_u_inv, out_r, out_identity_seen);
} BP_ABI_CATCH

// ---- proofs ----
void bp_proof_free(bp_proof* p) { if (p) { bp::curve_api(p->curve)->proof_free(p->impl); delete p; } }
int bp_proof_to_bytes(const bp_proof* p, uint8_t* out, size_t cap, size_t* len) try {
    if (!p || !len) return BP_ERR_ARG;
    std::vector<uint8_t> b;
    bp::curve_api(p->curve)->proof_to_bytes(p->impl, b);
    *len = b.size();
    if (!out) return BP_OK;
    if (cap < b.size()) return BP_ERR_LEN;
    memcpy(out, b.data(), b.size());
    return BP_OK;
} BP_ABI_CATCH
int bp_proof_from_bytes(int curve, const uint8_t* data, size_t len, bp_proof** out) try {
    auto a = bp::curve_api(curve);
    if (!a || !data || !out) return BP_ERR_ARG;
    void* pr = nullptr;
    int rc = a->proof_from_bytes(data, len, &pr);
    if (rc) return rc;
    *out = new bp_proof{curve, pr};
    return BP_OK;
} BP_ABI_CATCH
int bp_proofs_from_bytes_batch(bp_ctx* ctx, const uint8_t* const* data, const size_t* lens, size_t n, bp_proof** out, int* status) try {
    if (!ctx || (n && (!data || !lens || !out || !status))) return BP_ERR_ARG;
    auto a = bp::curve_api(ctx->curve);
    if (!a) return BP_ERR_UNSUPPORTED;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<void*> impl(n, nullptr);
    int rc = a->proofs_from_bytes_batch(ctx, data, lens, n, impl.data(), status);
    if (rc) return rc;
    for (size_t i = 0; i < n; i++) out[i] = impl[i] ? new bp_proof{ctx->curve, impl[i]} : nullptr;
    return BP_OK;
} BP_ABI_CATCH
bp_proof* bp_proof_clone(const bp_proof* p) { return p ? new bp_proof{p->curve, bp::curve_api(p->curve)->proof_clone(p->impl)} : nullptr; }
int bp_proof_get_field(const bp_proof* p, int which, uint8_t* buf) { return p && buf ? bp::curve_api(p->curve)->proof_field(p->impl, which, buf, 0) : BP_ERR_ARG; }
int bp_proof_set_field(bp_proof* p, int which, const uint8_t* buf) try {
    return p && buf ? bp::curve_api(p->curve)->proof_field(p->impl, which, const_cast<uint8_t*>(buf), 1) : BP_ERR_ARG;
} BP_ABI_CATCH
size_t bp_proof_rounds(const bp_proof* p) { return p ? bp::curve_api(p->curve)->proof_rounds(p->impl) : 0; }

// ---- InnerProductProof::create ----
int bp_ipa_create(bp_ctx* ctx, bp_transcript* transcript, const uint8_t Q[64], const uint8_t* G_factors, const uint8_t* H_factors,
                  const uint8_t* G_xy, const uint8_t* H_xy, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out_L, uint8_t* out_R,
                  uint8_t out_a[32], uint8_t out_b[32]) try {
    if (!ctx || !transcript || !Q || !G_factors || !H_factors || !G_xy || !H_xy || !a || !b || !out_a || !out_b) return BP_ERR_ARG;
    if (n > 1 && (!out_L || !out_R)) return BP_ERR_ARG;
    auto api = bp::curve_api(ctx->curve);
    if (!api) return BP_ERR_UNSUPPORTED;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return api->ipa_create_host(ctx, &transcript->t, Q, G_factors, H_factors, G_xy, H_xy, a, b, n, out_L, out_R, out_a, out_b);
} BP_ABI_CATCH


int bp_ipa_verify(bp_ctx* ctx, bp_transcript* transcript, size_t n, const uint8_t* L_xy, const uint8_t* R_xy, const uint8_t a[32], const uint8_t b[32],
                  const uint8_t* G_factors, const uint8_t* H_factors, const uint8_t P[64], const uint8_t Q[64], const uint8_t* G_xy,
                  const uint8_t* H_xy) try {
    if (!ctx || !transcript || !a || !b || !G_factors || !H_factors || !P || !Q || !G_xy || !H_xy) return BP_ERR_ARG;
    if (n > 1 && (!L_xy || !R_xy)) return BP_ERR_ARG;
    auto api = bp::curve_api(ctx->curve);
    if (!api) return BP_ERR_UNSUPPORTED;
    BP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return api->ipa_verify_host(ctx, &transcript->t, n, L_xy, R_xy, a, b, G_factors, H_factors, P, Q, G_xy, H_xy);
} BP_ABI_CATCH

}  // extern "C"

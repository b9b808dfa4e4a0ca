// Context: one curve on one GPU, one stream, a grow-only scratch arena.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include "../../include/bp_b200.h"
#include "ec.cuh"
#include "host/workers.hpp"

namespace bp {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace bp

#include <vector>
// host copy of the constraint store whose flattening inputs (and sorted keys / permutation) are resident in the f_* buffers
struct FlattenCache {
    bool valid = false;
    std::vector<uint32_t> key, cref, start, coeff;
};

static constexpr size_t BP_HOST_RESULT_BYTES = 128 * 1024;
static constexpr size_t BP_HOST_COLL_BYTES = 64 * 8 * 64;      // world <= 64 ranks x MSM_MAX_BATCH partial points

struct bp_ctx {
    int curve = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // H2D of the next MSM chunk while the current one computes
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    bp::DevBuf stage2_bases, stage2_scalars;
    std::string err;
    uint64_t launches = 0;
    int force_c = 0;
    size_t msm_chunk = (size_t)1 << 21;
    size_t ipa_nofold_n = (size_t)1 << 13;   // IPA rounds with n <= this use MSMs over the stage generators instead of folding them   // host-buffer MSMs above 1.5x this are chunked (copy/compute overlap)
    // Multi-GPU: one bp_ctx per process/GPU; generators are sharded cyclically by index
    // (rank g holds i = g mod world), every MSM over them yields a partial point per rank, and the partials are
    // exchanged through the host program's collective (NCCL / gloo all-gather) and added on every rank.
    int rank = 0, world = 1;
    bp_allgather_fn coll = nullptr;
    void* coll_user = nullptr;
    uint64_t coll_calls = 0, coll_bytes = 0;
    // library-owned exchange (bp_ctx_init_nccl): ncclAllGather of the 64-byte partial points on `stream`, no callback into
    // the host program. nccl_comm is an ncclComm_t (host/nccl_dyn.hpp keeps NCCL out of the link line).
    void* nccl_comm = nullptr;
    bp::DevBuf coll_send, coll_recv;
    void* h_coll = nullptr;     // pinned, BP_HOST_COLL_BYTES
    bool gens_on_device = true;              // BulletproofGens chains on the GPU where the stream is seekable (bp_gens_set_device_generation)
    bool pedersen_table = true;              // batched Pedersen commitments through the fixed-base table (bp_pedersen_set_table)
    bool ipa_jsf = true;                     // joint-sparse-form digits for the GLV fold (bp_ipa_set_glv(ctx, 2) = binary digits)
    bool ipa_glv = true;                     // GLV split of the uniform fold scalar where the curve has the endomorphism (bp_ipa_set_glv)
    bool ipa_geo = true;                     // use the uniform-scalar fold for geometric factor vectors (bp_ipa_set_geometric)
    int sm_count = 148;
    int msm_affine_rounds = 0;               // batched-affine pair rounds before the XYZZ accumulation (bp_msm_set_affine_rounds)
    size_t msm_affine_min_entries = (size_t)1 << 22;
    // pairs of single large MSMs are ordered by the two-pass bucket sort of msm_sort.cuh instead of cub::DeviceRadixSort
    // (bp_msm_set_sort); below msm_sort_min_entries (point, window) pairs, and for batched MSMs, the library sort stays
    int msm_sort_mode = 1;
    size_t msm_sort_min_entries = (size_t)1 << 22;
    bool msm_pair_reduce = true;             // two-level bucket reduction for large windows (msm_reduce_windows; bp_msm_set_two_level_reduce)
    int msm_tiny_max = 768;                  // MSMs of a batch with at most this many terms each take the single-launch path (0 = never; bp_msm_set_tiny)
    size_t msm_warp_partials_below = (size_t)1 << 17;   // partial-slot lists shorter than this are reduced by warp-segmented scans
    // MSM scratch
    bp::DevBuf keys_a, keys_b, vals_a, vals_b, cub_tmp, buckets, part_keys, part_pts, seg_out, win_out, result;
    bp::DevBuf stage_bases, stage_scalars, pairpts, pairpre, tr_in, tr_pts, tr_out, sort_scratch;
    FlattenCache flatten_cache;
    int dev_transcript_min = 32;             // batches of at least this many proofs derive their IPA challenges on the device (0 = never)
    // IPA / prover / verifier work buffers (r1cs.cuh)
    bp::DevBuf ipa_G, ipa_H, ipa_s, ipa_parts, small, c_v, c_b, c_out;
    bp::DevBuf p_aL, p_aR, p_aO, p_sL, p_sR, p_wL, p_wR, p_wO, p_ypow, p_yinv, p_l, p_r, p_Gf, p_Hf, v_pts, v_sc, v_g, v_h, v_accg, v_acch, f_kind, f_idx, f_coeff, f_start, f_keys, f_keys2, f_perm, f_perm2, f_contrib, f_sorted, f_ukeys, f_sums, f_tmp, f_wv;
    template <class F> void for_each_buf(F f) {
        bp::DevBuf* all[] = {&keys_a, &keys_b, &vals_a, &vals_b, &cub_tmp, &buckets, &part_keys, &part_pts, &seg_out, &win_out, &result,
                             &stage_bases, &stage_scalars, &stage2_bases, &stage2_scalars, &coll_send, &coll_recv, &pairpts, &pairpre, &tr_in, &tr_pts, &tr_out, &sort_scratch, &ipa_G, &ipa_H, &ipa_s, &ipa_parts, &small, &c_v, &c_b, &c_out, &p_aL, &p_aR, &p_aO, &p_sL, &p_sR,
                             &p_wL, &p_wR, &p_wO, &p_ypow, &p_yinv, &p_l, &p_r, &p_Gf, &p_Hf, &v_pts, &v_sc, &v_g, &v_h, &v_accg, &v_acch, &f_kind, &f_idx, &f_coeff, &f_start, &f_keys, &f_keys2, &f_perm, &f_perm2, &f_contrib, &f_sorted,
                             &f_ukeys, &f_sums, &f_tmp, &f_wv};
        for (auto* b : all) f(b);
    }
    bp::HostWorkers* workers = nullptr;      // persistent host threads for parallel Horner chains (created on first use)
    void* h_result = nullptr;   // pinned, BP_HOST_RESULT_BYTES
    // optional per-phase timing of the last MSM (cudaEvents on `stream`)
    bool timing = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // host wall-clock split of the last prove / verify (ms); see bp_ctx_last_stage_ms
    double stage_ms[16] = {0};
    int last_c = 0, last_W = 0;
    size_t last_entries = 0;
};

#define BP_CUDA_TRY(ctx, expr)                                                                       \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                         \
            return BP_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define BP_LAUNCH_CHECK(ctx)                                                                         \
    do {                                                                                             \
        (ctx)->launches++;                                                                           \
        BP_CUDA_TRY(ctx, cudaGetLastError());                                                        \
    } while (0)

// Pippenger variable-base MSM for sm_100a.
//
// Replaces ark-ec `VariableBaseMSM::msm` at the 17 call sites of SURVEY.md 8(a) row a1
// (src/inner_product_proof.rs:104,124,187,202; src/r1cs/prover.rs:516-559,607-648;
// src/r1cs/verifier.rs:574,685). Only the value of the sum is observable, so the algorithm is
// chosen for the GPU:
//   1. digits      : scalar (Montgomery) -> canonical -> signed c-bit digits; emits one
//                    (bucket key, point index | sign) pair per window          [HBM-bound]
//   2. sort        : radix sort of the pairs by key (cub::DeviceRadixSort)      [HBM-bound]
//   3. accumulate  : every thread walks a fixed-length chunk of the sorted pairs, gathers
//                    the 64-byte affine points (128-bit loads, next point prefetched during
//                    the current add) and sums runs in XYZZ; runs closed inside the chunk
//                    are stored straight to their bucket, runs cut by a chunk edge go to a
//                    partial list                                               [IMAD-bound]
//   4. partials    : partial sums of one bucket are adjacent; the first one folds the rest
//   5. reduce      : per window, sum_b (b+1)*B_b by per-thread running sums over bucket
//                    segments + a short double-and-add for the segment offset   [IMAD-bound]
//   6. window sums : block tree-reduction of the segment results
//   7. combine     : Horner over the windows (c doublings each) + to-affine
//
// Chunking by entries (not by bucket) keeps the work per thread constant for any scalar
// distribution (all-equal scalars, 50% zeros, ...) -- SURVEY.md 8(d) config 1's adversarial sets.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cstring>
#include <functional>
#include <vector>
#include "ctx.cuh"
#include "host/nccl_dyn.hpp"
#include "msm_sort.cuh"

namespace bp {

static constexpr uint32_t INVALID_KEY = 0xFFFFFFFFu;
#ifndef BP_ACC_MIN_BLOCKS
#define BP_ACC_MIN_BLOCKS 4
#endif

struct MsmPlan {
    int c;            // window width in bits
    int W;            // number of windows
    uint32_t nb;      // buckets per window = 2^(c-1); bucket b holds digit magnitude b+1
    int key_bits;     // radix-sort end bit
    size_t entries;   // n * W
    int L;            // sorted entries per accumulate thread
    size_t T;         // accumulate threads
    uint32_t seg;     // buckets per reduce thread
    uint32_t nseg;    // segments per window
};

static inline MsmPlan make_plan(size_t n, int nmsm, int force_c, int sm_count) {
    // n = total number of (point, scalar) terms over all nmsm batched MSMs
    MsmPlan p;
    int best_c = 4;
    double best = 1e300;
    for (int c = 3; c <= 20; c++) {
        int W = 256 / c + 1;
        double nb = (double)(1u << (c - 1));
        double cost = W * (10.0 * (double)n + 45.0 * nb * nmsm);   // modmul-equivalents
        if (cost < best) { best = cost; best_c = c; }
    }
    p.c = force_c > 0 ? force_c : best_c;
    p.W = 256 / p.c + 1;
    p.nb = 1u << (p.c - 1);
    uint64_t nkeys = (uint64_t)nmsm * p.W * p.nb;
    p.key_bits = 1;
    while ((1ull << p.key_bits) <= nkeys) p.key_bits++;
    p.entries = n * (size_t)p.W;
    size_t per = p.entries / ((size_t)sm_count * 512);   // one resident wave (4 blocks x 128 threads per SM) before chunks grow
    // longer chunks at 2^24 (96 ... 256 pairs per thread) were measured: the slot levels shrink by as much as the accumulate
    // kernel's tail grows (40.22 -> 40.02 ... 40.10 ms)
    p.L = (int)(per < 8 ? 8 : per > 64 ? 64 : per);
    p.T = (p.entries + p.L - 1) / p.L;
    uint32_t seg = p.nb / 1024;
    p.seg = seg < 4 ? 4 : seg > 32 ? 32 : seg;
    // the largest windows still fill the machine with 64-bucket segments, and the per-segment lo*run multiplication
    // (~28 additions) is amortised over twice as many buckets
    if ((uint64_t)nmsm * p.W * p.nb / 64 >= (uint64_t)sm_count * 512) p.seg = 64;
    if (p.seg > p.nb) p.seg = p.nb;
    p.nseg = (p.nb + p.seg - 1) / p.seg;
    return p;
}

// A batch of up to MSM_MAX_BATCH independent MSMs given as up to MSM_MAX_SEGS segments; each
// segment pairs `count` scalars with `count` affine bases from its own array and belongs to one
// MSM of the batch. (A_I = <a_L,G> + <a_R,H> + i_bl*B_blinding is three segments of one MSM;
// the L and R of an IPA round are two MSMs of three segments each.)
static constexpr int MSM_MAX_SEGS = 8;
static constexpr int MSM_MAX_BATCH = 8;
static constexpr uint32_t MSM_IDX_MASK = 0x0FFFFFFFu;   // val = sign<<31 | seg<<28 | index
struct MsmJob {
    int nseg = 0;
    int nmsm = 1;
    const affine* bases[MSM_MAX_SEGS];
    const fe* scalars[MSM_MAX_SEGS];
    uint32_t count[MSM_MAX_SEGS];
    uint32_t start[MSM_MAX_SEGS + 1];   // prefix sums of count
    uint32_t msm[MSM_MAX_SEGS];
    uint32_t sstride[MSM_MAX_SEGS];     // scalar of term k of a segment = scalars[k * sstride] (cyclic generator shards read
                                        // replicated scalar vectors with stride = world)
    void add(const affine* b, const fe* s, size_t n, int m, uint32_t stride = 1) {
        bases[nseg] = b; scalars[nseg] = s; count[nseg] = (uint32_t)n; msm[nseg] = (uint32_t)m; sstride[nseg] = stride;
        start[nseg] = nseg ? start[nseg - 1] + count[nseg - 1] : 0;
        nseg++;
        start[nseg] = start[nseg - 1] + (uint32_t)n;
        if (m + 1 > nmsm) nmsm = m + 1;
    }
    size_t total() const { return nseg ? start[nseg] : 0; }
};

static inline SortSegs sort_segs_of(const MsmJob& job) {
    SortSegs g;
    g.nseg = job.nseg;
    for (int k = 0; k <= MSM_MAX_SEGS; k++) g.start[k] = k <= job.nseg ? job.start[k] : 0xFFFFFFFFu;
    return g;
}

// ---- 1. digits -------------------------------------------------------------------------------
// KEYS_ONLY (the bucket-sort path): the sign rides in bit 31 of the key and no value array is written -- the value of
// pair i is sign | segment | index, which the scatter kernel rebuilds from i (msm_sort.cuh); saves 4 of the 8 bytes per pair
// that this kernel writes and the scatter kernel reads.
template <class C, bool KEYS_ONLY = false>
__global__ void __launch_bounds__(256) msm_digits_kernel(const __grid_constant__ MsmJob job, size_t n, int c, int W,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int sg = 0;
#pragma unroll
    for (int k = 1; k < MSM_MAX_SEGS; k++)
        if (k < job.nseg && i >= job.start[k]) sg = k;
    uint32_t local = (uint32_t)i - job.start[sg];
    fe s = Fp<typename C::Fr>::from_mont(ld_fe(job.scalars[sg] + (size_t)local * job.sstride[sg]));
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1u;
    const uint32_t wbase = job.msm[sg] * (uint32_t)W;
    const uint32_t vbase = ((uint32_t)sg << 28) | local;
    uint32_t carry = 0;
    // the scalar streams through a 64-bit bit buffer, one limb at a time (static limb indices: indexing s.v[] by the window's
    // bit position costs two 8-way select chains per window); c <= 20, so the buffer never holds more than 51 bits
    uint64_t buf = 0;
    int have = 0, w = 0;
    auto emit = [&](uint32_t raw) {
        uint32_t d = raw + carry;
        uint32_t neg = 0;
        if (d > half) { d = (1u << c) - d; neg = 1; carry = 1; } else carry = 0;
        uint32_t key = d ? (((wbase + (uint32_t)w) << (c - 1)) | (d - 1u)) : INVALID_KEY;
        if (KEYS_ONLY) {
            keys[(size_t)w * n + i] = d ? (key | (neg << 31)) : INVALID_KEY;
        } else {
            keys[(size_t)w * n + i] = key;
            vals[(size_t)w * n + i] = vbase | (neg << 31);
        }
        w++;
    };
#pragma unroll
    for (int l = 0; l < 8; l++) {
        buf |= (uint64_t)s.v[l] << have;
        have += 32;
        while (have >= c && w < W) {
            emit((uint32_t)buf & mask);
            buf >>= c;
            have -= c;
        }
    }
    while (w < W) {                                  // the bits left over (256 is no multiple of c), then zeros
        emit((uint32_t)buf & mask);
        buf >>= c;
    }
}

// ---- 3. accumulate ---------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ affine gather_point(const MsmJob& job, uint32_t v) {
    affine p = ld_affine(job.bases[(v >> 28) & 7u] + (v & MSM_IDX_MASK));
    if (v >> 31) p = GroupLaw<C>::neg(p);
    return p;
}

// Run bookkeeping shared by all levels: the first run of a chunk goes to slot 0, the last run
// (if it is not also the first) to slot 1, runs strictly inside the chunk own their bucket.
// A single-run chunk fills slot 1 with (key, identity) so the slot list stays dense and sorted.
// MODE 0: buckets are written once per MSM (plain stores). MODE 1 / 2 serve the streamed MSM (msm_run_streamed), whose
// chunks all add into one bucket array: at level 1 (MODE 1) a run that owns its bucket *starts* from the bucket's current
// value, so the store below already carries it; a run that turns out to be cut by the chunk edge moves that value into
// its slot and leaves the identity behind. At the slot levels (MODE 2) complete sums are added to the bucket.
template <class C, int MODE = 0>
struct RunSink {
    using E = GroupLaw<C>;
    xyzz* buckets;
    uint32_t* out_keys;
    xyzz* out_pts;
    size_t t;
    int nruns = 0;
    __device__ __forceinline__ void flush(uint32_t key, const xyzz& acc, bool is_last) {
        if (nruns == 0) {
            out_keys[2 * t] = key;
            st_xyzz(out_pts + 2 * t, acc);
            if (is_last) {
                out_keys[2 * t + 1] = key;
                st_xyzz(out_pts + 2 * t + 1, E::identity());
            }
        } else if (is_last) {
            out_keys[2 * t + 1] = key;
            st_xyzz(out_pts + 2 * t + 1, acc);
            if (MODE == 1) st_xyzz(buckets + key, E::identity());
        } else if (MODE == 2) {
            xyzz b = ld_xyzz(buckets + key);
            E::add(b, acc);
            st_xyzz(buckets + key, b);
        } else {
            st_xyzz(buckets + key, acc);
        }
        nruns++;
    }
    __device__ __forceinline__ void empty() {
        out_keys[2 * t] = INVALID_KEY;
        out_keys[2 * t + 1] = INVALID_KEY;
    }
};

// level 1: sorted (key, point index|sign) pairs -> run sums of gathered affine points.
// Software pipeline (ncu: 12.5 % long-scoreboard stalls before it): while point i is being added, the
// gather of point i+1 is in flight and (key, val) of entry i+2 are being loaded; the sign is applied
// when a point is consumed, not when it is loaded, so nothing waits on a load it has just issued.
template <class C, bool ACC = false>
__global__ void __launch_bounds__(128, BP_ACC_MIN_BLOCKS) msm_accumulate_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                             size_t M, int L, size_t T, const __grid_constant__ MsmJob job,
                                                             xyzz* __restrict__ buckets, uint32_t* __restrict__ out_keys,
                                                             xyzz* __restrict__ out_pts) {
    using E = GroupLaw<C>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    size_t s = t * (size_t)L;
    size_t e = s + L < M ? s + L : M;
    RunSink<C, ACC ? 1 : 0> sink{buckets, out_keys, out_pts, t};
    uint32_t cur = __ldg(keys + s);
    if (cur == INVALID_KEY) { sink.empty(); return; }
    uint32_t vcur = __ldg(vals + s);
    // entry i+1
    uint32_t k1 = INVALID_KEY, v1 = 0;
    if (s + 1 < e) { k1 = __ldg(keys + s + 1); v1 = __ldg(vals + s + 1); }
    affine p = ld_affine(job.bases[(vcur >> 28) & 7u] + (vcur & MSM_IDX_MASK));
    xyzz acc = E::identity();
    size_t i = s;
    while (true) {
        // issue the gather of entry i+1 and the (key, val) loads of entry i+2
        const bool have_next = k1 != INVALID_KEY;
        affine pnext;
        if (have_next) pnext = ld_affine(job.bases[(v1 >> 28) & 7u] + (v1 & MSM_IDX_MASK));
        uint32_t k2 = INVALID_KEY, v2 = 0;
        if (i + 2 < e) { k2 = __ldg(keys + i + 2); v2 = __ldg(vals + i + 2); }
        // streamed MSM: the next run starts from its bucket's sum of the earlier chunks (4-32 terms per run and chunk);
        // its 128-byte line is fetched while the addition below runs, or every run would begin with an exposed DRAM load
        // (bp_msm at 2^24: 47.0 -> 46.4 ms with 2^21-point chunks; reading the keys one entry further ahead to prefetch
        // two additions early was measured too and is no faster)
        if (ACC && have_next && k1 != cur) asm volatile("prefetch.global.L1 [%0];" ::"l"(buckets + k1));
        if (vcur >> 31) p = E::neg(p);
        E::madd(acc, p);
        if (k1 != cur) {
            sink.flush(cur, acc, !have_next);
            if (!have_next) break;
            cur = k1;
            if (ACC) acc = ld_xyzz(buckets + cur);     // sums of the earlier chunks (streamed MSM)
            else acc = E::identity();
        }
        p = pnext;
        vcur = v1;
        k1 = k2;
        v1 = v2;
        i++;
    }
}

// ---- 3a. batched-affine pre-addition ("pair rounds") ---------------------------------------------------------------
// A mixed XYZZ addition costs 8M + 2S; an affine addition costs one inversion + 2M + 1S, and Montgomery's trick shares
// the inversion: 3 more multiplications per addition plus one Fermat inversion (~450 modmul on secq256k1) per batch --
// 6 + 450/K modmul per addition. Before the XYZZ accumulation, round r = 0, 1, ... adds the neighbours (i, i + 2^r) of
// the *sorted* pair list that lie in the same bucket, K pairs per thread sharing one inversion:
//   pass A  walks the thread's pairs forward:  dx_j = x_Q - x_P, prefix products into a coalesced scratch column;
//   pass B  walks them backward: 1/dx_j from the running inverse, lambda, (x3, y3) -> pairpts[i / 2].
// The list keeps its length and its keys: the left entry's value becomes a reference to pairpts (segment 7 of the job),
// the right one a reference to a zero point, which the accumulate kernel's mixed addition skips on its identity test --
// so msm_accumulate_kernel runs unchanged on a list whose real entries have halved per round. Pairs the affine law
// cannot add (P = +-Q: dx = 0; an identity base; x = 0) are simply left for the XYZZ accumulation, which handles every
// exceptional case. Short-Weierstrass curves only (the twisted Edwards law has no cheap affine form).
// Pairs are dealt to the threads round-robin (pair p = j * T + t): neighbouring lanes read neighbouring pairs (coalesced
// keys, values and prefix products), every thread gets the same number of pairs K = ceil(npairs / T), and T is a whole
// number of resident waves -- the first version (contiguous 512-pair chunks per thread) ran 2.25 waves of long threads at
// 2^24 points and left the multiplier 47 % busy (profiles/r2_ncu_pair_affine_lg22_v1.txt).
template <class C>
__global__ void __launch_bounds__(128, 5) msm_pair_affine_kernel(const uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, size_t M, uint32_t stride,
                                                                 size_t npairs, int K, size_t T, const __grid_constant__ MsmJob job,
                                                                 affine* __restrict__ pairpts, fe* __restrict__ pre, uint32_t hole_val) {
    using F = Fp<typename C::Fq>;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    auto src = [&](uint32_t v) -> const affine* { return job.bases[(v >> 28) & 7u] + (v & MSM_IDX_MASK); };
    // dx of pair j, or false when the pair is not added in this round (identical decisions in both passes)
    auto pair_dx = [&](int j, size_t& left, uint32_t& v0, uint32_t& v1, fe& px, fe& qx, fe& dx) -> bool {
        const size_t pidx = (size_t)j * T + t;
        if (pidx >= npairs) return false;
        left = pidx * 2 * stride;
        const size_t right = left + stride;
        if (right >= M) return false;
        const uint32_t k0 = keys[left];
        if (k0 == INVALID_KEY || k0 != keys[right]) return false;
        v0 = vals[left];
        v1 = vals[right];
        if (v0 == hole_val || v1 == hole_val) return false;
        px = ld_fe_rw(&src(v0)->x);
        qx = ld_fe_rw(&src(v1)->x);
        if (F::is_zero(px) || F::is_zero(qx)) return false;      // the identity (0,0), or a point with x = 0: left to the XYZZ path
        dx = F::sub(qx, px);
        return !F::is_zero(dx);                                  // P = +-Q
    };
    fe prod = F::one();
#pragma unroll 1
    for (int j = 0; j < K; j++) {
        size_t left;
        uint32_t v0, v1;
        fe px, qx, dx;
        if (!pair_dx(j, left, v0, v1, px, qx, dx)) continue;
        st_fe(pre + (size_t)j * T + t, prod);
        prod = F::mul(prod, dx);
    }
    fe inv = F::inv(prod);
#pragma unroll 1
    for (int j = K - 1; j >= 0; j--) {
        size_t left;
        uint32_t v0, v1;
        fe px, qx, dx;
        if (!pair_dx(j, left, v0, v1, px, qx, dx)) continue;
        fe py = ld_fe_rw(&src(v0)->y), qy = ld_fe_rw(&src(v1)->y);
        if (v0 >> 31) py = F::neg(py);
        if (v1 >> 31) qy = F::neg(qy);
        const fe invj = F::mul(inv, ld_fe_rw(pre + (size_t)j * T + t));
        inv = F::mul(inv, dx);
        const fe lam = F::mul(F::sub(qy, py), invj);
        const fe x3 = F::sub(F::sub(F::sqr(lam), px), qx);
        const fe y3 = F::sub(F::mul(lam, F::sub(px, x3)), py);
        affine* o = pairpts + (left >> 1);
        st_fe(&o->x, x3);
        st_fe(&o->y, y3);
        vals[left] = (7u << 28) | (uint32_t)(left >> 1);
        vals[left + stride] = hole_val;
    }
}

// level >= 2: dense sorted (key, XYZZ partial) slots -> run sums
template <class C, bool ACC = false>
__global__ void __launch_bounds__(128) msm_partials_level_kernel(const uint32_t* __restrict__ in_keys, const xyzz* __restrict__ in_pts,
                                                                 size_t nslots, int L, size_t T, xyzz* __restrict__ buckets,
                                                                 uint32_t* __restrict__ out_keys, xyzz* __restrict__ out_pts) {
    using E = GroupLaw<C>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    size_t s = t * (size_t)L;
    size_t e = s + L < nslots ? s + L : nslots;
    RunSink<C, ACC ? 2 : 0> sink{buckets, out_keys, out_pts, t};
    uint32_t cur = in_keys[s];
    if (cur == INVALID_KEY) { sink.empty(); return; }
    xyzz acc = ld_xyzz(in_pts + s);
    for (size_t i = s + 1; i < e; i++) {
        uint32_t k = in_keys[i];
        if (k == INVALID_KEY) break;
        xyzz v = ld_xyzz(in_pts + i);
        if (k != cur) {
            sink.flush(cur, acc, false);
            cur = k;
            acc = v;
        } else {
            E::add(acc, v);
        }
    }
    sink.flush(cur, acc, true);
}

// Warp-segmented level for short slot lists (small MSMs are latency-bound: the serial 16-slot chains of
// msm_partials_level_kernel cost ~0.1 ms per level, and an IPA runs ~2 log n such MSMs back to back).
// One slot per lane; a segmented Hillis-Steele scan over the sorted keys sums every run in 5 dependent
// additions. Runs strictly inside the warp are complete (their key occurs nowhere else) and go to their
// bucket; the run touching lane 0 and the run touching lane 31 go to the two output slots of the warp,
// so a level shrinks the list 16x. With `final` set every run is complete and goes to its bucket.
template <class C>
__device__ __forceinline__ xyzz shfl_up_xyzz(const xyzz& v, int d) {
    xyzz r;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        r.x.v[k] = __shfl_up_sync(0xFFFFFFFFu, v.x.v[k], d);
        r.y.v[k] = __shfl_up_sync(0xFFFFFFFFu, v.y.v[k], d);
        r.zz.v[k] = __shfl_up_sync(0xFFFFFFFFu, v.zz.v[k], d);
        r.zzz.v[k] = __shfl_up_sync(0xFFFFFFFFu, v.zzz.v[k], d);
    }
    return r;
}

template <class C, bool ACC = false>
__global__ void __launch_bounds__(128) msm_partials_warp_kernel(const uint32_t* __restrict__ in_keys, const xyzz* __restrict__ in_pts,
                                                                size_t nslots, int final, xyzz* __restrict__ buckets,
                                                                uint32_t* __restrict__ out_keys, xyzz* __restrict__ out_pts) {
    using E = GroupLaw<C>;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t warp = p >> 5;
    const int lane = (int)(threadIdx.x & 31u);
    if ((warp << 5) >= nslots) return;                       // whole warp out of range
    uint32_t key = p < nslots ? in_keys[p] : INVALID_KEY;
    xyzz acc = E::identity();
    if (key != INVALID_KEY) acc = ld_xyzz(in_pts + p);
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t ok = __shfl_up_sync(0xFFFFFFFFu, key, d);
        xyzz o = shfl_up_xyzz<C>(acc, d);
        if (lane >= d && ok == key && key != INVALID_KEY) E::add(acc, o);
    }
    const uint32_t knext = __shfl_down_sync(0xFFFFFFFFu, key, 1);
    const uint32_t kfirst = __shfl_sync(0xFFFFFFFFu, key, 0);
    const uint32_t klast = __shfl_sync(0xFFFFFFFFu, key, 31);
    const bool run_end = lane == 31 || knext != key;
    if (final) {
        if (run_end && key != INVALID_KEY) {
            if (ACC) { xyzz b = ld_xyzz(buckets + key); E::add(b, acc); st_xyzz(buckets + key, b); }
            else st_xyzz(buckets + key, acc);
        }
        return;
    }
    if (lane == 0 && kfirst == klast) {                      // single run: keep the slot list dense
        out_keys[2 * warp + 1] = kfirst;
        st_xyzz(out_pts + 2 * warp + 1, E::identity());
    }
    if (!run_end) return;
    if (key == kfirst) {
        out_keys[2 * warp] = key;
        if (key != INVALID_KEY) st_xyzz(out_pts + 2 * warp, acc);
    } else if (key == klast) {
        out_keys[2 * warp + 1] = key;
        if (key != INVALID_KEY) st_xyzz(out_pts + 2 * warp + 1, acc);
    } else if (ACC) {
        xyzz b = ld_xyzz(buckets + key);
        E::add(b, acc);
        st_xyzz(buckets + key, b);
    } else {
        st_xyzz(buckets + key, acc);
    }
}

// last level: one thread per slot; the first slot of each key folds the following ones
template <class C>
__global__ void __launch_bounds__(128) msm_partials_final_kernel(const uint32_t* __restrict__ keys, const xyzz* __restrict__ pts,
                                                                 size_t nslots, xyzz* __restrict__ buckets) {
    using E = GroupLaw<C>;
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nslots) return;
    uint32_t k = keys[p];
    if (k == INVALID_KEY) return;
    if (p > 0 && keys[p - 1] == k) return;
    xyzz acc = ld_xyzz(pts + p);
    for (size_t q = p + 1; q < nslots && keys[q] == k; q++) {
        xyzz o = ld_xyzz(pts + q);
        E::add(acc, o);
    }
    st_xyzz(buckets + k, acc);
}

// ---- 5. bucket reduction ---------------------------------------------------------------------
// PAIR = false: seg_out[t] = sum over the segment of (b + 1) * B_b (the segment offset lo * run by double-and-add).
// PAIR = true (large windows): the offset is left to the next level -- seg_out[t] = sum (b - lo + 1) * B_b and
// seg_run[t] = sum B_b; msm_window_partial_kernel adds seg * sum_s s * run_s, a weighted sum over nseg values instead of a
// 13-bit double-and-add in every one of the nseg threads (11 % of this kernel's modmul at 64-bucket segments).
template <class C, bool PAIR = false>
__global__ void __launch_bounds__(128) msm_reduce_kernel(const xyzz* __restrict__ buckets, uint32_t nb, uint32_t seg, uint32_t nseg,
                                                         int W, xyzz* __restrict__ seg_out, xyzz* __restrict__ seg_run = nullptr) {
    using E = GroupLaw<C>;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)W * nseg) return;
    uint32_t w = (uint32_t)(t / nseg), sg = (uint32_t)(t % nseg);
    uint32_t lo = sg * seg;
    uint32_t hi = lo + seg < nb ? lo + seg : nb;
    const xyzz* B = buckets + (size_t)w * nb;
    xyzz run = E::identity(), acc = E::identity();
    xyzz v = ld_xyzz(B + hi - 1);
    for (uint32_t b = hi; b-- > lo;) {
        xyzz nxt = v;
        if (b > lo) nxt = ld_xyzz(B + b - 1);     // the next bucket is in flight during the two additions
        E::add(run, v);
        E::add(acc, run);
        v = nxt;
    }
    if (PAIR) {
        st_xyzz(seg_run + t, run);
    } else if (lo != 0 && !E::is_identity(run)) {
        // acc = sum (b - lo + 1) B_b ; weights are b + 1  ->  add lo * run
        xyzz m = E::mul_u32(run, lo);
        E::add(acc, m);
    }
    st_xyzz(seg_out + t, acc);
}

// Second level of the PAIR reduction: block (slice, w) takes a contiguous range of window w's segments, every thread `per`
// consecutive ones: sum_s acc_s + seg * sum_s s * run_s over the range (running sums inside the thread, one short
// double-and-add for the thread's first segment index), tree-summed over the block into part_out[w * slices + slice].
template <class C>
__global__ void __launch_bounds__(128) msm_window_partial_kernel(const xyzz* __restrict__ seg_acc, const xyzz* __restrict__ seg_run, uint32_t nseg,
                                                                 uint32_t seg, uint32_t per, xyzz* __restrict__ part_out) {
    using E = GroupLaw<C>;
    __shared__ xyzz sh[128];
    const uint32_t w = blockIdx.y, slices = gridDim.x;
    const uint32_t s0 = (blockIdx.x * blockDim.x + threadIdx.x) * per;
    xyzz val = E::identity();
    if (s0 < nseg) {
        const uint32_t s1 = s0 + per < nseg ? s0 + per : nseg;
        const xyzz* A = seg_acc + (size_t)w * nseg;
        const xyzz* R = seg_run + (size_t)w * nseg;
        xyzz r = E::identity(), a = E::identity();
        for (uint32_t sidx = s1; sidx-- > s0;) {
            xyzz rv = ld_xyzz(R + sidx);
            E::add(r, rv);
            E::add(a, r);                                  // a = sum (sidx - s0 + 1) * run
            xyzz av = ld_xyzz(A + sidx);
            E::add(val, av);
        }
        // sum sidx * run = a + (s0 - 1) * r
        if (s0 == 0) { xyzz nr = E::neg(r); E::add(a, nr); }
        else if (s0 > 1 && !E::is_identity(r)) { xyzz m = E::mul_u32(r, s0 - 1); E::add(a, m); }
        if (!E::is_identity(a)) { xyzz m = E::mul_u32(a, seg); E::add(val, m); }
    }
    sh[threadIdx.x] = val;
    __syncthreads();
    for (int stride = 64; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            xyzz x = sh[threadIdx.x];
            xyzz y = sh[threadIdx.x + stride];
            E::add(x, y);
            sh[threadIdx.x] = x;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(part_out + (size_t)w * slices + blockIdx.x, sh[0]);
}

// ---- 6. window sums --------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) msm_window_sum_kernel(const xyzz* __restrict__ seg_out, uint32_t nseg, xyzz* __restrict__ win_out) {
    using E = GroupLaw<C>;
    __shared__ xyzz sh[128];
    uint32_t w = blockIdx.x;
    xyzz acc = E::identity();
    for (uint32_t s = threadIdx.x; s < nseg; s += blockDim.x) {
        xyzz v = ld_xyzz(seg_out + (size_t)w * nseg + s);
        E::add(acc, v);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int stride = 64; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            xyzz a = sh[threadIdx.x];
            xyzz b = sh[threadIdx.x + stride];
            E::add(a, b);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(win_out + w, sh[0]);
}

// ---- tiny MSMs ---------------------------------------------------------------------------------
// Below ~256 terms per MSM the bucket pipeline is nothing but launch and dependency latency (0.45-0.55 ms for any size,
// and a small proof runs a dozen of them: every IPA round, the T commitments, the blinding-only phase-1 commitments).
// One launch instead: block (w, m) handles the 4-bit window w of MSM m -- every thread multiplies its (up to three)
// bases by their digits (<= 15: at most 4 doublings + 4 additions) and the block tree-sums the products in shared
// memory (8 levels).
// The 64 window sums per MSM go through the same host Horner as the bucket path (c = 4).
static constexpr int MSM_TINY_C = 4, MSM_TINY_W = 64, MSM_TINY_THREADS = 256;

template <class C>
__global__ void __launch_bounds__(MSM_TINY_THREADS) msm_tiny_kernel(const __grid_constant__ MsmJob job, xyzz* __restrict__ win_out) {
    using E = GroupLaw<C>;
    __shared__ xyzz sh[MSM_TINY_THREADS];
    const int w = blockIdx.x, m = blockIdx.y;
    // terms threadIdx.x, threadIdx.x + 256, ... of MSM m: walk the segments that belong to m
    int total = 0;
    for (int sg = 0; sg < job.nseg; sg++)
        if ((int)job.msm[sg] == m) total += (int)job.count[sg];
    xyzz acc = E::identity();
#pragma unroll 1
    for (int term = (int)threadIdx.x; term < total; term += MSM_TINY_THREADS) {
        int left = term;
#pragma unroll 1
        for (int sg = 0; sg < job.nseg; sg++) {
            if ((int)job.msm[sg] != m) continue;
            if (left < (int)job.count[sg]) {
                fe s = Fp<typename C::Fr>::from_mont(ld_fe(job.scalars[sg] + (size_t)left * job.sstride[sg]));
                const uint32_t d = (s.v[w >> 3] >> (4 * (w & 7))) & 15u;
                if (d) {
                    affine p = ld_affine(job.bases[sg] + left);
                    xyzz q = E::mul_u32(E::from_affine(p), d);
                    E::add(acc, q);
                }
                break;
            }
            left -= (int)job.count[sg];
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int stride = MSM_TINY_THREADS / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            xyzz a = sh[threadIdx.x];
            xyzz b = sh[threadIdx.x + stride];
            E::add(a, b);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(win_out + (size_t)m * MSM_TINY_W + w, sh[0]);
}

// synthetic workload: out[i] = (start + i + 1) * G
template <class C>
__global__ void __launch_bounds__(128) synth_points_kernel(affine* __restrict__ out, size_t n, uint64_t start) {
    using E = GroupLaw<C>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine g;
    for (int k = 0; k < 8; k++) { g.x.v[k] = C::gx(k); g.y.v[k] = C::gy(k); }
    uint64_t k = start + i + 1;
    uint32_t s[8] = {(uint32_t)k, (uint32_t)(k >> 32), 0, 0, 0, 0, 0, 0};
    xyzz acc = E::identity();
    for (int bit = 63; bit >= 0; bit--) {
        acc = E::dbl(acc);
        if ((s[bit >> 5] >> (bit & 31)) & 1u) E::madd(acc, g);
    }
    affine a = E::to_affine(acc);
    st_fe(&out[i].x, a.x);
    st_fe(&out[i].y, a.y);
}

// ---- host driver -----------------------------------------------------------------------------
int host_combine(int curve, const void* win, int W, int c, uint8_t out_xy[64], int* out_is_identity);

// Hierarchical reduction of the level-1 slot list (2 slots per accumulate thread, at the start of part_keys / part_pts;
// the levels ping-pong between that region and the one behind it). Long lists: serial 16-slot chunks (throughput);
// short lists: warp-segmented scans (latency).
static constexpr int MSM_PL = 16;
// pk / pp: a region of slots1 + 2 * ceil(slots1 / PL) + 64 slots whose first slots1 entries hold the level-1 list
template <class C, bool ACC>
int msm_fold_slots_at(bp_ctx* ctx, uint32_t* pk, xyzz* pp, size_t slots1, cudaStream_t st) {
    const int PL = MSM_PL;
    size_t nslots = slots1;
    uint32_t* in_k = pk;
    xyzz* in_p = pp;
    uint32_t* out_k = pk + slots1;
    xyzz* out_p = pp + slots1;
    size_t warp_below = ctx->msm_warp_partials_below;
    // sweep knob of the streamed MSM's per-chunk slot levels (2^24-point bp_msm: 45.2 / 45.1 / 45.2 / 45.1 ms for 2^16 / 2^17 /
    // 2^18 / 2^19, 45.9 ms for 2^15: flat, the default stays)
    if (ACC) { if (const char* e = getenv("BP_MSM_STREAM_WARP_BELOW")) { size_t v = strtoull(e, nullptr, 10); if (v) warp_below = v; } }
    while (nslots > warp_below) {
        size_t T2 = (nslots + PL - 1) / PL;
        msm_partials_level_kernel<C, ACC><<<(unsigned)((T2 + 127) / 128), 128, 0, st>>>(in_k, in_p, nslots, PL, T2, ctx->buckets.as<xyzz>(),
                                                                                        out_k, out_p);
        BP_LAUNCH_CHECK(ctx);
        nslots = 2 * T2;
        uint32_t* tk = in_k; in_k = out_k; out_k = tk;
        xyzz* tp = in_p; in_p = out_p; out_p = tp;
    }
    while (nslots > 32) {
        size_t nw = (nslots + 31) / 32;
        msm_partials_warp_kernel<C, ACC><<<(unsigned)((nslots + 127) / 128), 128, 0, st>>>(in_k, in_p, nslots, 0, ctx->buckets.as<xyzz>(), out_k, out_p);
        BP_LAUNCH_CHECK(ctx);
        nslots = 2 * nw;
        uint32_t* tk = in_k; in_k = out_k; out_k = tk;
        xyzz* tp = in_p; in_p = out_p; out_p = tp;
    }
    msm_partials_warp_kernel<C, ACC><<<1, 32, 0, st>>>(in_k, in_p, nslots, 1, ctx->buckets.as<xyzz>(), out_k, out_p);
    BP_LAUNCH_CHECK(ctx);
    return BP_OK;
}
template <class C, bool ACC>
int msm_fold_slots(bp_ctx* ctx, size_t slots1, cudaStream_t st) {
    return msm_fold_slots_at<C, ACC>(ctx, ctx->part_keys.as<uint32_t>(), ctx->part_pts.as<xyzz>(), slots1, st);
}
static inline size_t msm_fold_region(size_t slots1) { return slots1 + 2 * ((slots1 + MSM_PL - 1) / MSM_PL) + 64; }

// Bucket reduction + window sums of NW windows into ctx->win_out. Large windows (>= MSM_PAIR_MIN_NSEG segments) take the
// two-level PAIR path: reduce (acc, run) per segment -> per-slice partial sums with the segment offsets -> window sums.
static constexpr uint32_t MSM_PAIR_MIN_NSEG = 1024;
static constexpr uint32_t MSM_PAIR_SLICES = 16;
template <class C>
int msm_reduce_windows(bp_ctx* ctx, const MsmPlan& p, int NW, cudaStream_t st) {
    const size_t rt = (size_t)NW * p.nseg;
    if (ctx->msm_pair_reduce && p.nseg >= MSM_PAIR_MIN_NSEG) {
        BP_CUDA_TRY(ctx, ctx->seg_out.reserve((2 * rt + (size_t)NW * MSM_PAIR_SLICES) * sizeof(xyzz)));
        xyzz* seg_acc = ctx->seg_out.as<xyzz>();
        xyzz* seg_run = seg_acc + rt;
        xyzz* part = seg_run + rt;
        msm_reduce_kernel<C, true><<<(unsigned)((rt + 127) / 128), 128, 0, st>>>(ctx->buckets.as<xyzz>(), p.nb, p.seg, p.nseg, NW, seg_acc, seg_run);
        BP_LAUNCH_CHECK(ctx);
        // The second level is a chain of ~50 dependent additions per thread: pure latency, and a warp that has its
        // scheduler to itself runs it twice as fast as two sharing one -- so at most one 4-warp block per SM
        // (2^24 points: 13 windows x 11 slices = 143 blocks on 148 SMs).
        uint32_t slices = (uint32_t)(ctx->sm_count / NW);
        if (slices < 1) slices = 1;
        if (slices > MSM_PAIR_SLICES) slices = MSM_PAIR_SLICES;
        const uint32_t per = (p.nseg + slices * 128 - 1) / (slices * 128);
        msm_window_partial_kernel<C><<<dim3(slices, NW), 128, 0, st>>>(seg_acc, seg_run, p.nseg, p.seg, per, part);
        BP_LAUNCH_CHECK(ctx);
        msm_window_sum_kernel<C><<<NW, 128, 0, st>>>(part, slices, ctx->win_out.as<xyzz>());
        BP_LAUNCH_CHECK(ctx);
        return BP_OK;
    }
    msm_reduce_kernel<C, false><<<(unsigned)((rt + 127) / 128), 128, 0, st>>>(ctx->buckets.as<xyzz>(), p.nb, p.seg, p.nseg, NW,
                                                                             ctx->seg_out.as<xyzz>());
    BP_LAUNCH_CHECK(ctx);
    msm_window_sum_kernel<C><<<NW, 128, 0, st>>>(ctx->seg_out.as<xyzz>(), p.nseg, ctx->win_out.as<xyzz>());
    BP_LAUNCH_CHECK(ctx);
    return BP_OK;
}

// Batched-affine pair rounds over a sorted pair list (see msm_pair_affine_kernel); `job` gains segment 7 = the pair points.
template <class C>
int msm_pair_rounds(bp_ctx* ctx, MsmJob& job, const uint32_t* keys, uint32_t* vals, size_t entries, cudaStream_t st) {
    if (C::KIND != 0 || ctx->msm_affine_rounds <= 0 || job.nseg > 7 || entries < ctx->msm_affine_min_entries) return BP_OK;
    const size_t npairs = (entries + 1) / 2;
    if (npairs + 1 > MSM_IDX_MASK) return BP_OK;
    BP_CUDA_TRY(ctx, ctx->pairpts.reserve((npairs + 1) * sizeof(affine)));
    BP_CUDA_TRY(ctx, ctx->pairpre.reserve(npairs * sizeof(fe)));
    affine* pp = ctx->pairpts.as<affine>();
    BP_CUDA_TRY(ctx, cudaMemsetAsync(pp + npairs, 0, sizeof(affine), st));       // the zero point every hole refers to
    job.bases[7] = pp;
    const uint32_t hole = (7u << 28) | (uint32_t)npairs;
    for (int r = 0; r < ctx->msm_affine_rounds; r++) {
        const uint32_t stride = 1u << r;
        const size_t np = (entries + 2 * (size_t)stride - 1) / (2 * (size_t)stride);      // pairs (i, i + stride), i = 2 * stride * p
        // whole resident waves (5 blocks of 128 threads per SM), at least ~256 pairs per thread to amortise the inversion
        size_t T = (size_t)ctx->sm_count * 5 * 128;
        while (T * 2 * 320 <= np) T *= 2;
        if (T > np) T = np;
        const int K = (int)((np + T - 1) / T);
        msm_pair_affine_kernel<C><<<(unsigned)((T + 127) / 128), 128, 0, st>>>(keys, vals, entries, stride, np, K, T, job, pp, ctx->pairpre.as<fe>(), hole);
        BP_LAUNCH_CHECK(ctx);
    }
    return BP_OK;
}

// fn(0) ... fn(n - 1) in parallel on the context's persistent host workers (host/workers.hpp); fn(0) on the caller
template <class Fn>
static inline void host_parallel(bp_ctx* ctx, int n, Fn&& fn) {
    if (n <= 1) { if (n == 1) fn(0); return; }
    if (!ctx->workers) ctx->workers = new HostWorkers(MSM_MAX_BATCH - 1);
    ctx->workers->run(n, std::function<void(int)>(fn));
}

// Runs the batch; out_xy / out_is_identity have job.nmsm entries.
template <class C>
int msm_run_job(bp_ctx* ctx, const MsmJob& job, uint8_t (*out_xy)[64], int* out_is_identity) {
    cudaStream_t st = ctx->stream;
    size_t n = job.total();
    if (n == 0) {
        for (int m = 0; m < job.nmsm; m++) { memset(out_xy[m], 0, 64); if (out_is_identity) out_is_identity[m] = 1; }
        return BP_OK;
    }
    for (int k = 0; k < job.nseg; k++)
        if (job.count[k] > MSM_IDX_MASK) return BP_ERR_LEN;
    if (n >= (1ull << 31)) return BP_ERR_LEN;
    // tiny batches: one launch (msm_tiny_kernel)
    if (ctx->force_c == 0 && ctx->msm_tiny_max > 0) {
        uint32_t per[MSM_MAX_BATCH] = {0}, worst = 0;
        for (int k = 0; k < job.nseg; k++) per[job.msm[k]] += job.count[k];
        for (int m = 0; m < job.nmsm; m++) worst = per[m] > worst ? per[m] : worst;
        if (worst <= (uint32_t)ctx->msm_tiny_max) {
            const int NWt = job.nmsm * MSM_TINY_W;
            BP_CUDA_TRY(ctx, ctx->win_out.reserve((size_t)NWt * sizeof(xyzz)));
            if ((size_t)NWt * sizeof(xyzz) > BP_HOST_RESULT_BYTES) return BP_ERR_LEN;
            ctx->last_c = MSM_TINY_C; ctx->last_W = MSM_TINY_W; ctx->last_entries = n * MSM_TINY_W;
            msm_tiny_kernel<C><<<dim3(MSM_TINY_W, job.nmsm), MSM_TINY_THREADS, 0, st>>>(job, ctx->win_out.as<xyzz>());
            BP_LAUNCH_CHECK(ctx);
            BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_result, ctx->win_out.p, (size_t)NWt * sizeof(xyzz), cudaMemcpyDeviceToHost, st));
            BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            if (ctx->timing) for (int i = 0; i < 5; i++) ctx->phase_ms[i] = 0;
            int rcs[MSM_MAX_BATCH] = {0};
            auto combine = [&](int m) {
                rcs[m] = host_combine(ctx->curve, (const xyzz*)ctx->h_result + (size_t)m * MSM_TINY_W, MSM_TINY_W, MSM_TINY_C, out_xy[m],
                                      out_is_identity ? &out_is_identity[m] : nullptr);
            };
            host_parallel(ctx, job.nmsm, combine);
            for (int m = 0; m < job.nmsm; m++)
                if (rcs[m] != BP_OK) return rcs[m];
            return BP_OK;
        }
    }
    MsmPlan p = make_plan(n, job.nmsm, ctx->force_c, ctx->sm_count);
    if (C::KIND == 0 && ctx->msm_affine_rounds > 0 && job.nseg <= 7 && p.entries >= ctx->msm_affine_min_entries) {
        // after r pair rounds only ~1/2^r of the entries are real additions: keep the work per accumulate thread
        int L2 = p.L << ctx->msm_affine_rounds;
        p.L = L2 > 512 ? 512 : L2;
        p.T = (p.entries + p.L - 1) / p.L;
    }
    const int NW = job.nmsm * p.W;   // windows over the whole batch
    BP_CUDA_TRY(ctx, ctx->keys_a.reserve(p.entries * 4));
    BP_CUDA_TRY(ctx, ctx->keys_b.reserve(p.entries * 4));
    BP_CUDA_TRY(ctx, ctx->vals_a.reserve(p.entries * 4));
    BP_CUDA_TRY(ctx, ctx->vals_b.reserve(p.entries * 4));
    size_t nbuckets = (size_t)NW * p.nb;
    BP_CUDA_TRY(ctx, ctx->buckets.reserve(nbuckets * sizeof(xyzz)));
    // partial slot lists: level 1 has 2T slots, every further level shrinks by PL/2
    const int PL = MSM_PL;
    size_t slots1 = 2 * p.T;
    size_t slots2 = 2 * ((slots1 + PL - 1) / PL);
    BP_CUDA_TRY(ctx, ctx->part_keys.reserve((slots1 + slots2 + 64) * 4));
    BP_CUDA_TRY(ctx, ctx->part_pts.reserve((slots1 + slots2 + 64) * sizeof(xyzz)));
    BP_CUDA_TRY(ctx, ctx->seg_out.reserve((2 * (size_t)NW * p.nseg + (size_t)NW * MSM_PAIR_SLICES) * sizeof(xyzz)));
    BP_CUDA_TRY(ctx, ctx->win_out.reserve((size_t)NW * sizeof(xyzz)));
    if ((size_t)NW * sizeof(xyzz) > BP_HOST_RESULT_BYTES) return BP_ERR_LEN;
    size_t tmp_bytes = 0;
    BP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->keys_a.as<uint32_t>(), ctx->keys_b.as<uint32_t>(),
                                                     ctx->vals_a.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), p.entries, 0,
                                                     p.key_bits, st));
    BP_CUDA_TRY(ctx, ctx->cub_tmp.reserve(tmp_bytes));

    SortPlan splan;
    if (ctx->msm_sort_mode == 1 && job.nmsm == 1 && p.entries >= ctx->msm_sort_min_entries && !(C::KIND == 0 && ctx->msm_affine_rounds > 0)) {
        splan = make_sort_plan(n, p.W, p.c - 1);
        if (splan.ok) BP_CUDA_TRY(ctx, ctx->sort_scratch.reserve(sort_scratch_bytes(splan)));
    }
    ctx->last_c = p.c; ctx->last_W = p.W; ctx->last_entries = p.entries;
    auto mark = [&](int i) { if (ctx->timing) cudaEventRecord(ctx->ev[i], st); };
    mark(0);
    if (splan.ok)
        msm_digits_kernel<C, true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(job, n, p.c, p.W, ctx->keys_a.as<uint32_t>(), nullptr);
    else
        msm_digits_kernel<C><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(job, n, p.c, p.W, ctx->keys_a.as<uint32_t>(),
                                                                         ctx->vals_a.as<uint32_t>());
    BP_LAUNCH_CHECK(ctx);
    mark(7);
    if (splan.ok) {
        int nl = 0;
        BP_CUDA_TRY(ctx, sort_pairs_run(splan, sort_segs_of(job), ctx->keys_a.as<uint32_t>(), ctx->keys_b.as<uint32_t>(), ctx->vals_b.as<uint32_t>(),
                                        ctx->sort_scratch.p, st, &nl));
        ctx->launches += nl;
    } else {
        BP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, ctx->keys_a.as<uint32_t>(), ctx->keys_b.as<uint32_t>(),
                                                         ctx->vals_a.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), p.entries, 0,
                                                         p.key_bits, st));
    }
    mark(1);
    BP_CUDA_TRY(ctx, cudaMemsetAsync(ctx->buckets.p, 0, nbuckets * sizeof(xyzz), st));
    uint32_t* pk = ctx->part_keys.as<uint32_t>();
    xyzz* pp = ctx->part_pts.as<xyzz>();
    MsmJob ajob = job;
    if (int rc = msm_pair_rounds<C>(ctx, ajob, ctx->keys_b.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), p.entries, st)) return rc;
    msm_accumulate_kernel<C><<<(unsigned)((p.T + 127) / 128), 128, 0, st>>>(ctx->keys_b.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), p.entries, p.L,
                                                                            p.T, ajob, ctx->buckets.as<xyzz>(), pk, pp);
    BP_LAUNCH_CHECK(ctx);
    mark(2);
    if (int rc = msm_fold_slots<C, false>(ctx, slots1, st)) return rc;
    mark(3);
    if (int rc = msm_reduce_windows<C>(ctx, p, NW, st)) return rc;
    mark(4);
    // The Horner combination over windows is 256 dependent doublings: 1.35 ms on one GPU
    // thread (measured), ~0.1 ms on a host core. The W window sums (W*128 B) go to the host.
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_result, ctx->win_out.p, (size_t)NW * sizeof(xyzz), cudaMemcpyDeviceToHost, st));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (ctx->timing) {
        // phases: 0 digits, 1 sort, 2 memset+accumulate, 3 partial levels, 4 bucket reduce + window sums
        cudaEventElapsedTime(&ctx->phase_ms[0], ctx->ev[0], ctx->ev[7]);
        cudaEventElapsedTime(&ctx->phase_ms[1], ctx->ev[7], ctx->ev[1]);
        cudaEventElapsedTime(&ctx->phase_ms[2], ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&ctx->phase_ms[3], ctx->ev[2], ctx->ev[3]);
        cudaEventElapsedTime(&ctx->phase_ms[4], ctx->ev[3], ctx->ev[4]);
        ctx->phase_ms[5] = splan.ok ? 1.0f : 0.0f;
    }
    // one Horner chain (~256 dependent doublings, ~0.1 ms) per MSM of the batch: independent, so the extra ones run on
    // their own host threads (L and R of an IPA round, the three vector commitments, the five T commitments)
    int rcs[MSM_MAX_BATCH] = {0};
    auto combine = [&](int m) {
        rcs[m] = host_combine(ctx->curve, (const xyzz*)ctx->h_result + (size_t)m * p.W, p.W, p.c, out_xy[m],
                              out_is_identity ? &out_is_identity[m] : nullptr);
    };
    host_parallel(ctx, job.nmsm, combine);
    for (int m = 0; m < job.nmsm; m++)
        if (rcs[m] != BP_OK) return rcs[m];
    return BP_OK;
}

// All-gather of `bytes` bytes per rank in a multi-GPU context: recv[r * bytes ...] = rank r's send. Library-owned when the
// context has an NCCL communicator (bp_ctx_init_nccl): ncclAllGather on the context's stream over NVLink, staged through
// one pinned buffer -- no callback, no Python, no torch tensor; otherwise the host program's callback (bp_ctx_set_collective).
inline int ctx_allgather(bp_ctx* ctx, const uint8_t* send, size_t bytes, uint8_t* recv) {
    if (ctx->world <= 1) { memcpy(recv, send, bytes); return BP_OK; }
    if (ctx->nccl_comm) {
        if (bytes * ctx->world > BP_HOST_COLL_BYTES || bytes > ctx->coll_send.cap) return BP_ERR_LEN;
        uint8_t* hs = (uint8_t*)ctx->h_coll;
        memcpy(hs, send, bytes);
        cudaStream_t st = ctx->stream;
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coll_send.p, hs, bytes, cudaMemcpyHostToDevice, st));
        const NcclApi& nc = nccl_api();
        ncclResult_t r = nc.AllGather(ctx->coll_send.p, ctx->coll_recv.p, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, st);
        if (r != ncclSuccess) { ctx->err = std::string("ncclAllGather: ") + nc.GetErrorString(r); return BP_ERR_CUDA; }
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(hs, ctx->coll_recv.p, bytes * ctx->world, cudaMemcpyDeviceToHost, st));
        BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        memcpy(recv, hs, bytes * ctx->world);
    } else {
        if (!ctx->coll) return BP_ERR_ARG;
        if (ctx->coll(ctx->coll_user, send, recv, bytes)) { ctx->err = "collective all-gather failed"; return BP_ERR_CUDA; }
    }
    ctx->coll_calls++;
    ctx->coll_bytes += bytes * ctx->world;
    return BP_OK;
}

// Multi-GPU contexts: the job holds this rank's shard of each of the `nmsm` MSMs. The partial points (64 B per MSM and
// rank, identity = zeros) are all-gathered and added on the host, so every rank returns the full sums. With world == 1
// this is msm_run_job.
int host_points_sum(int curve, const uint8_t* pts_xy, size_t n, uint8_t out_xy[64], int* out_is_identity);
template <class C>
int msm_run_job_sharded(bp_ctx* ctx, MsmJob& job, int nmsm, uint8_t (*out_xy)[64], int* out_is_identity) {
    if (job.nmsm < nmsm) job.nmsm = nmsm;           // a rank may hold no term of some MSM
    if (int rc = msm_run_job<C>(ctx, job, out_xy, out_is_identity)) return rc;
    if (ctx->world <= 1) return BP_OK;
    const size_t bytes = (size_t)nmsm * 64;
    std::vector<uint8_t> send(bytes), recv(bytes * ctx->world), pts((size_t)ctx->world * 64);
    for (int m = 0; m < nmsm; m++) {
        if (out_is_identity[m]) memset(&send[(size_t)m * 64], 0, 64);
        else memcpy(&send[(size_t)m * 64], out_xy[m], 64);
    }
    if (int rc = ctx_allgather(ctx, send.data(), bytes, recv.data())) return rc;
    for (int m = 0; m < nmsm; m++) {
        for (int r = 0; r < ctx->world; r++) memcpy(&pts[(size_t)r * 64], &recv[(size_t)r * bytes + (size_t)m * 64], 64);
        if (int rc = host_points_sum(ctx->curve, pts.data(), ctx->world, out_xy[m], &out_is_identity[m])) return rc;
    }
    return BP_OK;
}

template <class C>
int msm_run(bp_ctx* ctx, const affine* d_bases, const fe* d_scalars, size_t n, uint8_t out_xy[64], int* out_is_identity) {
    if (n > MSM_IDX_MASK) return BP_ERR_LEN;
    MsmJob job;
    if (n) job.add(d_bases, d_scalars, n, 0);
    uint8_t out[1][64];
    int ident[1] = {0};
    int rc = msm_run_job<C>(ctx, job, out, ident);
    if (rc != BP_OK) return rc;
    memcpy(out_xy, out[0], 64);
    if (out_is_identity) *out_is_identity = ident[0];
    return BP_OK;
}

// Host-resident MSM streamed through the GPU (bp_msm above ~6 M points): the input is cut into chunks, the H2D copies run
// back to back on the copy stream while the kernels of the chunks already there run, and **all chunks add into one bucket
// array** with the window width of the whole MSM, so the total work is that of the one-shot MSM (one bucket reduction,
// W * n bucket additions) -- independent per-chunk MSMs, the first implementation, use narrower windows and a reduction
// each: +29 % modmul at 2^24 points in 2M/2M/4M/8M chunks, which was exactly the gap between `e2e` and `value`.
// digits -> sort -> accumulate<ACC> -> slot levels<ACC> per chunk; reduce, window sums and the host Horner once.
template <class C>
int msm_run_streamed(bp_ctx* ctx, const uint8_t* h_bases, const uint8_t* h_scalars, size_t n, const std::vector<size_t>& lo_of,
                     const std::vector<size_t>& cnt_of, uint8_t out_xy[64], int* out_is_identity, const affine* d_bases = nullptr) {
    // d_bases != nullptr: the bases already live on the device (bp_bases, e.g. generators) and only the scalars stream
    cudaStream_t st = ctx->stream;
    if (n > MSM_IDX_MASK) return BP_ERR_LEN;
    const size_t nchunks = lo_of.size();
    size_t maxc = 0;
    for (size_t c : cnt_of) maxc = c > maxc ? c : maxc;
    MsmPlan p = make_plan(n, 1, ctx->force_c, ctx->sm_count);        // windows of the whole MSM
    const size_t max_entries = (maxc * (size_t)p.W + 63) & ~(size_t)63;   // 256-byte aligned halves of the double buffers
    // the whole input is staged (96 B per point: 1.6 GB at 2^24, 26 GB at the 2^28 limit of the pair format -- HBM has
    // 180 GB), so the copy stream runs back to back from the first byte to the last and never waits for a kernel
    if (!d_bases) BP_CUDA_TRY(ctx, ctx->stage_bases.reserve(n * 64));
    BP_CUDA_TRY(ctx, ctx->stage_scalars.reserve(n * 32));
    BP_CUDA_TRY(ctx, ctx->keys_a.reserve(max_entries * 4));
    BP_CUDA_TRY(ctx, ctx->keys_b.reserve(2 * max_entries * 4));     // sorted pairs: one buffer per chunk parity
    BP_CUDA_TRY(ctx, ctx->vals_a.reserve(max_entries * 4));
    BP_CUDA_TRY(ctx, ctx->vals_b.reserve(2 * max_entries * 4));
    const size_t nbuckets = (size_t)p.W * p.nb;
    BP_CUDA_TRY(ctx, ctx->buckets.reserve(nbuckets * sizeof(xyzz)));
    {
        size_t worst1 = 0;
        for (size_t c : cnt_of) {
            MsmPlan q = make_plan(c, 1, p.c, ctx->sm_count);
            worst1 = 2 * q.T > worst1 ? 2 * q.T : worst1;
        }
        size_t slots2 = 2 * ((worst1 + MSM_PL - 1) / MSM_PL);
        BP_CUDA_TRY(ctx, ctx->part_keys.reserve((worst1 + slots2 + 64) * 4));
        BP_CUDA_TRY(ctx, ctx->part_pts.reserve((worst1 + slots2 + 64) * sizeof(xyzz)));
    }
    BP_CUDA_TRY(ctx, ctx->seg_out.reserve((2 * (size_t)p.W * p.nseg + (size_t)p.W * MSM_PAIR_SLICES) * sizeof(xyzz)));
    BP_CUDA_TRY(ctx, ctx->win_out.reserve((size_t)p.W * sizeof(xyzz)));
    if ((size_t)p.W * sizeof(xyzz) > BP_HOST_RESULT_BYTES) return BP_ERR_LEN;
    size_t tmp_bytes = 0;
    BP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->keys_a.as<uint32_t>(), ctx->keys_b.as<uint32_t>(),
                                                     ctx->vals_a.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), max_entries, 0,
                                                     p.key_bits, st));
    BP_CUDA_TRY(ctx, ctx->cub_tmp.reserve(tmp_bytes));
    if (ctx->msm_sort_mode == 1) {
        size_t worst = 0;
        for (size_t c : cnt_of) {
            SortPlan sp = make_sort_plan(c, p.W, p.c - 1);
            if (sp.ok && sort_scratch_bytes(sp) > worst) worst = sort_scratch_bytes(sp);
        }
        if (worst) BP_CUDA_TRY(ctx, ctx->sort_scratch.reserve(worst));
    }
    ctx->last_c = p.c; ctx->last_W = p.W; ctx->last_entries = n * (size_t)p.W;
    if (ctx->timing) for (int i = 0; i < 5; i++) ctx->phase_ms[i] = 0;

    // Three streams besides the copy stream, all queued up front (the host never blocks):
    //   prep : digits + radix sort of chunk k (HBM-bound) -- runs next to the accumulate kernel of chunk k-1
    //          (IMAD-bound, DRAM ~10 % busy) and fills that kernel's tail; sorted pairs are double-buffered
    //   main : accumulate of chunk k, after its pairs are sorted and the slots of chunk k-1 are folded
    //   fold : the slot levels of chunk k -- a handful of latency-bound launches -- next to the prep of chunk k+1
    // prep and fold have the higher priority, so their blocks take the SM slots the long accumulate blocks free.
    // Measured and not kept (profiles/r2_msm_stream_timeline.txt): the timeline (BP_MSM_TIMELINE=1) shows accumulate k+1
    // starting exactly when the slot levels of chunk k are done, 0.2-1.3 ms after accumulate k (5.6 ms over nine chunks).
    // Letting the slot levels add into a second bucket array with double-buffered slot lists, so that they run next to
    // accumulate k+1, and copying the scalars of a chunk before its bases so that its sort starts earlier, closes every gap
    // -- and the accumulate kernels get slower by the same amount (206-register slot-level blocks take the place of 1.6
    // accumulate blocks each and last 3x longer themselves), plus 1.3 ms for merging the second array: 46.95 ms against
    // 45.1 ms; with the fold stream at the accumulate kernel's priority 48.9 ms. The scalars-first copy order alone (the
    // sort of chunk k+1 then runs next to accumulate k instead of next to the slot levels of chunk k): 46.3 ms -- whatever
    // runs next to the accumulate kernel costs it more than the gap it would have filled.
    // Sorted-pair buffer k&1 is free again once the accumulate kernel of chunk k has run.
    struct Scoped {
        cudaEvent_t used[2] = {nullptr, nullptr}, sorted[2] = {nullptr, nullptr};
        cudaEvent_t acc_done = nullptr, fold_done = nullptr;
        cudaStream_t fold = nullptr, prep = nullptr;
        std::vector<cudaEvent_t> copied;
        ~Scoped() {
            for (auto x : copied) if (x) cudaEventDestroy(x);
            for (auto x : used) if (x) cudaEventDestroy(x);
            for (auto x : sorted) if (x) cudaEventDestroy(x);
            if (acc_done) cudaEventDestroy(acc_done);
            if (fold_done) cudaEventDestroy(fold_done);
            if (fold) cudaStreamDestroy(fold);
            if (prep) cudaStreamDestroy(prep);
        }
    } sc;
    for (auto& x : sc.used) BP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    for (auto& x : sc.sorted) BP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    BP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sc.acc_done, cudaEventDisableTiming));
    BP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sc.fold_done, cudaEventDisableTiming));
    int prio_lo = 0, prio_hi = 0;
    BP_CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    BP_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&sc.fold, cudaStreamNonBlocking, prio_hi));
    BP_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&sc.prep, cudaStreamNonBlocking, prio_hi));
    sc.copied.assign(nchunks, nullptr);
    // BP_MSM_TIMELINE=1: per-chunk timestamps (copy done, pairs sorted, accumulate start / end, slots folded) on stderr
    const bool timeline = getenv("BP_MSM_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> tl;
    struct TlFree { std::vector<cudaEvent_t>& v; ~TlFree() { for (auto x : v) if (x) cudaEventDestroy(x); } } tl_free{tl};
    if (timeline) {
        tl.assign(1 + 5 * nchunks + 1, nullptr);
        for (auto& x : tl) BP_CUDA_TRY(ctx, cudaEventCreate(&x));
        BP_CUDA_TRY(ctx, cudaEventRecord(tl[0], ctx->copy_stream));
    }
    for (size_t k = 0; k < nchunks; k++) {
        BP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sc.copied[k], cudaEventDisableTiming));
        const size_t lo = lo_of[k], cnt = cnt_of[k];
        if (!d_bases)
            BP_CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)ctx->stage_bases.p + lo * 64, h_bases + lo * 64, cnt * 64, cudaMemcpyHostToDevice, ctx->copy_stream));
        BP_CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t*)ctx->stage_scalars.p + lo * 32, h_scalars + lo * 32, cnt * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
        BP_CUDA_TRY(ctx, cudaEventRecord(sc.copied[k], ctx->copy_stream));
        if (timeline) BP_CUDA_TRY(ctx, cudaEventRecord(tl[1 + 5 * k], ctx->copy_stream));
    }
    BP_CUDA_TRY(ctx, cudaMemsetAsync(ctx->buckets.p, 0, nbuckets * sizeof(xyzz), st));
    for (size_t k = 0; k < nchunks; k++) {
        const int s = (int)(k & 1);
        const size_t lo = lo_of[k], cnt = cnt_of[k];
        uint32_t* skeys = ctx->keys_b.as<uint32_t>() + (size_t)s * max_entries;
        uint32_t* svals = ctx->vals_b.as<uint32_t>() + (size_t)s * max_entries;
        MsmJob job;
        job.add((d_bases ? d_bases : ctx->stage_bases.as<affine>()) + lo, ctx->stage_scalars.as<fe>() + lo, cnt, 0);
        MsmPlan q = make_plan(cnt, 1, p.c, ctx->sm_count);   // same c, W, nb, key_bits; L and T of this chunk
        // prep
        BP_CUDA_TRY(ctx, cudaStreamWaitEvent(sc.prep, sc.copied[k], 0));
        if (k >= 2) BP_CUDA_TRY(ctx, cudaStreamWaitEvent(sc.prep, sc.used[s], 0));   // chunk k-2 has consumed its sorted pairs
        SortPlan splan;
        if (ctx->msm_sort_mode == 1 && q.entries >= ctx->msm_sort_min_entries) splan = make_sort_plan(cnt, q.W, q.c - 1);
        if (splan.ok && sort_scratch_bytes(splan) > ctx->sort_scratch.cap) splan.ok = false;
        if (splan.ok)
            msm_digits_kernel<C, true><<<(unsigned)((cnt + 255) / 256), 256, 0, sc.prep>>>(job, cnt, q.c, q.W, ctx->keys_a.as<uint32_t>(), nullptr);
        else
            msm_digits_kernel<C><<<(unsigned)((cnt + 255) / 256), 256, 0, sc.prep>>>(job, cnt, q.c, q.W, ctx->keys_a.as<uint32_t>(),
                                                                                    ctx->vals_a.as<uint32_t>());
        BP_LAUNCH_CHECK(ctx);
        if (splan.ok) {
            int nl = 0;
            BP_CUDA_TRY(ctx, sort_pairs_run(splan, sort_segs_of(job), ctx->keys_a.as<uint32_t>(), skeys, svals, ctx->sort_scratch.p, sc.prep, &nl));
            ctx->launches += nl;
        } else {
            BP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, ctx->keys_a.as<uint32_t>(), skeys,
                                                             ctx->vals_a.as<uint32_t>(), svals, q.entries, 0, q.key_bits, sc.prep));
        }
        BP_CUDA_TRY(ctx, cudaEventRecord(sc.sorted[s], sc.prep));
        if (timeline) BP_CUDA_TRY(ctx, cudaEventRecord(tl[2 + 5 * k], sc.prep));
        // main
        BP_CUDA_TRY(ctx, cudaStreamWaitEvent(st, sc.sorted[s], 0));
        if (k >= 1) BP_CUDA_TRY(ctx, cudaStreamWaitEvent(st, sc.fold_done, 0));      // slots of chunk k-1 folded into the buckets
        if (timeline) BP_CUDA_TRY(ctx, cudaEventRecord(tl[3 + 5 * k], st));
        msm_accumulate_kernel<C, true><<<(unsigned)((q.T + 127) / 128), 128, 0, st>>>(skeys, svals, q.entries, q.L, q.T, job,
                                                                                      ctx->buckets.as<xyzz>(), ctx->part_keys.as<uint32_t>(),
                                                                                      ctx->part_pts.as<xyzz>());
        BP_LAUNCH_CHECK(ctx);
        BP_CUDA_TRY(ctx, cudaEventRecord(sc.used[s], st));
        BP_CUDA_TRY(ctx, cudaEventRecord(sc.acc_done, st));
        if (timeline) BP_CUDA_TRY(ctx, cudaEventRecord(tl[4 + 5 * k], st));
        // fold
        BP_CUDA_TRY(ctx, cudaStreamWaitEvent(sc.fold, sc.acc_done, 0));
        if (int rc = msm_fold_slots<C, true>(ctx, 2 * q.T, sc.fold)) return rc;
        BP_CUDA_TRY(ctx, cudaEventRecord(sc.fold_done, sc.fold));
        if (timeline) BP_CUDA_TRY(ctx, cudaEventRecord(tl[5 + 5 * k], sc.fold));
    }
    BP_CUDA_TRY(ctx, cudaStreamWaitEvent(st, sc.fold_done, 0));
    if (int rc = msm_reduce_windows<C>(ctx, p, p.W, st)) return rc;
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_result, ctx->win_out.p, (size_t)p.W * sizeof(xyzz), cudaMemcpyDeviceToHost, st));
    if (timeline) BP_CUDA_TRY(ctx, cudaEventRecord(tl[1 + 5 * nchunks], st));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (timeline) {
        auto at = [&](size_t i) { float ms = 0; cudaEventElapsedTime(&ms, tl[0], tl[i]); return ms; };
        for (size_t k = 0; k < nchunks; k++)
            fprintf(stderr, "[bp_msm timeline] chunk %zu: %8zu points  copied %7.3f  sorted %7.3f  accumulate %7.3f .. %7.3f  folded %7.3f ms\n", k,
                    cnt_of[k], at(1 + 5 * k), at(2 + 5 * k), at(3 + 5 * k), at(4 + 5 * k), at(5 + 5 * k));
        fprintf(stderr, "[bp_msm timeline] window sums on the host at %7.3f ms\n", at(1 + 5 * nchunks));
    }
    return host_combine(ctx->curve, (const xyzz*)ctx->h_result, p.W, p.c, out_xy, out_is_identity);
}

template <class C>
int synth_points_run(bp_ctx* ctx, void* d_out, size_t n, uint64_t start) {
    if (n == 0) return BP_OK;
    synth_points_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((affine*)d_out, n, start);
    BP_LAUNCH_CHECK(ctx);
    return BP_OK;
}

// The kernels are instantiated once per curve, in msm_<curve>.cu (which define BP_MSM_INSTANTIATE); every other
// translation unit that includes this header (the prover / verifier layer) links against those instances.
#if !defined(BP_MSM_INSTANTIATE)
#define BP_MSM_EXTERN(C)                                                                          \
    extern template int msm_run<C>(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);    \
    extern template int msm_run_job<C>(bp_ctx*, const MsmJob&, uint8_t (*)[64], int*);
BP_MSM_EXTERN(Secq256k1)
BP_MSM_EXTERN(Zorro)
BP_MSM_EXTERN(Curve25519)
#undef BP_MSM_EXTERN
#endif

}  // namespace bp

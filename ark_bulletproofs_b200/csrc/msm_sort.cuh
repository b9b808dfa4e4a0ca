// Bucket sort of the (bucket key, point index | sign) pairs of one large MSM -- step 2 of msm_kernels.cuh without a library
// radix sort (cub::DeviceRadixSort was 4.6 ms of the 2^24-point MSM at 0.35 of the HBM roofline; an LSD radix
// sort moves every pair three times and ranks it stably each time).
//
// What the accumulate kernel needs is weaker than a sort: equal keys adjacent, keys ascending, zero digits
// (INVALID_KEY) at the end. The order inside a bucket is irrelevant (the bucket sum is commutative), so ranking can use
// shared-memory atomics instead of a stable warp-match ranking, and two MSD passes suffice:
//   hist    : per tile of TS pairs of one window, counts per coarse bin (the top bits of the bucket id)        [reads keys]
//   colscan : per (window, bin) column, exclusive running sum over the tiles; column totals
//   binscan : exclusive scan of the column totals -> where every coarse bin starts in the output; zero digits last
//   scatter : the tile is ordered by coarse bin in shared memory (counting sort, shared atomics) and every (tile, bin) run
//             is written to its place: bin start + offset of the tile; the value of a pair (sign, segment, index) is
//             rebuilt from its position -- the digits kernel writes keys only, sign in bit 31          [reads keys, 1 write]
//   bins    : one block per coarse bin (8-16 K pairs): ranks by the remaining low bits with shared atomics, places the
//             pairs in shared memory in their final order and writes them out linearly                         [1 read, 1 write]
// Keys are read twice (4 B), pairs written once and moved once (8 B each way): 32 B of HBM traffic per pair against ~52 B for three
// onesweep passes plus their histogram. Any key distribution is handled: a coarse bin that does not fit the shared-memory
// stage (adversarial scalar sets: all scalars equal puts a whole window into one bucket) is ranked from global memory by
// the same block, slower but correct.
#pragma once
#include <cstddef>
#include <cstdint>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define BP_SORT_HD __host__ __device__
#else
#define BP_SORT_HD
#endif

namespace bp {

static constexpr uint32_t SORT_INVALID_KEY = 0xFFFFFFFFu;
static constexpr int SORT_TS = 8192;           // pairs per tile (hist / scatter): 64 KB of stage + 16 KB of index, two blocks per SM
static constexpr int SORT_BIN_TARGET = 8192;   // coarse bins hold between this many and twice as many pairs on average
static constexpr int SORT_BIN_CAP = 20480;     // pairs staged in shared memory by the bins kernel (160 KB + 40 KB of index)

struct SortPlan {
    bool ok = false;
    uint32_t n = 0;         // pairs per window (window w is pairs [w*n, (w+1)*n) of the input)
    uint32_t W = 0;         // windows
    int cb = 0;             // bucket bits per window (key = w << cb | bucket)
    int low_bits = 0;       // bits ranked by the bins kernel
    int low_top = 0;        // the same for the last window, whose digits only span 2^top_bits values (256 is no multiple of c):
                            // its coarse bins split that range, not the full bucket range; 0 = one bucket per coarse bin, and
                            // the scatter kernel writes the final arrays itself
    uint32_t nb1 = 0;       // coarse bins per window = 2^(cb - low_bits)
    uint32_t tiles = 0;     // tiles per window
    size_t hist_elems() const { return (size_t)W * tiles * (nb1 + 1); }
    size_t bins_total() const { return (size_t)W * nb1; }
    BP_SORT_HD int low_of(uint32_t w) const { return w + 1 == W ? low_top : low_bits; }
};

// segments of the MSM job (msm_kernels.cuh: MsmJob): term i of a window belongs to segment sg with start[sg] <= i < start[sg + 1];
// its value is sign << 31 | sg << 28 | (i - start[sg])
struct SortSegs {
    int nseg = 0;
    uint32_t start[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// One MSM (nmsm == 1) of n terms, W windows of cb bucket bits. Returns ok = false where the library sort stays.
static inline SortPlan make_sort_plan(size_t n, int W, int cb, int scalar_bits = 256) {
    SortPlan s;
    if (n < (size_t)SORT_BIN_TARGET * 4 || n > 0x7FFFFFFFu || cb < 3) return s;
    int lb1 = 0;                                   // log2 of the coarse bins per window
    while (((size_t)SORT_BIN_TARGET << (lb1 + 1)) <= n && lb1 < cb) lb1++;
    int low = cb - lb1;
    if (low > 10) { lb1 += low - 10; low = 10; }   // at most 1024 low-bit counters per block
    if (lb1 > 10) { low += lb1 - 10; lb1 = 10; }   // <= 1024 coarse bins per window (one counter per thread of the tile kernels)
    if (low > 10) return s;
    if ((n >> lb1) > (size_t)SORT_BIN_CAP * 4 / 5) return s;   // bins would not fit the stage: library sort
    s.n = (uint32_t)n; s.W = (uint32_t)W; s.cb = cb; s.low_bits = low; s.nb1 = 1u << lb1;
    s.tiles = (uint32_t)((n + SORT_TS - 1) / SORT_TS);
    // last window: digit magnitudes <= 2^top_bits (bits of the scalar left for it, plus the carry), buckets < 2^top_bits
    int top_bits = scalar_bits - (W - 1) * (cb + 1);
    if (top_bits < 0) top_bits = 0;
    if (top_bits > cb) top_bits = cb;
    s.low_top = top_bits - lb1 < 0 ? 0 : top_bits - lb1;
    if (s.low_top > low) s.low_top = low;
    s.ok = true;
    return s;
}

#if defined(__CUDACC__)
// ---- hist: tile_hist[w][tile][b] for b in [0, nb1] (b = nb1 counts the zero digits) ------------------------------------
static __global__ void __launch_bounds__(512) sort_hist_kernel(const uint32_t* __restrict__ keys, const __grid_constant__ SortPlan sp,
                                                        uint16_t* __restrict__ tile_hist) {
    extern __shared__ uint32_t sh_cnt[];
    const uint32_t w = blockIdx.y, tile = blockIdx.x, nbp = sp.nb1 + 1;
    for (uint32_t b = threadIdx.x; b < nbp; b += blockDim.x) sh_cnt[b] = 0;
    __syncthreads();
    const uint32_t lo = tile * SORT_TS, hi = lo + SORT_TS < sp.n ? lo + SORT_TS : sp.n;
    const uint32_t* K = keys + (size_t)w * sp.n;
    const uint32_t mask = sp.nb1 - 1;
    const int low = sp.low_of(w);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t k = __ldg(K + i);                 // bit 31 = sign of the digit; the mask drops it
        const uint32_t b = k == SORT_INVALID_KEY ? sp.nb1 : ((k >> low) & mask);
        atomicAdd(&sh_cnt[b], 1u);
    }
    __syncthreads();
    uint16_t* out = tile_hist + ((size_t)w * sp.tiles + tile) * nbp;
    for (uint32_t b = threadIdx.x; b < nbp; b += blockDim.x) out[b] = (uint16_t)sh_cnt[b];   // <= SORT_TS < 65536
}

// ---- colscan: rel[w][tile][b] = pairs of column (w, b) in earlier tiles; col_total[w][b] -------------------------------
// One thread per (column, group of tiles): 32 columns x 32 groups per block. A thread sums the counts of its tiles, the
// groups of a column are combined through shared memory, and a second sweep over the same tiles (L2 hits) writes the
// offsets -- a serial walk over all 2048 tiles of a column per thread was 0.15-0.2 ms of pure load latency at 2^24 points.
static __global__ void __launch_bounds__(1024) sort_colscan_kernel(const uint16_t* __restrict__ tile_hist, const __grid_constant__ SortPlan sp,
                                                           uint32_t* __restrict__ rel, uint32_t* __restrict__ col_total) {
    __shared__ uint32_t sh_sum[32][33];
    const uint32_t nbp = sp.nb1 + 1;
    const uint32_t c = threadIdx.x & 31u, g = threadIdx.x >> 5;
    const size_t col = (size_t)blockIdx.x * 32 + c;
    const bool live = col < (size_t)sp.W * nbp;
    const uint32_t w = live ? (uint32_t)(col / nbp) : 0u, b = live ? (uint32_t)(col % nbp) : 0u;
    const uint32_t tpg = (sp.tiles + 31) / 32;
    const uint32_t t0 = g * tpg < sp.tiles ? g * tpg : sp.tiles, t1 = t0 + tpg < sp.tiles ? t0 + tpg : sp.tiles;
    const uint16_t* h = tile_hist + (size_t)w * sp.tiles * nbp + b;
    uint32_t* r = rel + (size_t)w * sp.tiles * nbp + b;
    uint32_t sum = 0;
    if (live) {
        uint32_t t = t0;
        for (; t + 8 <= t1; t += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = h[(size_t)(t + j) * nbp];
#pragma unroll
            for (int j = 0; j < 8; j++) sum += v[j];
        }
        for (; t < t1; t++) sum += h[(size_t)t * nbp];
    }
    sh_sum[g][c] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (uint32_t gg = 0; gg < g; gg++) run += sh_sum[gg][c];
    if (live) {
        if (g == 31) col_total[col] = run + sum;
        uint32_t t = t0;
        for (; t + 8 <= t1; t += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = h[(size_t)(t + j) * nbp];
#pragma unroll
            for (int j = 0; j < 8; j++) { r[(size_t)(t + j) * nbp] = run; run += v[j]; }
        }
        for (; t < t1; t++) { const uint32_t v = h[(size_t)t * nbp]; r[(size_t)t * nbp] = run; run += v; }
    }
}

// ---- binscan: bin_start[w * nb1 + b] (valid bins in key order), then the zero-digit regions of the windows; one block ----
// bin_start has W * nb1 + W + 1 entries: valid bins, per-window invalid regions, and the end (= W * n).
static __global__ void __launch_bounds__(1024) sort_binscan_kernel(const uint32_t* __restrict__ col_total, const __grid_constant__ SortPlan sp,
                                                            uint32_t* __restrict__ bin_start) {
    __shared__ uint32_t sh_part[1024];
    __shared__ uint32_t sh_base;
    const uint32_t nbp = sp.nb1 + 1;
    const uint32_t nvalid = sp.W * sp.nb1, total = nvalid + sp.W;
    // element e < nvalid: column (e / nb1, e % nb1); e >= nvalid: the invalid column of window e - nvalid
    auto value = [&](uint32_t e) -> uint32_t {
        if (e < nvalid) return col_total[(size_t)(e / sp.nb1) * nbp + (e % sp.nb1)];
        return col_total[(size_t)(e - nvalid) * nbp + sp.nb1];
    };
    if (threadIdx.x == 0) sh_base = 0;
    __syncthreads();
    for (uint32_t chunk = 0; chunk < total; chunk += 1024 * 8) {
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t e = chunk + threadIdx.x * 8 + j;
            v[j] = e < total ? value(e) : 0u;
            sum += v[j];
        }
        sh_part[threadIdx.x] = sum;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {        // Hillis-Steele inclusive scan of the thread sums
            uint32_t add = threadIdx.x >= (unsigned)d ? sh_part[threadIdx.x - d] : 0u;
            __syncthreads();
            sh_part[threadIdx.x] += add;
            __syncthreads();
        }
        uint32_t run = sh_base + sh_part[threadIdx.x] - sum;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t e = chunk + threadIdx.x * 8 + j;
            if (e < total) bin_start[e] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) sh_base += sh_part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_start[total] = sh_base;
}

// exclusive scan of one value per thread over a 1024-thread block (sh_warp: 32 words); returns the exclusive prefix
__device__ __forceinline__ uint32_t sort_block_excl_scan_1024(uint32_t v, uint32_t* sh_warp) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (unsigned)d) inc += up;
    }
    if (lane == 31) sh_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = sh_warp[lane], ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, ti, d);
            if (lane >= (unsigned)d) ti += up;
        }
        sh_warp[lane] = ti - t;
    }
    __syncthreads();
    return inc - v + sh_warp[warp];
}

// ---- scatter: pairs to their coarse bins; zero digits straight to the tail of the final key array ----------------------
// A direct scatter (one 8-byte store per pair into <= 1024 moving cursors) leaves every in-flight tile with thousands of
// partially written sectors: measured at 2^24 points, 3.3x write and read amplification in DRAM (L2 evicts them before they
// fill; ncu: 5.7 GB written for 1.7 GB of pairs, 5.9 ms). So the tile is ordered by coarse bin inside shared memory first
// (counting sort with an inverse index: stage[] keeps the pairs, inv[] says which pair comes i-th) and written out
// linearly: every (tile, bin) run of ~8 pairs is one contiguous 64-byte store stream, completed to whole sectors in L2 by
// the neighbouring tiles' runs (ncu: DRAM bytes written = bytes of pairs). Tiles of 8192 pairs keep the stage at 92 KB, so
// two blocks share an SM and one block's loads overlap the other's stores (1.18 -> 1.01 ms at 2^24 against 16384-pair tiles).
static constexpr int SORT_TILE_SMEM = SORT_TS * 8 + SORT_TS * 2 + 3 * 1028 * 4 + 32 * 4;
static __global__ void __launch_bounds__(1024) sort_scatter_kernel(const uint32_t* __restrict__ keys, const __grid_constant__ SortSegs segs,
                                                           const __grid_constant__ SortPlan sp, const uint32_t* __restrict__ rel,
                                                           const uint32_t* __restrict__ bin_start, uint2* __restrict__ pairs,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    extern __shared__ __align__(16) uint8_t sh_raw[];
    uint2* stage = reinterpret_cast<uint2*>(sh_raw);
    uint16_t* inv = reinterpret_cast<uint16_t*>(sh_raw + (size_t)SORT_TS * 8);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(sh_raw + (size_t)SORT_TS * 10);      // counters, then cursors
    uint32_t* off = cnt + 1028;
    uint32_t* gbase = off + 1028;
    uint32_t* sh_warp = gbase + 1028;
    const uint32_t w = blockIdx.y, tile = blockIdx.x, nbp = sp.nb1 + 1;               // nbp <= 1025
    const uint32_t* r = rel + ((size_t)w * sp.tiles + tile) * nbp;
    const uint32_t nvalid = sp.W * sp.nb1;
    for (uint32_t b = threadIdx.x; b < nbp; b += blockDim.x) {
        cnt[b] = 0;
        gbase[b] = r[b] + (b < sp.nb1 ? bin_start[(size_t)w * sp.nb1 + b] : bin_start[nvalid + w]);
    }
    __syncthreads();
    const uint32_t lo = tile * SORT_TS, hi = lo + SORT_TS < sp.n ? lo + SORT_TS : sp.n, m = hi - lo;
    const uint32_t* K = keys + (size_t)w * sp.n + lo;
    const uint32_t mask = sp.nb1 - 1;
    const int low = sp.low_of(w);
    auto bin_of = [&](uint32_t k) -> uint32_t { return k == SORT_INVALID_KEY ? sp.nb1 : ((k >> low) & mask); };
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < m; i += 1024) {
        // keys carry the sign in bit 31 (msm_digits_kernel<KEYS_ONLY>); the value is rebuilt from the term index
        const uint32_t kr = __ldg(K + i);
        const uint32_t gi = lo + i;
        int sg = 0;
#pragma unroll
        for (int q = 1; q < 8; q++)
            if (q < segs.nseg && gi >= segs.start[q]) sg = q;
        const uint32_t v = (kr & 0x80000000u) | ((uint32_t)sg << 28) | (gi - segs.start[sg]);
        const uint32_t k = kr == SORT_INVALID_KEY ? kr : (kr & 0x7FFFFFFFu);
        stage[i] = make_uint2(k, v);
        atomicAdd(&cnt[bin_of(k)], 1u);
    }
    __syncthreads();
    {
        const uint32_t c0 = threadIdx.x < nbp ? cnt[threadIdx.x] : 0u;
        const uint32_t ex = sort_block_excl_scan_1024(c0, sh_warp);
        if (threadIdx.x < nbp) { off[threadIdx.x] = ex; cnt[threadIdx.x] = ex; }
        if (nbp == 1025 && threadIdx.x == 1023) { off[1024] = ex + c0; cnt[1024] = ex + c0; }   // the zero-digit column comes last
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += 1024) {
        const uint32_t pos = atomicAdd(&cnt[bin_of(stage[i].x)], 1u);
        inv[pos] = (uint16_t)i;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += 1024) {
        const uint2 p = stage[inv[i]];
        const uint32_t b = bin_of(p.x);
        const uint32_t g = gbase[b] + (i - off[b]);
        if (b == sp.nb1) keys_out[g] = SORT_INVALID_KEY;
        else if (low == 0) { keys_out[g] = p.x; vals_out[g] = p.y; }      // one bucket per coarse bin: already final
        else pairs[g] = p;
    }
}

// ---- bins: final order inside every coarse bin ------------------------------------------------------------------------
// The same block-local counting sort, keyed by the low bits; bins beyond the shared-memory stage take the global path.
static constexpr int SORT_BINS_SMEM = SORT_BIN_CAP * 8 + SORT_BIN_CAP * 2 + 2 * 1024 * 4 + 32 * 4;
static __global__ void __launch_bounds__(1024) sort_bins_kernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ bin_start,
                                                                     const __grid_constant__ SortPlan sp, uint32_t* __restrict__ keys_out,
                                                                     uint32_t* __restrict__ vals_out) {
    extern __shared__ __align__(16) uint8_t sh_raw[];
    uint2* stage = reinterpret_cast<uint2*>(sh_raw);
    uint16_t* inv = reinterpret_cast<uint16_t*>(sh_raw + (size_t)SORT_BIN_CAP * 8);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(sh_raw + (size_t)SORT_BIN_CAP * 10);
    uint32_t* off = cnt + 1024;
    uint32_t* sh_warp = off + 1024;
    const uint32_t bin = blockIdx.x;
    const int low = sp.low_of(bin / sp.nb1);
    if (low == 0) return;                           // written by the scatter kernel
    const uint32_t start = bin_start[bin], m = bin_start[bin + 1] - start;
    if (m == 0) return;
    const uint32_t nlow = 1u << low, lmask = nlow - 1;      // nlow <= 1024
    if (threadIdx.x < nlow) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint2* P = pairs + start;
    const bool staged = m <= (uint32_t)SORT_BIN_CAP;
    if (staged) {
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < m; i += 1024) {
            const uint2 p = P[i];
            stage[i] = p;
            atomicAdd(&cnt[p.x & lmask], 1u);
        }
    } else {
        // oversize bin (skewed keys): counted from global memory, one shared atomic per distinct key of a warp
        for (uint32_t i0 = 0; i0 < m; i0 += 1024) {
            const uint32_t i = i0 + threadIdx.x;
            const bool act = i < m;
            const uint32_t kk = act ? (P[i].x & lmask) : (0xFFFFFF00u | (threadIdx.x & 31u));
            const unsigned grp = __match_any_sync(0xFFFFFFFFu, kk);
            if (act && (threadIdx.x & 31u) == (unsigned)(__ffs(grp) - 1)) atomicAdd(&cnt[kk], (uint32_t)__popc(grp));
        }
    }
    __syncthreads();
    {
        const uint32_t c0 = threadIdx.x < nlow ? cnt[threadIdx.x] : 0u;
        const uint32_t ex = sort_block_excl_scan_1024(c0, sh_warp);
        if (threadIdx.x < nlow) { off[threadIdx.x] = ex; cnt[threadIdx.x] = ex; }
    }
    __syncthreads();
    if (staged) {
        for (uint32_t i = threadIdx.x; i < m; i += 1024) {
            const uint32_t pos = atomicAdd(&cnt[stage[i].x & lmask], 1u);
            inv[pos] = (uint16_t)i;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < m; i += 1024) {
            const uint2 p = stage[inv[i]];
            keys_out[start + i] = p.x;
            vals_out[start + i] = p.y;
        }
    } else {
        for (uint32_t i0 = 0; i0 < m; i0 += 1024) {
            const uint32_t i = i0 + threadIdx.x;
            const bool act = i < m;
            uint2 p = make_uint2(0u, 0u);
            if (act) p = P[i];
            const uint32_t kk = act ? (p.x & lmask) : (0xFFFFFF00u | (threadIdx.x & 31u));
            const unsigned grp = __match_any_sync(0xFFFFFFFFu, kk);
            const unsigned lane = threadIdx.x & 31u, leader = (unsigned)(__ffs(grp) - 1);
            uint32_t base = 0;
            if (act && lane == leader) base = atomicAdd(&cnt[kk], (uint32_t)__popc(grp));
            base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
            if (act) {
                const uint32_t pos = start + base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
                keys_out[pos] = p.x;
                vals_out[pos] = p.y;
            }
        }
    }
}

// scratch layout (one allocation): tile_hist (uint16) | rel | col_total | bin_start | pairs
struct SortScratch {
    uint16_t* tile_hist;
    uint32_t *rel, *col_total, *bin_start;
    uint2* pairs;
};
static inline size_t sort_scratch_bytes(const SortPlan& sp) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return al(sp.hist_elems() * 2) + al(sp.hist_elems() * 4) + al((size_t)sp.W * (sp.nb1 + 1) * 4) + al((sp.bins_total() + sp.W + 2) * 4) +
           al((size_t)sp.W * sp.n * 8);
}
static inline SortScratch sort_scratch_at(void* base, const SortPlan& sp) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    uint8_t* p = (uint8_t*)base;
    SortScratch s;
    s.tile_hist = (uint16_t*)p; p += al(sp.hist_elems() * 2);
    s.rel = (uint32_t*)p; p += al(sp.hist_elems() * 4);
    s.col_total = (uint32_t*)p; p += al((size_t)sp.W * (sp.nb1 + 1) * 4);
    s.bin_start = (uint32_t*)p; p += al((sp.bins_total() + sp.W + 2) * 4);
    s.pairs = (uint2*)p;
    return s;
}

// keys_in: window-major keys (W * n) with the digit's sign in bit 31; keys_out / vals_out: sorted pairs, zero digits last.
static inline cudaError_t sort_pairs_run(const SortPlan& sp, const SortSegs& segs, const uint32_t* keys_in, uint32_t* keys_out, uint32_t* vals_out,
                                         void* scratch, cudaStream_t st, int* launches) {
    {   // per device; cheap driver calls
        cudaError_t e = cudaFuncSetAttribute(sort_bins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_BINS_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_TILE_SMEM);
        if (e != cudaSuccess) return e;
    }
    const SortScratch s = sort_scratch_at(scratch, sp);
    const uint32_t nbp = sp.nb1 + 1;
    sort_hist_kernel<<<dim3(sp.tiles, sp.W), 512, nbp * 4, st>>>(keys_in, sp, s.tile_hist);
    const size_t cols = (size_t)sp.W * nbp;
    sort_colscan_kernel<<<(unsigned)((cols + 31) / 32), 1024, 0, st>>>(s.tile_hist, sp, s.rel, s.col_total);
    sort_binscan_kernel<<<1, 1024, 0, st>>>(s.col_total, sp, s.bin_start);
    sort_scatter_kernel<<<dim3(sp.tiles, sp.W), 1024, SORT_TILE_SMEM, st>>>(keys_in, segs, sp, s.rel, s.bin_start, s.pairs, keys_out, vals_out);
    sort_bins_kernel<<<(unsigned)sp.bins_total(), 1024, SORT_BINS_SMEM, st>>>(s.pairs, s.bin_start, sp, keys_out, vals_out);
    *launches += 5;
    return cudaGetLastError();
}

#endif  // __CUDACC__

}  // namespace bp

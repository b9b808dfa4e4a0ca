// BulletproofGens generation on the device.
//
// Replaces the serial loop of GeneratorsChain (src/generators.rs:71-121: `G::rand(&mut ChaChaRng)` until `capacity`
// points exist) for curves whose `Affine::rand` attempt consumes a fixed number of ChaCha20 words. On secq256k1 the
// base field is within 2^129 of 2^256, so the x draw (4 x next_u64) is never rejected and attempt j reads exactly
// words [9j, 9j+9) of the keystream: 8 words of x (raw limbs = the Montgomery representation, ark-ff Fp::rand) and
// one word whose top bit is `greatest` (ark-ec get_point_from_x_unchecked). The keystream is seekable, so every
// attempt is independent: one thread per attempt computes its ChaCha block(s), y = sqrt(x^3 + a*x + b) by
// Tonelli-Shanks (2-adicity 6 for secq256k1's base field), and picks the root `greatest` asks for; the accepted
// points are compacted in stream order (cub::DeviceSelect) and the first `capacity` kept. A raw x >= q (probability
// 2^-128) would shift the stream: it raises a flag and the caller falls back to the host path.
// The host implementation (host/gens_host.hpp) stays for the other curves (zorro's x draw rejects with p ~ 1/2,
// curve25519 clears the cofactor) and is what the device result is tested against.
#pragma once
#include <cub/device/device_select.cuh>
#include "ctx.cuh"
#include "host/fp_host.hpp"
#include "host/merlin.hpp"

namespace bp {

struct ChaChaKey { uint32_t k[8]; };
struct SqrtParams {
    uint32_t t[8];      // q - 1 = 2^s * t
    uint32_t t1h[8];    // (t + 1) / 2
    fe z;               // z^t for a quadratic non-residue z
    int s;
};

__device__ __forceinline__ uint32_t gk_rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
// rand_chacha 0.3 ChaCha20Rng block: 64-bit counter in words 12-13, stream id 0 (same as host/merlin.hpp)
__device__ inline void chacha20_block_dev(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = s[i];
#define BP_GQR(a, b, c, d)                                                                                        \
    w[a] += w[b]; w[d] = gk_rotl(w[d] ^ w[a], 16); w[c] += w[d]; w[b] = gk_rotl(w[b] ^ w[c], 12);                 \
    w[a] += w[b]; w[d] = gk_rotl(w[d] ^ w[a], 8);  w[c] += w[d]; w[b] = gk_rotl(w[b] ^ w[c], 7);
#pragma unroll 1
    for (int i = 0; i < 10; i++) {
        BP_GQR(0, 4, 8, 12) BP_GQR(1, 5, 9, 13) BP_GQR(2, 6, 10, 14) BP_GQR(3, 7, 11, 15)
        BP_GQR(0, 5, 10, 15) BP_GQR(1, 6, 11, 12) BP_GQR(2, 7, 8, 13) BP_GQR(3, 4, 9, 14)
    }
#undef BP_GQR
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = w[i] + s[i];
}

// Tonelli-Shanks without the Euler pre-check: a non-residue shows up as "no i < m with t^(2^i) = 1"
template <class F>
__device__ bool fq_sqrt_dev(const fe& a, const SqrtParams& sp, fe& out) {
    if (F::is_zero(a)) { out = a; return true; }
    const fe one = F::one();
    int m = sp.s;
    fe c = sp.z, t = F::pow(a, sp.t), r = F::pow(a, sp.t1h);
    while (!F::eq(t, one)) {
        int i = 0;
        fe t2 = t;
        while (!F::eq(t2, one)) {
            t2 = F::sqr(t2);
            if (++i == m) return false;
        }
        fe b = c;
        for (int k = 0; k < m - i - 1; k++) b = F::sqr(b);
        m = i;
        c = F::sqr(b);
        t = F::mul(t, c);
        r = F::mul(r, b);
    }
    out = r;
    return true;
}

// One `Affine::rand` attempt per thread. The attempt's nine keystream words start at word offsets[j] (curves whose x
// draw rejects: the host parses the stream, see gens_chain_device) or at 9 * (attempt0 + j) (fixed stride).
//   short Weierstrass: x = the 8 words, y = sqrt(x^3 + a x + b), root chosen by `greatest`;
//   twisted Edwards:   y = the 8 words with the top bit shaved (255-bit field), x = sqrt((1 - y^2) / (a - d y^2)) chosen by
//                      `greatest`, then mul_by_cofactor (three doublings) -- ark-ec's Affine::rand for both models.
template <class C>
__global__ void __launch_bounds__(128) gens_attempt_kernel(const __grid_constant__ ChaChaKey key, uint64_t attempt0, size_t count,
                                                           const uint32_t* __restrict__ offsets, const __grid_constant__ SqrtParams sp,
                                                           affine* __restrict__ pts, uint8_t* __restrict__ ok, int* __restrict__ irregular) {
    using F = Fp<typename C::Fq>;
    using E = GroupLaw<C>;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const uint64_t w0 = offsets ? (uint64_t)offsets[j] : (attempt0 + j) * 9u;
    const uint64_t blk = w0 / 16u;
    const int off = (int)(w0 % 16u);
    uint32_t buf[32];
    chacha20_block_dev(key.k, blk, buf);
    if (off + 9 > 16) chacha20_block_dev(key.k, blk + 1, buf + 16);
    fe x;
#pragma unroll
    for (int k = 0; k < 8; k++) x.v[k] = buf[off + k];        // 4 x next_u64 = 8 little-endian words; raw = Montgomery repr
    if (C::Fq::BITS < 256) x.v[7] &= 0xFFFFFFFFu >> (256 - C::Fq::BITS);   // Fp::rand shaves the bits above the modulus' length
    const bool greatest = (buf[off + 8] >> 31) & 1u;           // bool::rand = top bit of next_u32
    // a raw value >= q makes Fp::rand redraw and shifts the stream: impossible here for pre-parsed offsets, astronomically
    // unlikely for the fixed-stride curves (the caller falls back to the host generator)
    bool geq = true;
    for (int k = 7; k >= 0; k--) {
        uint32_t mk = C::Fq::m(k);
        if (x.v[k] != mk) { geq = x.v[k] > mk; break; }
    }
    if (geq) { atomicExch(irregular, 1); ok[j] = 0; return; }
    auto larger = [](const fe& v) {                            // v > -v as canonical integers
        fe nv = F::neg(v);
        fe vc = F::from_mont(v), nc = F::from_mont(nv);
        for (int k = 7; k >= 0; k--)
            if (vc.v[k] != nc.v[k]) return vc.v[k] > nc.v[k];
        return false;
    };
    affine p;
    if (C::KIND == 1) {
        const fe y = x;
        const fe y2 = F::sqr(y);
        const fe num = F::sub(F::one(), y2);
        const fe den = F::sub(F::neg(F::one()), F::mul(F::template curve_b<C>(), y2));
        if (F::is_zero(den)) { ok[j] = 0; return; }
        fe xr;
        if (!fq_sqrt_dev<F>(F::mul(num, F::inv(den)), sp, xr)) { ok[j] = 0; return; }
        p.x = (greatest == larger(xr)) ? xr : F::neg(xr);
        p.y = y;
        xyzz acc = E::from_affine(p);
        for (int k = 0; k < 3; k++) acc = E::dbl(acc);         // cofactor 8
        p = E::to_affine(acc);
        if (E::is_identity(p)) { p.x = F::zero(); p.y = F::one(); }
    } else {
        fe rhs = F::add(F::mul(F::sqr(x), x), F::template curve_b<C>());
        if (C::A_SMALL != 0) rhs = F::add(rhs, F::mul_small(x, C::A_SMALL));
        fe y;
        if (!fq_sqrt_dev<F>(rhs, sp, y)) { ok[j] = 0; return; }
        p.x = x;
        p.y = (greatest == larger(y)) ? y : F::neg(y);
    }
    st_fe(&pts[j].x, p.x);
    st_fe(&pts[j].y, p.y);
    ok[j] = 1;
}

// ---- batched point decompression --------------------------------------------------------
// ark-serialize `deserialize_compressed` with validation for short-Weierstrass points (src/r1cs/proof.rs:83-91 calls it
// for the 11 + 2k points of every proof; 1024 proofs of a batch verification are ~44 000 square roots, ~0.4 s on one
// host core): 32 bytes of x (little endian, canonical) + a flag byte (bit 7: y is the larger root, bit 6: infinity).
// One thread per point; same accept/reject decisions as HostCurve::point_from_compressed.
template <class C>
__global__ void __launch_bounds__(128) points_decompress_kernel(const uint8_t* __restrict__ comp, size_t n, const __grid_constant__ SqrtParams sp,
                                                                affine* __restrict__ out, uint8_t* __restrict__ ok) {
    using F = Fp<typename C::Fq>;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint8_t* b = comp + j * 33;
    fe x;
#pragma unroll
    for (int k = 0; k < 8; k++) x.v[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
    const uint32_t flags = b[32];
    affine p;
    p.x = F::zero(); p.y = F::zero();
    bool good = (flags & 0xC0u) != 0xC0u;                      // both SWFlags bits: UnexpectedFlags; the six low bits are padding
    bool geq = true;                                           // x >= q is not canonical
    for (int k = 7; k >= 0; k--) {
        uint32_t mk = C::Fq::m(k);
        if (x.v[k] != mk) { geq = x.v[k] > mk; break; }
    }
    good = good && !geq;
    if (good && (flags & 0x40u)) {                             // infinity: the identity whatever x (< q) is, as ark-ec decides
    } else if (good) {
        fe xm = F::to_mont(x);
        fe rhs = F::add(F::mul(F::sqr(xm), xm), F::template curve_b<C>());
        if (C::A_SMALL != 0) rhs = F::add(rhs, F::mul_small(xm, C::A_SMALL));
        fe y;
        good = fq_sqrt_dev<F>(rhs, sp, y);
        if (good) {
            fe ny = F::neg(y);
            fe yc = F::from_mont(y), nc = F::from_mont(ny);
            bool y_larger = false;
            for (int k = 7; k >= 0; k--)
                if (yc.v[k] != nc.v[k]) { y_larger = yc.v[k] > nc.v[k]; break; }
            p.x = xm;
            p.y = (((flags & 0x80u) != 0) == y_larger) ? y : ny;
        }
    }
    st_fe(&out[j].x, p.x);
    st_fe(&out[j].y, p.y);
    ok[j] = good ? 1 : 0;
}

template <class C>
int points_decompress_device(bp_ctx* ctx, const uint8_t* comp, size_t n, const SqrtParams& sp, affine* out, uint8_t* ok) {
    if (n == 0) return BP_OK;
    DevBuf dc, dp, dk;
    struct Guard { DevBuf* b[3]; ~Guard() { for (auto* x : b) x->release(); } } guard{{&dc, &dp, &dk}};
    BP_CUDA_TRY(ctx, dc.reserve(n * 33));
    BP_CUDA_TRY(ctx, dp.reserve(n * sizeof(affine)));
    BP_CUDA_TRY(ctx, dk.reserve(n));
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(dc.p, comp, n * 33, cudaMemcpyHostToDevice, ctx->stream));
    points_decompress_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(dc.as<uint8_t>(), n, sp, dp.as<affine>(), dk.as<uint8_t>());
    BP_LAUNCH_CHECK(ctx);
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(out, dp.p, n * sizeof(affine), cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(ok, dk.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

// Validation of caller-supplied generator vectors (bp_gens_from_points): canonical coordinates, on the curve, and -- on the
// twisted Edwards curve (cofactor 8) -- r * P = O. Any failure sets *bad.
template <class C>
__global__ void __launch_bounds__(128) points_validate_kernel(const affine* __restrict__ pts, size_t n, int* __restrict__ bad) {
    using E = GroupLaw<C>;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    affine p = ld_affine(pts + j);
    bool good = true;
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        const fe& v = c ? p.y : p.x;
        bool geq = true;
        for (int k = 7; k >= 0; k--) {
            uint32_t mk = C::Fq::m(k);
            if (v.v[k] != mk) { geq = v.v[k] > mk; break; }
        }
        good = good && !geq;
    }
    if (good && !E::is_identity(p)) {
        good = E::on_curve(p);
        if (good && C::KIND == 1) {
            uint32_t rl[8];
            for (int i = 0; i < 8; i++) rl[i] = C::Fr::m(i);
            good = E::is_identity(E::mul_scalar(p, rl));
        }
    }
    if (!good) atomicOr(bad, 1);
}

template <class C>
int points_validate_device(bp_ctx* ctx, const affine* d_pts, size_t n) {
    if (n == 0) return BP_OK;
    DevBuf flag;
    struct Guard { DevBuf* b; ~Guard() { b->release(); } } guard{&flag};
    BP_CUDA_TRY(ctx, flag.reserve(sizeof(int)));
    BP_CUDA_TRY(ctx, cudaMemsetAsync(flag.p, 0, sizeof(int), ctx->stream));
    points_validate_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_pts, n, flag.as<int>());
    BP_LAUNCH_CHECK(ctx);
    int bad = 0;
    BP_CUDA_TRY(ctx, cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return bad ? BP_ERR_FORMAT : BP_OK;
}

// out[j] = in[j * stride + offset]  (this rank's cyclic shard of the chain)
static __global__ void __launch_bounds__(256) gens_stride_kernel(const affine* __restrict__ in, size_t n_out, size_t stride, size_t offset,
                                                                 affine* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    affine p = ld_affine(in + j * stride + offset);
    st_fe(&out[j].x, p.x);
    st_fe(&out[j].y, p.y);
}

// First `count` points of the chain seeded by `seed`, shard (rank, world) of them into d_out. Returns BP_ERR_UNSUPPORTED when the
// stream turned out irregular (caller falls back to the host generator).
template <class C>
int gens_chain_device(bp_ctx* ctx, const uint8_t seed[32], const SqrtParams& sp, size_t count, int rank, int world, affine* d_out) {
    cudaStream_t st = ctx->stream;
    ChaChaKey key;
    memcpy(key.k, seed, 32);
    DevBuf full, att, okb, sel, misc, tmp;
    struct Guard { DevBuf* b[6]; ~Guard() { for (auto* x : b) x->release(); } } guard{{&full, &att, &okb, &sel, &misc, &tmp}};
    BP_CUDA_TRY(ctx, full.reserve((count + 1) * sizeof(affine)));
    BP_CUDA_TRY(ctx, misc.reserve(64));
    int* d_irregular = misc.as<int>();
    int* d_nsel = misc.as<int>() + 4;
    BP_CUDA_TRY(ctx, cudaMemsetAsync(misc.p, 0, 64, st));
    size_t produced = 0;
    uint64_t attempt0 = 0;
    // A 256-bit base field well below 2^256 (zorro: q ~ 2^255) rejects about half of the x draws, each rejected draw eating
    // eight words: where an attempt starts depends on every draw before it. That scan is inherently serial and cheap (a
    // ChaCha20 keystream and one 256-bit comparison per draw), so the host walks the stream and hands the device the word
    // offset of every draw that Fp::rand accepts; the square roots -- all of the arithmetic -- stay on the device.
    constexpr bool x_rejects = C::Fq::BITS == 256 && !(C::Fq::m(7) == 0xFFFFFFFFu && C::Fq::m(6) == 0xFFFFFFFFu && C::Fq::m(5) == 0xFFFFFFFFu);
    ChaCha20Rng walker(seed);
    std::vector<uint32_t> h_off;
    DevBuf d_off;
    struct OffGuard { DevBuf* b; ~OffGuard() { b->release(); } } off_guard{&d_off};
    while (produced < count) {
        size_t want = (count - produced) * 2 + 1024;
        if (want > ((size_t)1 << 30)) return BP_ERR_LEN;
        BP_CUDA_TRY(ctx, att.reserve(want * sizeof(affine)));
        BP_CUDA_TRY(ctx, sel.reserve(want * sizeof(affine)));
        BP_CUDA_TRY(ctx, okb.reserve(want));
        const uint32_t* offs = nullptr;
        if (x_rejects) {
            h_off.resize(want);
            for (size_t k = 0; k < want;) {
                const uint64_t at = walker.words_used;
                if (at + 9 > 0xFFFFFFFFull) return BP_ERR_LEN;
                uint64_t l[4];
                for (int q = 0; q < 4; q++) l[q] = walker.next_u64();
                if (HostFp<typename C::Fq>::geq_m(l)) continue;        // Fp::rand redraws: the next draw starts 8 words on
                walker.next_u32();                                     // the `greatest` word of this attempt
                h_off[k++] = (uint32_t)at;
            }
            BP_CUDA_TRY(ctx, d_off.reserve(want * sizeof(uint32_t)));
            BP_CUDA_TRY(ctx, cudaMemcpyAsync(d_off.p, h_off.data(), want * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            offs = d_off.as<uint32_t>();
        }
        gens_attempt_kernel<C><<<(unsigned)((want + 127) / 128), 128, 0, st>>>(key, attempt0, want, offs, sp, att.as<affine>(), okb.as<uint8_t>(), d_irregular);
        BP_LAUNCH_CHECK(ctx);
        size_t tb = 0;
        BP_CUDA_TRY(ctx, cub::DeviceSelect::Flagged(nullptr, tb, att.as<affine>(), okb.as<uint8_t>(), sel.as<affine>(), d_nsel, (int)want, st));
        BP_CUDA_TRY(ctx, tmp.reserve(tb));
        BP_CUDA_TRY(ctx, cub::DeviceSelect::Flagged(tmp.p, tb, att.as<affine>(), okb.as<uint8_t>(), sel.as<affine>(), d_nsel, (int)want, st));
        int h[8];
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(h, misc.p, 32, cudaMemcpyDeviceToHost, st));
        BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        if (h[0]) return BP_ERR_UNSUPPORTED;
        size_t got = (size_t)h[4];
        size_t take = got < count - produced ? got : count - produced;
        BP_CUDA_TRY(ctx, cudaMemcpyAsync(full.as<affine>() + produced, sel.p, take * sizeof(affine), cudaMemcpyDeviceToDevice, st));
        produced += take;
        attempt0 += want;
    }
    size_t r = (size_t)rank, w = (size_t)world;
    size_t n_out = count > r ? (count - r + w - 1) / w : 0;
    if (n_out) {
        gens_stride_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(full.as<affine>(), n_out, w, r, d_out);
        BP_LAUNCH_CHECK(ctx);
    }
    BP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return BP_OK;
}

}  // namespace bp

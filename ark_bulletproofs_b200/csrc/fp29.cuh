// Unsaturated 9 x 29-bit field elements for the bucket-accumulation inner loop.
//
// Why: on B200 every carry-*producing* IMAD.WIDE occupies the fmaheavy pipe for two passes
// (31 wide MADs/clk/SM, profiles/r1_mul29_experiment.txt), while a plain IMAD.WIDE issues at the full
// IMAD rate (59/clk/SM). The 8 x 32-bit CIOS of fp.cuh is therefore pinned at ~514 pipe cycles per warp
// modmul. Here products of 29-bit limbs are accumulated into 18 64-bit columns with *no* carries
// (<= 14 products of < 2^58 per column), and -- unlike the round-1 experiment that converted at every
// multiplication and lost to ALU pressure -- elements STAY in this form across a whole kernel:
//   * Montgomery domain 2^261 (nine 29-bit steps): no alignment shifts, and since 2^261 >= 32 m the
//     product of two elements < 2m reduces to (-m, m/8): adding m gives a positive result < 2m with
//     no conditional correction and no top-bit fold after a multiplication;
//   * m = 2^E - c with c < 2^145: a reduction step is q*c (5 multiplies) - q*2^E (one signed multiply by a
//     constant) instead of 9 multiplies;
//   * add/sub are limb-wise with one "weak reduction" (ripple + fold the bits above 2^E with 2^E = c mod m).
// Contract ("tight"): limbs 0..7 < 2^29, limb 8 < 2^(TOP_SH+1), value < 2m, value = x * 2^261 mod m up to
// one multiple of m. Only moduli of the form 2^E - c are supported (RED_NEG29): secq256k1's base field and
// 2^255 - 19; zorro's base field keeps the 32-bit path.
//
// Plain C++ (64-bit integer arithmetic): the same code runs in the host unit tests.
#pragma once
#include "fp.cuh"

// mul/sqr are real functions on the device: fully inlined, the 10 multiplications of a mixed addition are
// ~4 k instructions and the accumulate loop stalls on instruction fetch (ncu: stall_no_instruction 6.7/issue)
#if defined(__CUDACC__)
#define BP_HD_NOINL29 __host__ __device__ __noinline__
#else
#define BP_HD_NOINL29 inline
#endif

namespace bp {

struct fl { uint32_t v[9]; };

template <class M>
struct Fp29 {
    using Mod = M;
    using el = fl;
    static constexpr uint32_t MASK = 0x1FFFFFFFu;
    static_assert(M::RED_NEG29, "Fp29 needs a modulus of the form 2^E - c, c < 2^145");

    template <class C> BP_HD static fl te_d2() { fl r; for (int i = 0; i < 9; i++) r.v[i] = C::d2_29(i); return r; }
    BP_HD static fl from_u32(uint32_t) { return zero(); }   // only reached for curves with a != 0, which keep the 32-bit path
    BP_HD static fl zero() { fl r; for (int i = 0; i < 9; i++) r.v[i] = 0; return r; }
    BP_HD static fl one() { fl r; for (int i = 0; i < 9; i++) r.v[i] = M::one29(i); return r; }

    // ripple + fold + ripple: input limbs < 2^32 - 2^29, value < 8m  ->  tight
    BP_HD static fl weak_reduce(fl a) {
#pragma unroll
        for (int k = 0; k < 8; k++) { a.v[k + 1] += a.v[k] >> 29; a.v[k] &= MASK; }
        const uint32_t t = a.v[8] >> M::TOP_SH;            // value >> E, at most 7
        a.v[8] &= (1u << M::TOP_SH) - 1u;
#pragma unroll
        for (int k = 0; k < 5; k++)
            if (M::c29(k) != 0) a.v[k] += t * M::c29(k);   // 2^E = c (mod m)
#pragma unroll
        for (int k = 0; k < 8; k++) { a.v[k + 1] += a.v[k] >> 29; a.v[k] &= MASK; }
        return a;
    }
    BP_HD static fl add(const fl& a, const fl& b) {
        fl r;
#pragma unroll
        for (int k = 0; k < 9; k++) r.v[k] = a.v[k] + b.v[k];
        return weak_reduce(r);
    }
    BP_HD static fl sub(const fl& a, const fl& b) {          // a + 4m - b, every limb stays non-negative
        fl r;
#pragma unroll
        for (int k = 0; k < 9; k++) r.v[k] = a.v[k] + M::k4m29(k) - b.v[k];
        return weak_reduce(r);
    }
    BP_HD static fl dbl(const fl& a) { return add(a, a); }
    BP_HD static fl mul3(const fl& a) {
        fl r;
#pragma unroll
        for (int k = 0; k < 9; k++) r.v[k] = 3u * a.v[k];
        return weak_reduce(r);
    }
    BP_HD static fl neg(const fl& a) { return sub(zero(), a); }
    BP_HD static fl mul_small(const fl& a, int k) {
        fl r = zero(), p = a;
        for (int bit = 0; bit < 4; bit++) {
            if ((k >> bit) & 1) r = add(r, p);
            p = dbl(p);
        }
        return r;
    }
    // exact: a tight element is 0 mod m iff it is the integer 0 or the integer m
    BP_HD static bool is_zero(const fl& a) {
        uint32_t z = 0, e = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) { z |= a.v[k]; e |= a.v[k] ^ M::m29(k); }
        return z == 0 || e == 0;
    }
    BP_HD static bool eq(const fl& a, const fl& b) { return is_zero(sub(a, b)); }

    // Montgomery reduction of 18 columns (columns 9..17 pre-loaded with m) -> tight element
    BP_HD static fl reduce(int64_t (&c)[18]) {
        constexpr int64_t ETERM = -(int64_t)(1ull << M::E_SH);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            if (i > 0) c[i] += c[i - 1] >> 29;
            const uint32_t q = ((uint32_t)c[i] * M::MINV29) & MASK;
#pragma unroll
            for (int j = 0; j < 5; j++)
                if (M::c29(j) != 0) c[i + j] += (int64_t)((uint64_t)q * M::c29(j));
            c[i + M::E_COL] += (int64_t)q * ETERM;
        }
        fl r;
        c[9] += c[8] >> 29;
#pragma unroll
        for (int k = 9; k < 17; k++) {
            c[k + 1] += c[k] >> 29;
            r.v[k - 9] = (uint32_t)c[k] & MASK;
        }
        r.v[8] = (uint32_t)c[17];
        return r;
    }
    BP_HD_NOINL29 static fl mul(const fl& a, const fl& b) {
        int64_t c[18];
#pragma unroll
        for (int k = 0; k < 9; k++) { c[k] = 0; c[9 + k] = M::m29(k); }
#pragma unroll
        for (int i = 0; i < 9; i++)
#pragma unroll
            for (int j = 0; j < 9; j++) c[i + j] += (int64_t)((uint64_t)a.v[i] * b.v[j]);
        return reduce(c);
    }
    BP_HD_NOINL29 static fl sqr(const fl& a) {
        int64_t c[18];
        uint32_t a2[9];
#pragma unroll
        for (int k = 0; k < 9; k++) { c[k] = 0; c[9 + k] = M::m29(k); a2[k] = a.v[k] << 1; }
#pragma unroll
        for (int i = 0; i < 9; i++) {
            c[2 * i] += (int64_t)((uint64_t)a.v[i] * a.v[i]);
#pragma unroll
            for (int j = i + 1; j < 9; j++) c[i + j] += (int64_t)((uint64_t)a.v[i] * a2[j]);
        }
        return reduce(c);
    }

    // ---- conversions ------------------------------------------------------------------------
    // plain re-limbing of a canonical 256-bit value (no domain change)
    BP_HD static fl unpack(const fe& a) {
        fl r;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            int bit = 29 * k, w = bit >> 5, sh = bit & 31;
            uint64_t lo = a.v[w];
            uint64_t hi = (w + 1 < 8) ? a.v[w + 1] : 0u;
            r.v[k] = (uint32_t)(((hi << 32) | lo) >> sh) & MASK;
        }
        return r;
    }
    // tight element -> canonical integer in [0, m) packed as 8 x 32 bits (no domain change)
    BP_HD static fe pack_canonical(const fl& a) {
        // a < 2m: subtract m if a >= m (29-bit limb compare/subtract with borrow)
        uint32_t d[9];
        int32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            int64_t t = (int64_t)a.v[k] - (int64_t)M::m29(k) + borrow;
            d[k] = (uint32_t)t & MASK;
            borrow = (int32_t)(t >> 29);
        }
        // for k = 8 the mask above is harmless: both operands are < 2^26
        const bool ge = borrow >= 0;
        uint32_t r[10];
#pragma unroll
        for (int k = 0; k < 9; k++) r[k] = ge ? d[k] : a.v[k];
        r[9] = 0;
        fe o;
#pragma unroll
        for (int w = 0; w < 8; w++) o.v[w] = (uint32_t)(((((uint64_t)r[w + 1]) << 29) | r[w]) >> (3 * w));
        return o;
    }
    // storage (Montgomery 2^256, canonical) -> working domain (2^261), given the value already
    // multiplied by 2^5 in storage form (see msm_to29_kernel); and back
    BP_HD static fe to_storage(const fl& a) {
        fl k;
#pragma unroll
        for (int i = 0; i < 9; i++) k.v[i] = M::r256_29(i);
        return pack_canonical(mul(a, k));                  // (x*2^261) * 2^256 / 2^261 = x*2^256
    }
};

}  // namespace bp

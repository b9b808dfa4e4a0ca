// Balanced 9 x 29-bit field elements with carry-free column accumulation.
//
// EXPERIMENT, not linked into the product (profiles/r2_microbench3_fp29_vs_cios.json). The premise
// -- round 1's reading that only a *carry-chained* IMAD.WIDE is a two-pass instruction -- turned out to be wrong: a plain
// IMAD.WIDE Rd = a*b + Rc also holds the multiplier for two issue slots (4 clk per warp instruction). This layer has not
// a single flag instruction in a multiplication (81 plain IMAD.WIDE into 18 signed 64-bit columns, a Montgomery reduction
// of 9 x (1 IMAD + one IMAD.WIDE per non-zero limb of m), carries moved with IMAD.WIDE too: 152 + 9 pipe instructions for
// secq256k1's base field, 116 + 9 for a squaring) and runs at 56.9 G modmul/s against 66.2 for the 8 x 32 CIOS of fp.cuh
// (128 + 8 products); only its squaring wins (72.3). It stays as a verified reference point: host-tested against the
// oracle on all five moduli and three curves (tests/test_hostmath.py), bit-identical to the CIOS on the device
// (tools/microbench3.cu).
//
// What makes the flag-free form possible is the **balanced** digit set: limbs 0..7 lie in [-2^28, 2^28], so |a_i * b_j| <= 2^56
// and a column (9 products + <= 6 reduction terms + a carry) stays below 2^60 -- its carry (column >> 29) fits the
// 32-bit multiplicand of IMAD.WIDE. With unsigned 29-bit limbs the same column reaches 2^61.5 and the carry needs
// 33 bits (round 1's fp29 resolved carries with IADD3/IADD3.X pairs and lost to the ALU traffic; ptxas moreover
// re-associates chains of `mad.wide` into IMAD.WIDE + 3-input IADD3/IADD3.X trees -- the (mad.lo.cc, madc.hi) pair
// form below is what keeps the accumulate inside the multiplier).
//
// Representation: fl = 9 x int32, value V = sum v[k] * 2^(29k), V = x * 2^261 (mod m) (Montgomery radix 2^261: nine
// 29-bit steps, no alignment shifts). Not canonical: any V = x * 2^261 + k*m is a valid representative.
//   "normalized"  limbs 0..7 in [-2^28, 2^28], limb 8 free with |v[8]| < 2^28 (i.e. |V| < 16 m): the only
//                 requirement on the operands of mul/sqr.
//   mul/sqr       return a normalized element with |V| <= |Va|*|Vb| / 2^261 + m/2  (<= 1.01 m for operands <= 4 m).
//   add/sub/...   limb-wise followed by one ripple (norm); the *_l variants skip the ripple and may only feed
//                 other additions (limbs must stay below 2^31).
// Moduli: any of the five with at most 6 non-zero balanced limbs (all of them: a sparse high part).
//
// The whole file is plain C++ on 64-bit integers apart from the IMAD.WIDE pair, so the same code runs in the host
// unit tests (tests/test_hostmath.py) against the oracle's big integers.
#pragma once
#include "../fp.cuh"

namespace bp {

struct fl { int32_t v[9]; };

// Host unit tests compile with BP_FP29_CHECK: every multiplication verifies the operand contract (balanced limbs,
// |v[8]| < 2^28) and that no column leaves the range whose carry fits 32 bits; violations are counted.
#if defined(BP_FP29_CHECK) && !defined(__CUDA_ARCH__)
inline long& fp29_violations() { static long n = 0; return n; }
inline void fp29_check_operand(const fl& a) {
    for (int i = 0; i < 8; i++)
        if (a.v[i] < -(1 << 28) || a.v[i] > (1 << 28)) fp29_violations()++;
    if (a.v[8] <= -(1 << 28) || a.v[8] >= (1 << 28)) fp29_violations()++;
}
inline void fp29_check_column(int64_t c) {
    if (c >= (1ll << 60) || c < -(1ll << 60)) fp29_violations()++;
}
#define BP_FP29_OPERAND(a) fp29_check_operand(a)
#define BP_FP29_COLUMN(c) fp29_check_column(c)
#else
#define BP_FP29_OPERAND(a)
#define BP_FP29_COLUMN(c)
#endif

#if defined(__CUDACC__)
// ptxas folds an immediate 1 into IADD3 + IMAD.HI; a constant-bank operand keeps `column += carry * 1` one IMAD.WIDE
static __device__ __constant__ int32_t bp_one29 = 1;
#endif

template <class M>
struct Fp29 {
    using Mod = M;
    using el = fl;
    static constexpr uint32_t MASK = 0x1FFFFFFFu;
    static constexpr int32_t HALF = 1 << 28;

    // c += a * b on a 64-bit column, no flags in or out (IMAD.WIDE R, a, b, R)
    BP_HD static void mac(int64_t& c, int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
        uint32_t lo = (uint32_t)c, hi = (uint32_t)((uint64_t)c >> 32);
        asm("mad.lo.cc.s32 %0, %2, %3, %0; madc.hi.s32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
        c = (int64_t)(((uint64_t)hi << 32) | lo);
#else
        c += (int64_t)a * (int64_t)b;
#endif
    }
    BP_HD static void addcarry(int64_t& c, int32_t carry) {
#if defined(__CUDA_ARCH__)
        mac(c, carry, bp_one29);
#else
        c += carry;
#endif
    }
    BP_HD static int32_t sx29(uint32_t x) {   // sign-extend the low 29 bits (SGXT)
#if defined(__CUDA_ARCH__)
        int32_t r;
        asm("bfe.s32 %0, %1, 0, 29;" : "=r"(r) : "r"(x));
        return r;
#else
        return (int32_t)(x << 3) >> 3;
#endif
    }
    BP_HD static int32_t sar29(int32_t x) { return x >> 29; }

    BP_HD static fl zero() { fl r; for (int i = 0; i < 9; i++) r.v[i] = 0; return r; }
    BP_HD static fl one() { fl r; for (int i = 0; i < 9; i++) r.v[i] = M::oneb(i); return r; }
    template <class C> BP_HD static fl te_d2() { fl r; for (int i = 0; i < 9; i++) r.v[i] = C::d2b(i); return r; }
    template <class C> BP_HD static fl curve_b() { fl r; for (int i = 0; i < 9; i++) r.v[i] = C::bb(i); return r; }

    // ---- multiplication ----------------------------------------------------------------------
    // Montgomery reduction of the 18 columns (9..16 pre-loaded with 2^28 for the balanced digit extraction)
    BP_HD static fl reduce(int64_t (&c)[18]) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int32_t q = sx29((uint32_t)c[i] * M::NINV29);      // c[i] + q*m_0 = 0 (mod 2^29), |q| <= 2^28
#pragma unroll
            for (int j = 0; j < 9; j++)
                if (M::mb(j) != 0) mac(c[i + j], q, M::mb(j));
            BP_FP29_COLUMN(c[i]);
            addcarry(c[i + 1], (int32_t)(c[i] >> 29));               // exact: the low 29 bits are zero
        }
        fl r;
#pragma unroll
        for (int k = 9; k < 17; k++) {
            BP_FP29_COLUMN(c[k]);
            r.v[k - 9] = (int32_t)((uint32_t)c[k] & MASK) - HALF;    // digit of (c - 2^28) in [-2^28, 2^28)
            addcarry(c[k + 1], (int32_t)(c[k] >> 29));
        }
        r.v[8] = (int32_t)c[17];
        return r;
    }
    BP_HD static fl mul(const fl& a, const fl& b) {
        BP_FP29_OPERAND(a); BP_FP29_OPERAND(b);
        int64_t c[18];
#pragma unroll
        for (int k = 0; k < 18; k++) c[k] = (k >= 9 && k < 17) ? (int64_t)HALF : 0;
#pragma unroll
        for (int i = 0; i < 9; i++)
#pragma unroll
            for (int j = 0; j < 9; j++) mac(c[i + j], a.v[i], b.v[j]);
        return reduce(c);
    }
    BP_HD static fl sqr(const fl& a) {
        BP_FP29_OPERAND(a);
        int64_t c[18];
        int32_t a2[9];
#pragma unroll
        for (int k = 0; k < 18; k++) c[k] = (k >= 9 && k < 17) ? (int64_t)HALF : 0;
#pragma unroll
        for (int k = 0; k < 9; k++) a2[k] = a.v[k] * 2;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            mac(c[2 * i], a.v[i], a.v[i]);
#pragma unroll
            for (int j = i + 1; j < 9; j++) mac(c[i + j], a.v[i], a2[j]);
        }
        return reduce(c);
    }

    // ---- additive layer ------------------------------------------------------------------------
    // ripple to balanced digits; input limbs anywhere in int32 with |limb| < 2^31 - 2^28
    BP_HD static fl norm(fl a) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int32_t t = a.v[k] + HALF;
            a.v[k + 1] += sar29(t);
            a.v[k] = (int32_t)((uint32_t)t & MASK) - HALF;
        }
        return a;
    }
    BP_HD static fl add_l(const fl& a, const fl& b) { fl r; for (int k = 0; k < 9; k++) r.v[k] = a.v[k] + b.v[k]; return r; }
    BP_HD static fl sub_l(const fl& a, const fl& b) { fl r; for (int k = 0; k < 9; k++) r.v[k] = a.v[k] - b.v[k]; return r; }
    BP_HD static fl dbl_l(const fl& a) { fl r; for (int k = 0; k < 9; k++) r.v[k] = a.v[k] * 2; return r; }
    BP_HD static fl add(const fl& a, const fl& b) { return norm(add_l(a, b)); }
    BP_HD static fl sub(const fl& a, const fl& b) { return norm(sub_l(a, b)); }
    BP_HD static fl mul_sub(const fl& a, const fl& b, const fl& c, const fl& d) { return sub(mul(a, b), mul(c, d)); }
    BP_HD static fl dbl(const fl& a) { return norm(dbl_l(a)); }
    BP_HD static fl mul3(const fl& a) { fl r; for (int k = 0; k < 9; k++) r.v[k] = a.v[k] * 3; return norm(r); }
    BP_HD static fl neg(const fl& a) { fl r; for (int k = 0; k < 9; k++) r.v[k] = -a.v[k]; return r; }
    // k * 2^261 for a small curve coefficient (k <= 8)
    BP_HD static fl from_u32(uint32_t k) { fl r; for (int i = 0; i < 9; i++) r.v[i] = M::oneb(i) * (int32_t)k; return norm(r); }
    // a * k for small k: a full multiplication, so the value bound does not grow
    BP_HD static fl mul_small(const fl& a, int k) { return mul(a, from_u32((uint32_t)k)); }

    // V = 0 (mod m)?  Requires |V| < 8 m and limbs < 2^31. k = round(V / m) from the top limb, then V - k*m must be the
    // integer 0 (checked with a floor ripple in 64-bit arithmetic; rare path).
    BP_HD_NOINL static bool is_zero_slow(const fl& a, int32_t k) {
        int64_t carry = 0;
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int64_t t = (int64_t)a.v[i] - (int64_t)k * M::mb(i) + carry;
            carry = t >> 29;
            nz |= (i < 8) ? ((uint32_t)t & MASK) : ((uint32_t)t | (uint32_t)((uint64_t)t >> 32));
        }
        return nz == 0;
    }
    BP_HD static bool is_zero(const fl& a) {
        const int32_t k = (a.v[8] + (1 << (M::TOPB - 1))) >> M::TOPB;   // m ~ 2^(232 + TOPB)
        // a multiple of m matches k*m in its low 29 bits; anything else fails here with probability 1 - 2^-29
        if ((((uint32_t)a.v[0] - (uint32_t)k * (uint32_t)M::mb(0)) & MASK) != 0) return false;
        return is_zero_slow(a, k);
    }
    BP_HD static bool eq(const fl& a, const fl& b) { return is_zero(sub_l(a, b)); }

    // ---- conversions ---------------------------------------------------------------------------
    // re-limbing of a 256-bit integer into balanced digits (no domain change)
    BP_HD static fl unpack(const fe& a) {
        fl r;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int bit = 29 * k, w = bit >> 5, sh = bit & 31;
            const uint64_t lo = a.v[w];
            const uint64_t hi = (w + 1 < 8) ? a.v[w + 1] : 0u;
            r.v[k] = (int32_t)((uint32_t)(((hi << 32) | lo) >> sh) & MASK);
        }
        return norm(r);
    }
    // storage (Montgomery 2^256, canonical) -> working form: one multiplication by 2^266
    BP_HD static fl from_storage(const fe& a) {
        fl k;
#pragma unroll
        for (int i = 0; i < 9; i++) k.v[i] = M::k266b(i);
        return mul(unpack(a), k);
    }
    // The MSM keeps its bases pre-converted in 64 bytes per point (msm_to29_kernel): the canonical integer
    // U = 32*X + D (mod m), D = 2^28 * sum_{k<8} 2^(29k), with bit 28 of every 29-bit group flipped. Then limb k of the
    // working form is just the sign extension of bit group k -- SHF + SGXT per limb, no ripple: U - D = sum (u_k - 2^28) 2^(29k).
    BP_HD static fe to_biased(const fe& storage_x32_plus_d) {
        fe r = storage_x32_plus_d;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int bit = 29 * k + 28;
            r.v[bit >> 5] ^= 1u << (bit & 31);
        }
        return r;
    }
    BP_HD static fe bias_d() {   // D as a 256-bit integer
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int bit = 29 * k + 28;
            r.v[bit >> 5] |= 1u << (bit & 31);
        }
        return r;
    }
    BP_HD static fl load_biased(const fe& a) {
        fl r;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int bit = 29 * k, w = bit >> 5, sh = bit & 31;
            const uint64_t lo = a.v[w];
            const uint64_t hi = (w + 1 < 8) ? a.v[w + 1] : 0u;
            const uint32_t g = (uint32_t)(((hi << 32) | lo) >> sh);
            r.v[k] = (k < 8) ? sx29(g) : (int32_t)g;
        }
        return r;
    }
    // canonical integer in [0, m) of V mod m, packed as 8 x 32 bits (no domain change); |V| < 8 m
    BP_HD_NOINL static fe pack_canonical(const fl& a0) {
        const fl a = norm(a0);
        const int32_t k = (a.v[8] + (1 << (M::TOPB - 1))) >> M::TOPB;      // V - k*m lies in (-0.51 m, 0.51 m)
        int32_t u[9];
        int64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {                                      // floor ripple: limbs 0..7 in [0, 2^29), sign in limb 8
            const int64_t t = (int64_t)a.v[i] - (int64_t)k * M::mb(i) + carry;
            carry = (i < 8) ? (t >> 29) : 0;
            u[i] = (i < 8) ? (int32_t)((uint32_t)t & MASK) : (int32_t)t;
        }
        auto ripple = [](int32_t (&w)[9]) {
#pragma unroll
            for (int i = 0; i < 8; i++) { w[i + 1] += w[i] >> 29; w[i] = (int32_t)((uint32_t)w[i] & MASK); }
        };
        if (u[8] < 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) u[i] += (int32_t)M::m29u(i);
            ripple(u);
        }
        {
            int32_t d[9];
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] = u[i] - (int32_t)M::m29u(i);
            ripple(d);
            if (d[8] >= 0) {
#pragma unroll
                for (int i = 0; i < 9; i++) u[i] = d[i];
            }
        }
        fe o;
#pragma unroll
        for (int w = 0; w < 8; w++) {                                      // bits [32w, 32w + 32) of sum u[k] 2^(29k)
            const int k0 = (32 * w) / 29, off = 32 * w - 29 * k0;
            uint64_t acc = (uint64_t)(uint32_t)u[k0] >> off;
            acc |= (uint64_t)(uint32_t)u[k0 + 1] << (29 - off);
            if (k0 + 2 < 9) acc |= (uint64_t)(uint32_t)u[k0 + 2] << (58 - off);
            o.v[w] = (uint32_t)acc;
        }
        return o;
    }
    // working form -> canonical storage (Montgomery 2^256, [0, m), 8 x 32 bits): one multiplication by 2^256
    BP_HD static fe to_storage(const fl& a) {
        fl k;
#pragma unroll
        for (int i = 0; i < 9; i++) k.v[i] = M::r256b(i);
        return pack_canonical(mul(a, k));
    }

    // a^e, e as 8 little-endian 32-bit limbs
    BP_HD_NOINL static fl pow(const fl& a, const uint32_t* e) {
        fl r = one();
        bool started = false;
        for (int i = 7; i >= 0; i--) {
            for (int bit = 31; bit >= 0; bit--) {
                if (started) r = sqr(r);
                if ((e[i] >> bit) & 1u) {
                    r = started ? mul(r, a) : a;
                    started = true;
                }
            }
        }
        return r;
    }
    // Fermat inverse; inv(0) = 0
    BP_HD_NOINL static fl inv(const fl& a) {
        uint32_t e[8];
        e[0] = M::m(0) - 2u;
        for (int i = 1; i < 8; i++) e[i] = M::m(i);
        return pow(a, e);
    }
};

#if defined(__CUDACC__)
// 36-byte elements are moved as whole 144-byte XYZZ records (nine 16-byte words)
template <int NW>
__device__ __forceinline__ void ld_words(int32_t* dst, const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < NW / 4; i++) {
        uint4 a = q[i];
        dst[4 * i] = (int32_t)a.x; dst[4 * i + 1] = (int32_t)a.y; dst[4 * i + 2] = (int32_t)a.z; dst[4 * i + 3] = (int32_t)a.w;
    }
}
template <int NW>
__device__ __forceinline__ void st_words(void* p, const int32_t* src) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < NW / 4; i++)
        q[i] = make_uint4((uint32_t)src[4 * i], (uint32_t)src[4 * i + 1], (uint32_t)src[4 * i + 2], (uint32_t)src[4 * i + 3]);
}
#endif

}  // namespace bp

// 256-bit Montgomery prime-field arithmetic for sm_100a (8 x 32-bit limbs, R = 2^256).
//
// Replaces ark-ff `Fp256<MontBackend<_,4>>` (reference: Cargo.toml:28-35; zorro
// Fq at src/curve/zorro/fq.rs:3-7) for the five moduli of SURVEY.md App. A.1.
// Memory layout is identical to ark-ff's ([u64;4] little-endian limbs holding
// value*2^256 mod m), so buffers cross the C ABI without conversion.
//
// The multiplier is an even/odd-column CIOS: every 32x32 product is issued as an
// adjacent (mad.lo.cc, madc.hi.cc) pair on a 64-bit column, which ptxas fuses into one
// IMAD.WIDE.U32(.X) with predicate carries -- 128 wide multiplies + 8 for the quotients per
// modmul (checked with cuobjdump; see profiles/).
//
// Two unsaturated variants were built and measured against it: round 1's 9 x 29-bit columns with ALU
// carries (46 G modmul/s vs 66.5), and round 2's balanced 9 x 29-bit layer without a single flag instruction
// (experimental/fp29.cuh: 56.9 vs 66.2). Both lose for one reason: an IMAD.WIDE costs two multiplier passes with or without
// carry predicates, and this CIOS has the fewest 32x32->64 products (128 + 8) of any 8-limb formulation.
//
// Every value handed between functions is fully reduced (< m): the secq256k1
// moduli are within 2^129 of 2^256, so there are no spare bits for lazy reduction.
//
// The carry primitives have a host emulation (a thread-local carry flag) so that the very
// same templates run in the CPU unit tests and in the O(log n) host-side glue; the bulk
// paths only ever run on the device.
#pragma once
#include <cstdint>
#include "consts.cuh"

#if defined(__CUDACC__)
#define BP_HD __host__ __device__ __forceinline__
#define BP_HD_NOINL __host__ __device__
#else
#define BP_HD inline
#define BP_HD_NOINL inline
#endif

namespace bp {

struct alignas(16) fe { uint32_t v[8]; };

#if !defined(__CUDA_ARCH__)
namespace hostcc {
inline uint32_t& cc() { static thread_local uint32_t c = 0; return c; }
inline uint32_t addx(uint32_t a, uint32_t b, uint32_t cin, bool setcc) {
    uint64_t s = (uint64_t)a + b + cin;
    if (setcc) cc() = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t subx(uint32_t a, uint32_t b, uint32_t bin, bool setcc) {
    uint64_t s = (uint64_t)a - b - bin;
    if (setcc) cc() = (uint32_t)((s >> 32) & 1);
    return (uint32_t)s;
}
}  // namespace hostcc
#endif

// ---- carry-chain primitives -------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define BP_ASM2(name, ins) \
    __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b) { uint32_t r; asm volatile(ins " %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BP_ASM2(add_cc, "add.cc.u32")
BP_ASM2(addc_cc, "addc.cc.u32")
BP_ASM2(addc, "addc.u32")
BP_ASM2(sub_cc, "sub.cc.u32")
BP_ASM2(subc_cc, "subc.cc.u32")
BP_ASM2(subc, "subc.u32")
#undef BP_ASM2
// 64-bit column (hi:lo) += a*b, starting a carry chain
__device__ __forceinline__ void wmad_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b + carry, continuing the chain
__device__ __forceinline__ void wmadc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0,%2,%3,%0; madc.hi.cc.u32 %1,%2,%3,%1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (dhi:dlo) = a*b + (shi:slo) + carry
__device__ __forceinline__ void wmadc_to_cc(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t slo, uint32_t shi) {
    asm volatile("madc.lo.cc.u32 %0,%2,%3,%4; madc.hi.cc.u32 %1,%2,%3,%5;" : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b), "r"(slo), "r"(shi));
}
#else
inline uint32_t add_cc(uint32_t a, uint32_t b) { return hostcc::addx(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return hostcc::addx(a, b, hostcc::cc(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return hostcc::addx(a, b, hostcc::cc(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return hostcc::subx(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return hostcc::subx(a, b, hostcc::cc(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return hostcc::subx(a, b, hostcc::cc(), false); }
inline void wmad_impl(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t slo, uint32_t shi, uint32_t cin) {
    unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)shi << 32) | slo) + cin;
    dlo = (uint32_t)t;
    dhi = (uint32_t)(t >> 32);
    hostcc::cc() = (uint32_t)(t >> 64);
}
inline void wmad_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { wmad_impl(lo, hi, a, b, lo, hi, 0); }
inline void wmadc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { wmad_impl(lo, hi, a, b, lo, hi, hostcc::cc()); }
inline void wmadc_to_cc(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t slo, uint32_t shi) {
    wmad_impl(dlo, dhi, a, b, slo, shi, hostcc::cc());
}
#endif

template <class M>
struct Fp {
    using Mod = M;
    using el = fe;
    template <class C> BP_HD static fe te_d2() { fe r; for (int i = 0; i < 8; i++) r.v[i] = C::d2(i); return r; }
    template <class C> BP_HD static fe curve_b() { fe r; for (int i = 0; i < 8; i++) r.v[i] = C::b(i); return r; }
    // r = t - m if t >= m, where t carries a possible 257th bit `top`
    BP_HD static void final_sub(fe& r, const uint32_t* t, uint32_t top) {
        uint32_t s[8];
        s[0] = sub_cc(t[0], M::m(0));
#pragma unroll
        for (int i = 1; i < 8; i++) s[i] = subc_cc(t[i], M::m(i));
        uint32_t borrow = subc(top, 0u);   // negative iff t < m
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = ((int32_t)borrow < 0) ? t[i] : s[i];
    }

    BP_HD static fe add(const fe& a, const fe& b) {
        uint32_t t[8];
        t[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) t[i] = addc_cc(a.v[i], b.v[i]);
        uint32_t top = addc(0u, 0u);
        fe r;
        final_sub(r, t, top);
        return r;
    }

    BP_HD static fe sub(const fe& a, const fe& b) {
        uint32_t t[8];
        t[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) t[i] = subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = subc(0u, 0u);    // 0xFFFFFFFF if a < b
        fe r;
        r.v[0] = add_cc(t[0], borrow & M::m(0));
#pragma unroll
        for (int i = 1; i < 7; i++) r.v[i] = addc_cc(t[i], borrow & M::m(i));
        r.v[7] = addc(t[7], borrow & M::m(7));
        return r;
    }

    BP_HD static fe zero() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    BP_HD static bool is_zero(const fe& a) {
        uint32_t o = a.v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) o |= a.v[i];
        return o == 0;
    }
    BP_HD static bool eq(const fe& a, const fe& b) {
        uint32_t o = a.v[0] ^ b.v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) o |= a.v[i] ^ b.v[i];
        return o == 0;
    }
    BP_HD static fe neg(const fe& a) { return is_zero(a) ? a : sub(zero(), a); }
    BP_HD static fe dbl(const fe& a) { return add(a, a); }
    // lazy variants of the unsaturated field layer (fp29.cuh); every value is fully reduced here
    BP_HD static fe add_l(const fe& a, const fe& b) { return add(a, b); }
    BP_HD static fe sub_l(const fe& a, const fe& b) { return sub(a, b); }
    BP_HD static fe dbl_l(const fe& a) { return add(a, a); }
    BP_HD static fe norm(const fe& a) { return a; }
    BP_HD static fe mul3(const fe& a) { return add(dbl(a), a); }
    BP_HD static fe mul_small(const fe& a, int k) {   // k in {0..8}, by additions
        fe r = zero();
        fe p = a;
        for (int bit = 0; bit < 4; bit++) {
            if ((k >> bit) & 1) r = add(r, p);
            p = dbl(p);
        }
        return r;
    }

    // Montgomery product a*b*2^-256 mod m.
    // Running total T = X + 2^32*Y. Each round adds a*b_k and q*m (even-indexed limbs of
    // a / m into the aligned array, odd-indexed into the offset array), then divides by
    // 2^32 by swapping roles: X' = Y + X[1] (carry handed to Y'), Y' = X >> 64.
    // Bounds: T < 2^33*m inside a round, so Y < 2m < 2^257 and X < 2^259: 9 limbs each.
    // Same CIOS for moduli with a sparse high part, m = m_low + 2^A (+|-) 2^B with m_low < 2^(32*RED_LOW) -- all five
    // moduli of this repository (secq256k1's base field is 2^256 - 2^129 + m_low, four limbs). q*m then needs only
    // RED_LOW wide multiplies per round (the two-pass IMAD.WIDE.X is the bottleneck pipe); the
    // rest, D = q*2^A (+|-) q*2^B >= 0, is built with shifts and one borrow chain and added on the single carry
    // chain that X needs anyway to ripple the carry of its multiplies. All of D goes to X, so X grows to 10 limbs
    // (X < 2^258 + 2^288); Y' = X >> 64 still fits 8 limbs before a_odd*b_k is added.
    BP_HD static fe mul_sparse(const fe& a, const fe& b) {
        constexpr int LC = M::RED_LOW, IA = M::RED_A / 32, SA = M::RED_A % 32;
        constexpr int IB = M::RED_B >= 0 ? M::RED_B / 32 : 99, SB = M::RED_B >= 0 ? M::RED_B % 32 : 0;
        static_assert(LC >= 2 && LC % 2 == 0 && LC <= 6, "RED_LOW");
        static_assert(M::RED_BSIGN == 0 || (IB >= LC && IA >= IB + 2), "shifted terms must sit above the multiplies and not overlap");
        static_assert(IA >= LC && IA <= 8, "RED_A");
        uint32_t X[10], Y[10];
#pragma unroll
        for (int i = 0; i < 10; i++) { X[i] = 0; Y[i] = 0; }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t bk = b.v[k];
            uint32_t Xn[10], Yn[10];
            Xn[0] = add_cc(Y[0], X[1]);
#pragma unroll
            for (int j = 1; j < 9; j++) Xn[j] = Y[j];
            Xn[9] = 0;
#pragma unroll
            for (int j = 0; j < 8; j += 2) wmadc_to_cc(Yn[j], Yn[j + 1], a.v[j + 1], bk, X[j + 2], X[j + 3]);
            Yn[8] = addc(0u, 0u);
            wmad_cc(Xn[0], Xn[1], a.v[0], bk);
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], a.v[j], bk);
            Xn[8] = addc(Xn[8], 0u);
            const uint32_t q = Xn[0] * M::INV32;
            // D = q*2^A (+|-) q*2^B, limbs LC..9 (compile-time zero elsewhere)
            uint32_t D[10];
#pragma unroll
            for (int j = 0; j < 10; j++) D[j] = 0;
            const uint32_t qa_lo = q << SA, qa_hi = SA ? (q >> ((32 - SA) & 31)) : 0u;
            if (M::RED_BSIGN < 0) {
                const uint32_t qb_lo = q << SB, qb_hi = SB ? (q >> ((32 - SB) & 31)) : 0u;
                D[IB] = sub_cc(0u, qb_lo);
                D[IB + 1] = subc_cc(0u, qb_hi);
                const uint32_t ext = subc(0u, 0u);            // 0, or 0xFFFFFFFF: sign extension of -q*2^B
#pragma unroll
                for (int j = IB + 2; j < IA; j++) D[j] = ext;
                D[IA] = add_cc(qa_lo, ext);
                if (IA + 1 < 10) D[IA + 1] = addc(qa_hi, ext);   // limbs above cancel to zero (D >= 0)
            } else {
                if (M::RED_BSIGN > 0) {
                    D[IB] = q << SB;
                    D[IB + 1] = SB ? (q >> ((32 - SB) & 31)) : 0u;
                }
                D[IA] = qa_lo;
                if (IA + 1 < 10) D[IA + 1] = qa_hi;
            }
            // Y' += q * (odd limbs of m_low), carry rippled to the top
            wmad_cc(Yn[0], Yn[1], q, M::m(1));
#pragma unroll
            for (int j = 2; j < LC; j += 2) wmadc_cc(Yn[j], Yn[j + 1], q, M::m(j + 1));
#pragma unroll
            for (int j = LC; j < 8; j++) Yn[j] = addc_cc(Yn[j], 0u);
            Yn[8] = addc(Yn[8], 0u);
            // X' += q * (even limbs of m_low) + D on one chain
            wmad_cc(Xn[0], Xn[1], q, M::m(0));
#pragma unroll
            for (int j = 2; j < LC; j += 2) wmadc_cc(Xn[j], Xn[j + 1], q, M::m(j));
#pragma unroll
            for (int j = LC; j < 9; j++) Xn[j] = addc_cc(Xn[j], D[j]);
            Xn[9] = addc(Xn[9], D[9]);
#pragma unroll
            for (int j = 0; j < 10; j++) { X[j] = Xn[j]; }
#pragma unroll
            for (int j = 0; j < 9; j++) { Y[j] = Yn[j]; }
            Y[9] = 0;
        }
        uint32_t r[9];
        r[0] = add_cc(Y[0], X[1]);
#pragma unroll
        for (int j = 1; j < 8; j++) r[j] = addc_cc(Y[j], X[j + 1]);
        r[8] = addc(Y[8], X[9]);
        fe o;
        final_sub(o, r, r[8]);
        return o;
    }

    // Measured on B200 (round 1, gpurun_out/microbench_{sparse,generic}.json, msm_phases_sparse.log): mul_sparse has
    // 96 instead of 128 two-pass IMAD.WIDE.X per product but ~145 instead of ~50 carry-chained IADD3.X, and is NOT
    // faster -- Fp::mul 64.4 vs 66.8 G/s, mixed addition 5.3 vs 6.2 G/s, msm_accumulate 41.3 vs 36.0 ms at 2^24: the
    // carry-chained adds do not overlap with the carry-chained multiplies the way plain ALU work does. It stays as a
    // tested alternative (BP_SPARSE_MONT selects it; tests/test_hostmath.py checks it against the oracle on all five
    // moduli); the default is the generic CIOS.
    BP_HD static fe mul(const fe& a, const fe& b) {
#if defined(BP_SPARSE_MONT)
        if (M::RED_SPARSE) return mul_sparse(a, b);
#endif
        return mul_generic(a, b);
    }

    BP_HD static fe mul_generic(const fe& a, const fe& b) {
        uint32_t X[10], Y[10];
#pragma unroll
        for (int i = 0; i < 10; i++) { X[i] = 0; Y[i] = 0; }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t bk = b.v[k];
            uint32_t Xn[10], Yn[10];
            Xn[0] = add_cc(Y[0], X[1]);
#pragma unroll
            for (int j = 1; j < 9; j++) Xn[j] = Y[j];
#pragma unroll
            for (int j = 0; j < 8; j += 2) wmadc_to_cc(Yn[j], Yn[j + 1], a.v[j + 1], bk, X[j + 2], X[j + 3]);
            Yn[8] = addc(0u, 0u);
            wmad_cc(Xn[0], Xn[1], a.v[0], bk);
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], a.v[j], bk);
            Xn[8] = addc(Xn[8], 0u);
            const uint32_t q = Xn[0] * M::INV32;
            wmad_cc(Yn[0], Yn[1], q, M::m(1));
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Yn[j], Yn[j + 1], q, M::m(j + 1));
            Yn[8] = addc(Yn[8], 0u);
            wmad_cc(Xn[0], Xn[1], q, M::m(0));
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], q, M::m(j));
            Xn[8] = addc(Xn[8], 0u);
#pragma unroll
            for (int j = 0; j < 9; j++) { X[j] = Xn[j]; Y[j] = Yn[j]; }
            X[9] = 0; Y[9] = 0;
        }
        uint32_t r[9];
        r[0] = add_cc(Y[0], X[1]);
#pragma unroll
        for (int j = 1; j < 8; j++) r[j] = addc_cc(Y[j], X[j + 1]);
        r[8] = addc(Y[8], 0u);
        fe o;
        final_sub(o, r, r[8]);
        return o;
    }

    // r = t - m if t >= m for a 9-limb t < 3m: the result keeps its 9th limb (t9 < 2m afterwards)
    BP_HD static void cond_sub9(uint32_t* t) {
        uint32_t s[9];
        s[0] = sub_cc(t[0], M::m(0));
#pragma unroll
        for (int i = 1; i < 8; i++) s[i] = subc_cc(t[i], M::m(i));
        s[8] = subc_cc(t[8], 0u);
        uint32_t borrow = subc(0u, 0u);    // 0xFFFFFFFF iff t < m
#pragma unroll
        for (int i = 0; i < 9; i++) t[i] = borrow ? t[i] : s[i];
    }

    // a*b + c*d (SUB = false) or a*b - c*d (SUB = true), times 2^-256 mod m: two products, ONE Montgomery reduction.
    // The same even/odd-column CIOS as mul_generic with a second pair of product chains per round: 128 + 64 wide
    // multiplies + 8 quotients instead of 2 x (128 + 8). Every XYZZ addition and doubling ends in one of these
    // (Y3 = R*(Q - X3) - Y1*PPP). The difference is formed as a*b + c*(m - d), so all addends stay non-negative:
    // T' <= (T + 3*(2^32 - 1)*m) / 2^32 keeps T < 3m, i.e. X, Y < 2^260 (9 limbs each), and the result needs two
    // conditional subtractions instead of one.
    template <bool SUB>
    BP_HD static fe mul2(const fe& a, const fe& b, const fe& c, const fe& d_in) {
        fe d = d_in;
        if (SUB) {                                   // d <- m - d  (in (0, m]; d = 0 gives m = 0 mod m)
            d.v[0] = sub_cc(M::m(0), d_in.v[0]);
#pragma unroll
            for (int i = 1; i < 7; i++) d.v[i] = subc_cc(M::m(i), d_in.v[i]);
            d.v[7] = subc(M::m(7), d_in.v[7]);
        }
        uint32_t X[10], Y[10];
#pragma unroll
        for (int i = 0; i < 10; i++) { X[i] = 0; Y[i] = 0; }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t bk = b.v[k], dk = d.v[k];
            uint32_t Xn[10], Yn[10];
            Xn[0] = add_cc(Y[0], X[1]);
#pragma unroll
            for (int j = 1; j < 9; j++) Xn[j] = Y[j];
#pragma unroll
            for (int j = 0; j < 8; j += 2) wmadc_to_cc(Yn[j], Yn[j + 1], a.v[j + 1], bk, X[j + 2], X[j + 3]);
            Yn[8] = addc(0u, 0u);
            wmad_cc(Xn[0], Xn[1], a.v[0], bk);
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], a.v[j], bk);
            Xn[8] = addc(Xn[8], 0u);
            wmad_cc(Yn[0], Yn[1], c.v[1], dk);
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Yn[j], Yn[j + 1], c.v[j + 1], dk);
            Yn[8] = addc(Yn[8], 0u);
            wmad_cc(Xn[0], Xn[1], c.v[0], dk);
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], c.v[j], dk);
            Xn[8] = addc(Xn[8], 0u);
            const uint32_t q = Xn[0] * M::INV32;
            wmad_cc(Yn[0], Yn[1], q, M::m(1));
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Yn[j], Yn[j + 1], q, M::m(j + 1));
            Yn[8] = addc(Yn[8], 0u);
            wmad_cc(Xn[0], Xn[1], q, M::m(0));
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], q, M::m(j));
            Xn[8] = addc(Xn[8], 0u);
#pragma unroll
            for (int j = 0; j < 9; j++) { X[j] = Xn[j]; Y[j] = Yn[j]; }
            X[9] = 0; Y[9] = 0;
        }
        uint32_t r[9];
        r[0] = add_cc(Y[0], X[1]);
#pragma unroll
        for (int j = 1; j < 8; j++) r[j] = addc_cc(Y[j], X[j + 1]);
        r[8] = addc(Y[8], 0u);
        cond_sub9(r);
        fe o;
        final_sub(o, r, r[8]);
        return o;
    }
#if defined(BP_NO_MUL2)
    BP_HD static fe mul_add(const fe& a, const fe& b, const fe& c, const fe& d) { return add(mul(a, b), mul(c, d)); }
    BP_HD static fe mul_sub(const fe& a, const fe& b, const fe& c, const fe& d) { return sub(mul(a, b), mul(c, d)); }
#else
    BP_HD static fe mul_add(const fe& a, const fe& b, const fe& c, const fe& d) { return mul2<false>(a, b, c, d); }
    BP_HD static fe mul_sub(const fe& a, const fe& b, const fe& c, const fe& d) { return mul2<true>(a, b, c, d); }
#endif

    // r (9 limbs) = (L + Q*m) / 2^256 for the 8-limb L: the eight quotient rounds of the CIOS alone (X aligned, Y offset by
    // one limb, as in mul_generic); r <= m. 64 wide + 8 quotient multiplies.
    BP_HD static void redc8(const uint32_t* L, uint32_t* r) {
        uint32_t X[10], Y[10];
#pragma unroll
        for (int i = 0; i < 8; i++) { X[i] = L[i]; Y[i] = 0; }
        X[8] = 0; X[9] = 0; Y[8] = 0; Y[9] = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t Xn[10], Yn[10];
            if (k == 0) {                            // nothing to shift in yet: q from X[0]
#pragma unroll
                for (int j = 0; j < 9; j++) { Xn[j] = X[j]; Yn[j] = 0; }
            } else {
                Xn[0] = add_cc(Y[0], X[1]);
#pragma unroll
                for (int j = 1; j < 9; j++) Xn[j] = Y[j];
            }
            const uint32_t q = Xn[0] * M::INV32;
            if (k == 0) {
                wmad_cc(Yn[0], Yn[1], q, M::m(1));
#pragma unroll
                for (int j = 2; j < 8; j += 2) wmadc_cc(Yn[j], Yn[j + 1], q, M::m(j + 1));
                Yn[8] = addc(0u, 0u);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j += 2) wmadc_to_cc(Yn[j], Yn[j + 1], q, M::m(j + 1), X[j + 2], X[j + 3]);
                Yn[8] = addc(0u, 0u);
            }
            wmad_cc(Xn[0], Xn[1], q, M::m(0));
#pragma unroll
            for (int j = 2; j < 8; j += 2) wmadc_cc(Xn[j], Xn[j + 1], q, M::m(j));
            Xn[8] = addc(Xn[8], 0u);
#pragma unroll
            for (int j = 0; j < 9; j++) { X[j] = Xn[j]; Y[j] = Yn[j]; }
            X[9] = 0; Y[9] = 0;
        }
        r[0] = add_cc(Y[0], X[1]);
#pragma unroll
        for (int j = 1; j < 8; j++) r[j] = addc_cc(Y[j], X[j + 1]);
        r[8] = addc(Y[8], 0u);
    }

    // a^2 * 2^-256 mod m with 36 + 64 wide multiplies instead of 128: the 28 off-diagonal products a_i*a_j (i < j) are
    // summed once (even/odd columns as above, one row per a_i), doubled by a one-bit funnel shift, the 8 squares a_i^2
    // are added on one chain, and the 512-bit square is reduced by the quotient rounds of the CIOS alone (the high half
    // joins at the end: (L + Q*m)/2^256 + H < 2m).
    BP_HD static fe sqr_sos(const fe& a) {
        // off-diagonal sum S = sum_{i<j} a_i a_j 2^(32(i+j)) < 2^511: aligned columns E[p], offset columns O[p] (value 2^32 * O)
        uint32_t E[16], O[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
        for (int i = 0; i < 7; i++) {
            // row i: products at positions p = i + j, j = i+1..7. Odd p go to O[p-1], O[p]; even p to E[p], E[p+1].
            // Each parity is one carry chain over contiguous limbs; its carry-out lands on the limb above the last product,
            // which earlier rows have touched at most as their own carry limb -- by S < 2^511 the ripple ends there.
            const uint32_t ai = a.v[i];
            // first position of the row, i + (i+1) = 2i+1, is odd
            {
                bool first = true;
                int last = -1;
#pragma unroll
                for (int j = i + 1; j < 8; j += 2) {           // odd positions p = i + j (j - i odd)
                    const int p = i + j;
                    if (first) { wmad_cc(O[p - 1], O[p], ai, a.v[j]); first = false; }
                    else wmadc_cc(O[p - 1], O[p], ai, a.v[j]);
                    last = p;
                }
                if (last >= 0) {
                    O[last + 1] = addc_cc(O[last + 1], 0u);
                    O[last + 2] = addc(O[last + 2], 0u);
                }
            }
            {
                bool first = true;
                int last = -1;
#pragma unroll
                for (int j = i + 2; j < 8; j += 2) {           // even positions
                    const int p = i + j;
                    if (first) { wmad_cc(E[p], E[p + 1], ai, a.v[j]); first = false; }
                    else wmadc_cc(E[p], E[p + 1], ai, a.v[j]);
                    last = p;
                }
                if (last >= 0) {
                    E[last + 2] = addc_cc(E[last + 2], 0u);
                    if (last + 3 < 16) E[last + 3] = addc(E[last + 3], 0u);
                }
            }
        }
        // S = E + 2^32 * O ; T = 2*S + sum a_i^2 2^(64 i)
        uint32_t S[16];
        S[0] = E[0];
        S[1] = add_cc(E[1], O[0]);
#pragma unroll
        for (int i = 2; i < 15; i++) S[i] = addc_cc(E[i], O[i - 1]);
        S[15] = addc(E[15], O[14]);
        uint32_t T[16];
        T[0] = S[0] << 1;
#pragma unroll
        for (int i = 1; i < 16; i++) T[i] = (S[i] << 1) | (S[i - 1] >> 31);
        wmad_cc(T[0], T[1], a.v[0], a.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) wmadc_cc(T[2 * i], T[2 * i + 1], a.v[i], a.v[i]);
        // Montgomery-reduce the low half (redc8), then the high half H = T[8..15] joins
        uint32_t r[9];
        redc8(T, r);
        r[0] = add_cc(r[0], T[8]);
#pragma unroll
        for (int j = 1; j < 8; j++) r[j] = addc_cc(r[j], T[8 + j]);
        r[8] = addc(r[8], 0u);
        fe o;
        final_sub(o, r, r[8]);
        return o;
    }

    BP_HD static fe sqr(const fe& a) {
#if defined(BP_NO_SQR)
        return mul(a, a);
#else
        return sqr_sos(a);
#endif
    }

    BP_HD static fe one() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = M::one(i);
        return r;
    }
    BP_HD static fe r2() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = M::r2(i);
        return r;
    }
    BP_HD static fe to_mont(const fe& a) { return mul(a, r2()); }
    BP_HD static fe from_mont(const fe& a) {         // a * 2^-256: the reduction without a product
        uint32_t r[9];
        redc8(a.v, r);
        fe o;
        final_sub(o, r, r[8]);
        return o;
    }
    BP_HD static fe from_u32(uint32_t x) {
        fe o = zero();
        o.v[0] = x;
        return to_mont(o);
    }

    // a^e, e as 8 little-endian limbs: fixed 4-bit windows, MSB first -- 14 multiplications for the table, then four
    // squarings and at most one multiplication per nibble (252 S + <= 77 M; plain square-and-multiply needs a
    // multiplication per set bit: ~190 for the Fermat exponents of the secq256k1 fields, 250 for 2^255 - 19).
    // The table lives in local memory (dynamic index); the chain is latency, not bandwidth.
    BP_HD_NOINL static fe pow(const fe& a, const uint32_t* e) {
        fe tbl[16];
        tbl[0] = one();
        tbl[1] = a;
        for (int i = 2; i < 16; i++) tbl[i] = mul(tbl[i - 1], a);
        fe r = one();
        bool started = false;
        for (int i = 7; i >= 0; i--) {
            for (int nib = 7; nib >= 0; nib--) {
                const uint32_t d = (e[i] >> (4 * nib)) & 15u;
                if (started) {
                    r = sqr(sqr(sqr(sqr(r))));
                    if (d) r = mul(r, tbl[d]);
                } else if (d) {
                    r = tbl[d];
                    started = true;
                }
            }
        }
        return r;
    }

    // Fermat inverse a^(m-2); inv(0) = 0
    BP_HD_NOINL static fe inv(const fe& a) {
        uint32_t e[8];
        e[0] = M::m(0) - 2u;   // all five moduli have low limb >= 2: no borrow
        for (int i = 1; i < 8; i++) e[i] = M::m(i);
        return pow(a, e);
    }
};

#if defined(__CUDACC__)
__device__ __forceinline__ fe ld_fe(const void* p) {   // read-only path (inputs never written by the same kernel)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ fe ld_fe_rw(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fe(void* p, const fe& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
#endif

}  // namespace bp

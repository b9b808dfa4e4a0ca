// Short-Weierstrass group law for the bucket / fold kernels (secq256k1: a = 0; zorro: a = 6).
//
// Replaces ark-ec `short_weierstrass::{Affine,Projective}` arithmetic that the reference
// reaches through `G::Group::msm`, `mul_bigint`, `into_affine` (SURVEY.md 8(a) rows a1, a11;
// src/generators.rs:39-44, src/inner_product_proof.rs:104-156). Only the *values* matter for
// parity (results are canonicalised to affine before they are hashed or serialised), so the
// coordinates are chosen for the GPU: XYZZ (x = X/ZZ, y = Y/ZZZ) for accumulation --
// mixed add 8M+2S, full add 12M+2S -- and affine (x,y) Montgomery pairs in HBM (64 B).
//
// Identity encodings: XYZZ with ZZ == 0; affine (0,0) (never on a curve with b != 0).
#pragma once
#include "fp.cuh"

namespace bp {

template <class EL> struct affine_t { EL x, y; };
template <class EL> struct xyzz_t { EL x, y, zz, zzz; };
using affine = affine_t<fe>;   // 64 B in HBM (fe is 16-byte aligned)
using xyzz = xyzz_t<fe>;       // 128 B in HBM

template <class C, class F_ = Fp<typename C::Fq>>
struct SW {
    static constexpr bool IS_TE = false;
    using F = F_;
    using el = typename F::el;
    using aff = affine_t<el>;
    using ext = xyzz_t<el>;

    BP_HD static bool is_identity(const aff& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
    BP_HD static bool is_identity(const ext& p) { return F::is_zero(p.zz); }
    BP_HD static ext identity() {
        ext r;
        r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    BP_HD static aff affine_identity() {
        aff r;
        r.x = F::zero(); r.y = F::zero();
        return r;
    }
    BP_HD static ext from_affine(const aff& p) {
        ext r;
        if (is_identity(p)) return identity();
        r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
        return r;
    }
    BP_HD static aff neg(const aff& p) {
        aff r;
        r.x = p.x; r.y = F::neg(p.y);
        return r;
    }
    BP_HD static ext neg(const ext& p) {
        ext r = p;
        r.y = F::neg(p.y);
        return r;
    }
    BP_HD static el mul_a(const el& t) { return F::mul_small(t, C::A_SMALL); }

    // 2*(x,y) -> XYZZ   [mdbl-2008-s-1]
    BP_HD static ext dbl_affine(const aff& p) {
        if (is_identity(p) || F::is_zero(p.y)) return identity();
        ext r;
        el U = F::dbl(p.y);
        el V = F::sqr(U);
        el W = F::mul(U, V);
        el S = F::mul(p.x, V);
        el M = F::mul3(F::sqr(p.x));
        if (C::A_SMALL != 0) M = F::add(M, F::from_u32(C::A_SMALL));
        r.x = F::sub(F::sqr(M), F::dbl_l(S));
        r.y = F::mul_sub(M, F::sub(S, r.x), W, p.y);
        r.zz = V;
        r.zzz = W;
        return r;
    }

    // 2*P, XYZZ   [dbl-2008-s-1]
    BP_HD static ext dbl(const ext& p) {
        if (is_identity(p) || F::is_zero(p.y)) return identity();
        ext r;
        el U = F::dbl(p.y);
        el V = F::sqr(U);
        el W = F::mul(U, V);
        el S = F::mul(p.x, V);
        el M = F::mul3(F::sqr(p.x));
        if (C::A_SMALL != 0) M = F::add(M, mul_a(F::sqr(p.zz)));
        r.x = F::sub(F::sqr(M), F::dbl_l(S));
        r.y = F::mul_sub(M, F::sub(S, r.x), W, p.y);
        r.zz = F::mul(V, p.zz);
        r.zzz = F::mul(W, p.zzz);
        return r;
    }

    // The doubling branches of madd / add (P == Q) are taken for adversarial inputs only. Moving them out of line
    // (__noinline__) was measured: +3 % / +17 % on dependent chains of madd / add (tools/microbench7.cu), but the MSM kernels
    // got slower (accumulate 31.95 -> 32.8 ms, reduce 3.77 -> 4.07 ms at 2^24: call ABI, stack frames, spills at the
    // 128-register cap), so they stay inline.
    // acc += (x2,y2)   [madd-2008-s], all special cases handled
    BP_HD static void madd(ext& acc, const aff& q) {
        if (is_identity(q)) return;
        if (is_identity(acc)) { acc = from_affine(q); return; }
        el U2 = F::mul(q.x, acc.zz);
        el S2 = F::mul(q.y, acc.zzz);
        el P = F::sub(U2, acc.x);
        el R = F::sub(S2, acc.y);
        if (F::is_zero(P)) {
            if (F::is_zero(R)) acc = dbl_affine(q);
            else acc = identity();
            return;
        }
        el PP = F::sqr(P);
        el PPP = F::mul(P, PP);
        el Q = F::mul(acc.x, PP);
        el X3 = F::sub(F::sub_l(F::sqr(R), PPP), F::dbl_l(Q));
        el Y3 = F::mul_sub(R, F::sub(Q, X3), acc.y, PPP);
        acc.x = X3;
        acc.y = Y3;
        acc.zz = F::mul(acc.zz, PP);
        acc.zzz = F::mul(acc.zzz, PPP);
    }

    // acc += q, both XYZZ   [add-2008-s]
    BP_HD static void add(ext& acc, const ext& q) {
        if (is_identity(q)) return;
        if (is_identity(acc)) { acc = q; return; }
        el U1 = F::mul(acc.x, q.zz);
        el U2 = F::mul(q.x, acc.zz);
        el S1 = F::mul(acc.y, q.zzz);
        el S2 = F::mul(q.y, acc.zzz);
        el P = F::sub(U2, U1);
        el R = F::sub(S2, S1);
        if (F::is_zero(P)) {
            if (F::is_zero(R)) acc = dbl(acc);
            else acc = identity();
            return;
        }
        el PP = F::sqr(P);
        el PPP = F::mul(P, PP);
        el Q = F::mul(U1, PP);
        el X3 = F::sub(F::sub_l(F::sqr(R), PPP), F::dbl_l(Q));
        el Y3 = F::mul_sub(R, F::sub(Q, X3), S1, PPP);
        acc.x = X3;
        acc.y = Y3;
        acc.zz = F::mul(F::mul(acc.zz, q.zz), PP);
        acc.zzz = F::mul(F::mul(acc.zzz, q.zzz), PPP);
    }

    // XYZZ -> aff with one field inversion (x = X/ZZ, y = Y/ZZZ)
    BP_HD_NOINL static aff to_affine(const ext& p) {
        if (is_identity(p)) return affine_identity();
        // 1/ZZZ, then 1/ZZ = (1/ZZZ)^2 * ZZ^2 ... simpler: invert ZZ*ZZZ once
        el t = F::mul(p.zz, p.zzz);
        el ti = F::inv(t);
        aff r;
        r.x = F::mul(p.x, F::mul(ti, p.zzz));
        r.y = F::mul(p.y, F::mul(ti, p.zz));
        return r;
    }

    // k*P for a small non-negative k (double-and-add, MSB first); used by the bucket reduction
    BP_HD_NOINL static ext mul_u32(const ext& p, uint32_t k) {
        ext acc = identity();
        int top = 31;
        while (top >= 0 && !((k >> top) & 1u)) top--;
        for (int bit = top; bit >= 0; bit--) {
            acc = dbl(acc);
            if ((k >> bit) & 1u) add(acc, p);
        }
        return acc;
    }

    // s*P for a canonical (non-Montgomery) 256-bit scalar given as 8 LE limbs
    BP_HD_NOINL static ext mul_scalar(const aff& p, const uint32_t* s) {
        ext acc = identity();
        for (int i = 7; i >= 0; i--) {
            for (int bit = 31; bit >= 0; bit--) {
                acc = dbl(acc);
                if ((s[i] >> bit) & 1u) madd(acc, p);
            }
        }
        return acc;
    }

    BP_HD static bool on_curve(const aff& p) {
        if (is_identity(p)) return true;
        el b = F::template curve_b<C>();
        el rhs = F::add(F::mul(F::sqr(p.x), p.x), b);
        if (C::A_SMALL != 0) rhs = F::add(rhs, mul_a(p.x));
        return F::eq(F::sqr(p.y), rhs);
    }
};

// ---- twisted Edwards, a = -1 (curve25519 as instantiated by tests/r1cs_curve25519.rs) -----------------
// Extended coordinates (X:Y:Z:T), x = X/Z, y = Y/Z, T = XY/Z, stored in the same 128-byte slot as
// XYZZ (fields x, y, zz, zzz = X, Y, Z, T). The unified addition law (add-2008-hwcd-3) is complete
// for a = -1 (a square) and d a non-square, so there are no exceptional cases; "all zero" (Z = 0)
// is accepted as an additional encoding of the identity so zero-initialised buckets work unchanged.
// ABI: aff (x,y); the identity is (0,1) and, on input, also (0,0).
template <class C, class F_ = Fp<typename C::Fq>>
struct TE {
    static constexpr bool IS_TE = true;
    using F = F_;
    using el = typename F::el;
    using aff = affine_t<el>;
    using ext = xyzz_t<el>;
    BP_HD static el d2() { return F::template te_d2<C>(); }
    BP_HD static el dcoef() { return F::template curve_b<C>(); }
    BP_HD static bool is_identity(const aff& p) { return F::is_zero(p.x) && (F::is_zero(p.y) || F::eq(p.y, F::one())); }
    BP_HD static bool is_identity(const ext& p) { return F::is_zero(p.zz) || (F::is_zero(p.x) && F::eq(p.y, p.zz)); }
    BP_HD static ext identity() {
        ext r;
        r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    BP_HD static aff affine_identity() {
        aff r;
        r.x = F::zero(); r.y = F::zero();
        return r;
    }
    BP_HD static ext from_affine(const aff& p) {
        if (is_identity(p)) return identity();
        ext r;
        r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::mul(p.x, p.y);
        return r;
    }
    BP_HD static aff neg(const aff& p) {
        aff r;
        r.x = F::neg(p.x); r.y = p.y;
        return r;
    }
    BP_HD static ext neg(const ext& p) {
        ext r = p;
        r.x = F::neg(p.x); r.zzz = F::neg(p.zzz);
        return r;
    }
    // dbl-2008-hwcd with a = -1: 4M + 4S
    BP_HD static ext dbl(const ext& p) {
        if (F::is_zero(p.zz)) return identity();
        el A = F::sqr(p.x), B = F::sqr(p.y), Cc = F::dbl_l(F::sqr(p.zz));
        el D = F::neg(A);
        el E = F::sub(F::sub_l(F::sqr(F::add(p.x, p.y)), A), B);
        el Gl = F::add_l(D, B);
        el G = F::norm(Gl), Fv = F::sub(Gl, Cc), H = F::sub(D, B);
        ext r;
        r.x = F::mul(E, Fv); r.y = F::mul(G, H); r.zzz = F::mul(E, H); r.zz = F::mul(Fv, G);
        return r;
    }
    BP_HD static ext dbl_affine(const aff& p) { return dbl(from_affine(p)); }
    // acc += q (both extended): add-2008-hwcd-3, 8M + 1 (2d)
    BP_HD static void add(ext& acc, const ext& q) {
        if (F::is_zero(q.zz)) return;
        if (F::is_zero(acc.zz)) { acc = q; return; }
        el A = F::mul(F::sub(acc.y, acc.x), F::sub(q.y, q.x));
        el B = F::mul(F::add(acc.y, acc.x), F::add(q.y, q.x));
        el Cc = F::mul(F::mul(acc.zzz, d2()), q.zzz);
        el D = F::dbl_l(F::mul(acc.zz, q.zz));
        el E = F::sub(B, A), Fv = F::sub(D, Cc), G = F::add(D, Cc), H = F::add(B, A);
        acc.x = F::mul(E, Fv); acc.y = F::mul(G, H); acc.zzz = F::mul(E, H); acc.zz = F::mul(Fv, G);
    }
    // acc += aff q: 9M (T2 = x2*y2 computed on the fly)
    BP_HD static void madd(ext& acc, const aff& q) {
        if (is_identity(q)) return;
        if (F::is_zero(acc.zz)) { acc = from_affine(q); return; }
        el A = F::mul(F::sub(acc.y, acc.x), F::sub(q.y, q.x));
        el B = F::mul(F::add(acc.y, acc.x), F::add(q.y, q.x));
        el Cc = F::mul(F::mul(acc.zzz, d2()), F::mul(q.x, q.y));
        el D = F::dbl_l(acc.zz);
        el E = F::sub(B, A), Fv = F::sub(D, Cc), G = F::add(D, Cc), H = F::add(B, A);
        acc.x = F::mul(E, Fv); acc.y = F::mul(G, H); acc.zzz = F::mul(E, H); acc.zz = F::mul(Fv, G);
    }
    BP_HD_NOINL static aff to_affine(const ext& p) {
        if (is_identity(p)) return affine_identity();
        el zi = F::inv(p.zz);
        aff r;
        r.x = F::mul(p.x, zi);
        r.y = F::mul(p.y, zi);
        return r;
    }
    BP_HD_NOINL static ext mul_u32(const ext& p, uint32_t k) {
        ext acc = identity();
        int top = 31;
        while (top >= 0 && !((k >> top) & 1u)) top--;
        for (int bit = top; bit >= 0; bit--) {
            acc = dbl(acc);
            if ((k >> bit) & 1u) add(acc, p);
        }
        return acc;
    }
    BP_HD_NOINL static ext mul_scalar(const aff& p, const uint32_t* s) {
        ext acc = identity();
        for (int i = 7; i >= 0; i--) {
            for (int bit = 31; bit >= 0; bit--) {
                acc = dbl(acc);
                if ((s[i] >> bit) & 1u) madd(acc, p);
            }
        }
        return acc;
    }
    BP_HD static bool on_curve(const aff& p) {
        if (is_identity(p)) return true;
        el x2 = F::sqr(p.x), y2 = F::sqr(p.y);
        el lhs = F::sub(y2, x2);
        el rhs = F::add(F::one(), F::mul(dcoef(), F::mul(x2, y2)));
        return F::eq(lhs, rhs);
    }
};

// group law selected by the curve descriptor (consts.cuh: KIND 0 = short Weierstrass, 1 = twisted Edwards)
template <class C, class F, int KIND> struct GroupLawSel { using type = SW<C, F>; };
template <class C, class F> struct GroupLawSel<C, F, 1> { using type = TE<C, F>; };
template <class C, class F = Fp<typename C::Fq>> using GroupLaw = typename GroupLawSel<C, F, C::KIND>::type;

#if defined(__CUDACC__)
__device__ __forceinline__ affine ld_affine(const affine* p) {
    affine r;
    r.x = ld_fe(&p->x);
    r.y = ld_fe(&p->y);
    return r;
}
__device__ __forceinline__ void st_xyzz(xyzz* p, const xyzz& v) {
    st_fe(&p->x, v.x); st_fe(&p->y, v.y); st_fe(&p->zz, v.zz); st_fe(&p->zzz, v.zzz);
}
__device__ __forceinline__ xyzz ld_xyzz(const xyzz* p) {
    xyzz r;
    r.x = ld_fe_rw(&p->x); r.y = ld_fe_rw(&p->y); r.zz = ld_fe_rw(&p->zz); r.zzz = ld_fe_rw(&p->zzz);
    return r;
}
#endif

}  // namespace bp

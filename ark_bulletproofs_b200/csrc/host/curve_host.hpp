// Host-side restatement of the arkworks value semantics the reference relies on
// (SURVEY.md App. A.3, A.5): `Fp::rand`, `Affine::rand`, canonical (de)serialisation, plus
// small host group operations (single scalar multiplications for Pedersen commitments and
// Q = w*B, src/generators.rs:39-44, src/r1cs/prover.rs:779).
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>
#include "../ec.cuh"
#include "fp_host.hpp"
#include "merlin.hpp"

namespace bp {

template <class C>
struct HostCurve {
    using Fq = HostFp<typename C::Fq>;
    using Fr = HostFp<typename C::Fr>;
    using E = GroupLaw<C, Fq>;
    static constexpr bool IS_TE = C::KIND == 1;
    // ark-serialize sizes: SW 256-bit fields need a 33rd byte for the two flag bits; TE (255-bit field) packs
    // its one flag into the top bit and has no flags in uncompressed form
    static constexpr int POINT_COMPRESSED = IS_TE ? 32 : 33, POINT_UNCOMPRESSED = IS_TE ? 64 : 65;

    // ---- ark-ff Fp::rand: 4 x next_u64, shave, accept iff < modulus; raw = Montgomery repr ----
    template <class M>
    static fe fp_rand(Rng& rng) {
        const int shave = 256 - M::BITS;
        const uint64_t mask = shave == 64 ? 0 : (~0ull >> shave);
        while (true) {
            uint64_t l[4];
            rng.draw_u64(l, 4);
            l[3] &= mask;
            if (!HostFp<M>::geq_m(l)) return HostFp<M>::put(l);
        }
    }
    static fe scalar_rand(Rng& rng) { return fp_rand<typename C::Fr>(rng); }
    // n consecutive Fr::rand draws (prover.rs:510-513,599-602). The word stream is sequential, so drawing
    // exactly the words still needed (4 per missing scalar) in bulk and parsing them in order consumes the
    // same words as n single draws, rejections included.
    static void scalar_rand_bulk(Rng& rng, fe* out, size_t n) {
        using M = typename C::Fr;
        const int shave = 256 - M::BITS;
        const uint64_t mask = shave == 64 ? 0 : (~0ull >> shave);
        std::vector<uint64_t> w;
        size_t done = 0;
        while (done < n) {
            size_t want = n - done;
            if (want > 4096) want = 4096;
            w.resize(4 * want);
            rng.draw_u64(w.data(), 4 * want);
            for (size_t k = 0; k < want; k++) {
                uint64_t* l = &w[4 * k];
                l[3] &= mask;
                if (!HostFp<M>::geq_m(l)) out[done++] = HostFp<M>::put(l);
            }
        }
    }

    // ---- canonical little-endian bytes -----------------------------------------------------
    template <class F>
    static void fe_to_bytes(const fe& a, uint8_t out[32]) {
        fe c = F::from_mont(a);
        memcpy(out, c.v, 32);
    }
    template <class F>
    static bool fe_from_bytes(const uint8_t in[32], fe& out) {   // false if >= modulus
        uint64_t l[4];
        memcpy(l, in, 32);
        if (F::geq_m(l)) return false;
        out = F::to_mont(F::put(l));
        return true;
    }
    static void scalar_to_bytes(const fe& a, uint8_t out[32]) { fe_to_bytes<Fr>(a, out); }
    static bool scalar_from_bytes(const uint8_t in[32], fe& out) { return fe_from_bytes<Fr>(in, out); }

    // v > -v as canonical integers (ark SWFlags::from_y_coordinate / TEFlags::from_x_coordinate)
    static bool y_is_larger(const fe& y) {
        fe yc = Fq::from_mont(y), nc = Fq::from_mont(Fq::neg(y));
        return Fq::canonical_lt(nc, yc);
    }
    static affine te_identity() { affine r; r.x = Fq::zero(); r.y = Fq::one(); return r; }
    static void point_uncompressed(const affine& p, uint8_t* out) {      // src/transcript.rs:75-79
        if constexpr (IS_TE) {
            affine q = E::is_identity(p) ? te_identity() : p;
            fe_to_bytes<Fq>(q.x, out);
            fe_to_bytes<Fq>(q.y, out + 32);
        } else {
            if (E::is_identity(p)) { memset(out, 0, 65); out[64] = 0x40; return; }
            fe_to_bytes<Fq>(p.x, out);
            fe_to_bytes<Fq>(p.y, out + 32);
            out[64] = y_is_larger(p.y) ? 0x80 : 0;
        }
    }
    static void point_compressed(const affine& p, uint8_t* out) {        // src/r1cs/proof.rs:74-78
        if constexpr (IS_TE) {
            affine q = E::is_identity(p) ? te_identity() : p;
            fe_to_bytes<Fq>(q.y, out);
            if (y_is_larger(q.x)) out[31] |= 0x80;
        } else {
            if (E::is_identity(p)) { memset(out, 0, 33); out[32] = 0x40; return; }
            fe_to_bytes<Fq>(p.x, out);
            out[32] = y_is_larger(p.y) ? 0x80 : 0;
        }
    }

    static fe curve_b() { fe b; for (int i = 0; i < 8; i++) b.v[i] = C::b(i); return b; }
    static affine generator() {
        affine g;
        for (int i = 0; i < 8; i++) { g.x.v[i] = C::gx(i); g.y.v[i] = C::gy(i); }
        return g;
    }
    static fe rhs(const fe& x) {
        fe r = Fq::add(Fq::mul(Fq::sqr(x), x), curve_b());
        if (C::A_SMALL != 0) r = Fq::add(r, Fq::mul_small(x, C::A_SMALL));
        return r;
    }

    // Tonelli-Shanks; returns false for non-residues
    struct TS { int s; uint32_t t[8]; uint32_t t1h[8]; fe z; uint32_t half[8]; };
    static const TS& ts_params() {
        static const TS ts = [] {
            TS r;
            uint32_t e[8];
            for (int i = 0; i < 8; i++) e[i] = C::Fq::m(i);
            e[0] -= 1;   // q - 1
            // (q-1)/2 for Euler's criterion
            for (int i = 0; i < 8; i++) r.half[i] = (e[i] >> 1) | (i < 7 ? (e[i + 1] << 31) : 0);
            r.s = 0;
            while (!(e[0] & 1)) {
                for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? (e[i + 1] << 31) : 0);
                r.s++;
            }
            memcpy(r.t, e, 32);
            // (t+1)/2 = (t >> 1) + 1 for odd t
            uint32_t h[8];
            for (int i = 0; i < 8; i++) h[i] = (e[i] >> 1) | (i < 7 ? (e[i + 1] << 31) : 0);
            uint64_t c = 1;
            for (int i = 0; i < 8; i++) { c += h[i]; h[i] = (uint32_t)c; c >>= 32; }
            memcpy(r.t1h, h, 32);
            // a non-residue z, then z^t
            fe one = Fq::one();
            for (uint32_t k = 2;; k++) {
                fe cand = Fq::from_u32(k);
                fe l = Fq::pow(cand, r.half);
                if (!Fq::eq(l, one) && !Fq::is_zero(l)) { r.z = Fq::pow(cand, r.t); break; }
            }
            return r;
        }();
        return ts;
    }
    static bool fq_sqrt(const fe& a, fe& out) {
        const TS& ts = ts_params();
        if (Fq::is_zero(a)) { out = a; return true; }
        fe one = Fq::one();
        if (!Fq::eq(Fq::pow(a, ts.half), one)) return false;
        int m = ts.s;
        fe c = ts.z, t = Fq::pow(a, ts.t), r = Fq::pow(a, ts.t1h);
        while (!Fq::eq(t, one)) {
            int i = 0;
            fe t2 = t;
            while (!Fq::eq(t2, one)) { t2 = Fq::sqr(t2); i++; }
            fe b = c;
            for (int k = 0; k < m - i - 1; k++) b = Fq::sqr(b);
            m = i;
            c = Fq::sqr(b);
            t = Fq::mul(t, c);
            r = Fq::mul(r, b);
        }
        out = r;
        return true;
    }

    // ark-ec get_point_from_x_unchecked(x, greatest)
    static bool point_from_x(const fe& x, bool greatest, affine& out) {
        fe y;
        if (!fq_sqrt(rhs(x), y)) return false;
        fe ny = Fq::neg(y);
        bool y_larger = y_is_larger(y);
        out.x = x;
        out.y = (greatest == y_larger) ? y : ny;
        return true;
    }

    // ark-ec TE get_point_from_y_unchecked(y, greatest): x^2 = (1 - y^2) / (a - d*y^2), a = -1
    static bool point_from_y(const fe& y, bool greatest, affine& out) {
        fe y2 = Fq::sqr(y);
        fe num = Fq::sub(Fq::one(), y2);
        fe den = Fq::sub(Fq::neg(Fq::one()), Fq::mul(curve_b(), y2));
        if (Fq::is_zero(den)) return false;
        fe x;
        if (!fq_sqrt(Fq::mul(num, Fq::inv(den)), x)) return false;
        fe nx = Fq::neg(x);
        bool x_larger = y_is_larger(x);
        out.x = (greatest == x_larger) ? x : nx;
        out.y = y;
        return true;
    }

    // ark-ec `Affine::rand`: SW: x = Fq::rand, greatest = top bit of next_u32, retry until on curve;
    // TE: the same with y drawn and x recovered, then mul_by_cofactor (src/generators.rs:63,99,115)
    static affine affine_rand(Rng& rng) {
        while (true) {
            fe c0 = fp_rand<typename C::Fq>(rng);
            bool greatest = (rng.next_u32() >> 31) & 1;
            affine p;
            if constexpr (IS_TE) {
                if (!point_from_y(c0, greatest, p)) continue;
                xyzz acc = E::from_affine(p);
                for (int k = 0; k < 3; k++) acc = E::dbl(acc);     // cofactor 8
                affine r = E::to_affine(acc);
                return E::is_identity(r) ? te_identity() : r;
            } else {
                if (point_from_x(c0, greatest, p)) return p;
            }
        }
    }

    // What ark-serialize's validation guarantees for every `G` the reference ever holds, applied to points that enter the
    // C ABI in the raw 64-byte form (statement commitments, caller-supplied generators, bp_proof_set_field): both
    // coordinates canonical (Montgomery residues < q), on the curve, and in the prime-order subgroup (TE, cofactor 8).
    static bool point_valid(const affine& p) {
        uint64_t l[4];
        memcpy(l, p.x.v, 32);
        if (Fq::geq_m(l)) return false;
        memcpy(l, p.y.v, 32);
        if (Fq::geq_m(l)) return false;
        if (E::is_identity(p)) return true;
        if (!E::on_curve(p)) return false;
        if constexpr (IS_TE) {
            uint32_t rl[8];
            for (int i = 0; i < 8; i++) rl[i] = C::Fr::m(i);
            if (!E::is_identity(E::mul_scalar(p, rl))) return false;
        }
        return true;
    }

    // deserialize_compressed with validation (on curve, and in the prime-order subgroup for TE)
    static bool point_from_compressed(const uint8_t* in, affine& out) {
        if constexpr (IS_TE) {
            uint8_t buf[32];
            memcpy(buf, in, 32);
            bool sign = (buf[31] & 0x80) != 0;
            buf[31] &= 0x7F;
            uint64_t l[4];
            memcpy(l, buf, 32);
            if (Fq::geq_m(l)) return false;
            fe y = Fq::to_mont(Fq::put(l));
            if (!point_from_y(y, sign, out)) return false;
            // is_in_correct_subgroup_assuming_on_curve: r * P == 0
            uint32_t rl[8];
            for (int i = 0; i < 8; i++) rl[i] = C::Fr::m(i);
            xyzz t = E::mul_scalar(out, rl);
            if (!E::is_identity(t)) return false;
            if (E::is_identity(out)) out = E::affine_identity();
            return true;
        } else {
            // ark-ff 0.4 deserialize_with_flags: only the two SWFlags bits of byte 32 are interpreted (both set ->
            // UnexpectedFlags), the integer comes from the first 32 bytes (its six padding bits are never read), and
            // ark-ec returns the identity on the infinity flag without looking at x -- the same accept/reject decisions,
            // including malleated encodings.
            uint8_t flags = in[32];
            if ((flags & 0xC0) == 0xC0) return false;
            uint64_t l[4];
            memcpy(l, in, 32);
            if (Fq::geq_m(l)) return false;
            if (flags & 0x40) {
                out = E::affine_identity();
                return true;
            }
            fe x = Fq::to_mont(Fq::put(l));
            return point_from_x(x, (flags & 0x80) != 0, out);
        }
    }

    // k*P with k a Montgomery-form scalar (mul_bigint(k.into_bigint())) -> affine
    static affine mul(const affine& p, const fe& k_mont) {
        fe k = Fr::from_mont(k_mont);
        xyzz acc = E::mul_scalar(p, k.v);
        return E::to_affine(acc);
    }
    static affine add(const affine& a, const affine& b) {
        xyzz acc = E::from_affine(a);
        E::madd(acc, b);
        return E::to_affine(acc);
    }
};

}  // namespace bp

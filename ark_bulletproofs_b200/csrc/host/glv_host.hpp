// GLV decomposition on the host (secq256k1: j-invariant 0, phi(x, y) = (beta*x, y) = lambda*(x, y)).
// A uniform fold scalar kappa is split once per IPA round into kappa = k1 + k2*lambda (mod r) with |k1|, |k2| < 2^130,
// so the generator fold's double-and-add chain is ~129 steps instead of 256 (vec_kernels.cuh:
// ipa_fold_points_glv_kernel). Constants and their derivation: tools/gen_consts.py. The result is verified against
// kappa with two field multiplications; on any mismatch the caller uses the plain 256-step fold.
#pragma once
#include <cstdint>
#include <cstring>
#include "fp_host.hpp"

namespace bp {

struct GlvSplit {
    uint32_t k1[5], k2[5];   // magnitudes, little-endian 32-bit limbs (< 2^160)
    int neg1, neg2;          // signs
    int top;                 // index of the highest set bit over both magnitudes (-1 if both are zero)
};

// Joint sparse form (Solinas) of two non-negative magnitudes < 2^160: digits u1[j], u2[j] in {-1, 0, 1} with
// sum u_i[j] 2^j = k_i and at most one zero column... precisely: of any two consecutive columns at least one is (0, 0),
// so a joint double-and-add over {p1, p2, p1 + p2, p1 - p2} adds in about half of its steps (3/4 for plain binary
// digits). Step j is the 4-bit code (u1 + 1) | (u2 + 1) << 2 at bits 4*(j & 7) of code[j >> 3]; top = last step.
struct JsfDigits {
    uint32_t code[21];
    int top;
};
inline bool jsf_digits(const uint32_t k1[5], const uint32_t k2[5], JsfDigits& out) {
    uint32_t a[6] = {k1[0], k1[1], k1[2], k1[3], k1[4], 0}, b[6] = {k2[0], k2[1], k2[2], k2[3], k2[4], 0};
    auto nz = [](const uint32_t* x) { return (x[0] | x[1] | x[2] | x[3] | x[4] | x[5]) != 0; };
    auto shr1 = [](uint32_t* x) {
        for (int i = 0; i < 5; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
        x[5] >>= 1;
    };
    memset(out.code, 0, sizeof(out.code));
    int d1 = 0, d2 = 0, j = 0;
    while (nz(a) || nz(b) || d1 || d2) {
        if (j >= 168) return false;
        const int l1 = (int)((a[0] & 7u) + (uint32_t)d1) & 7, l2 = (int)((b[0] & 7u) + (uint32_t)d2) & 7;
        int u1 = 0, u2 = 0;
        if (l1 & 1) { u1 = 2 - (l1 & 3); if ((l1 == 3 || l1 == 5) && (l2 & 3) == 2) u1 = -u1; }
        if (l2 & 1) { u2 = 2 - (l2 & 3); if ((l2 == 3 || l2 == 5) && (l1 & 3) == 2) u2 = -u2; }
        if (2 * d1 == 1 + u1) d1 = 1 - d1;
        if (2 * d2 == 1 + u2) d2 = 1 - d2;
        shr1(a);
        shr1(b);
        out.code[j >> 3] |= (uint32_t)((u1 + 1) | ((u2 + 1) << 2)) << (4 * (j & 7));
        j++;
    }
    out.top = j - 1;
    // check: the digits reproduce both magnitudes (signed accumulation, MSB first)
    for (int which = 0; which < 2; which++) {
        __int128 hi = 0;            // value >> 64 would overflow 128 bits at 160-bit magnitudes: track two halves
        uint64_t acc[3] = {0, 0, 0};   // 192-bit two's complement
        for (int s = out.top; s >= 0; s--) {
            // acc = 2*acc + u
            uint64_t c = 0;
            for (int i = 0; i < 3; i++) { uint64_t n = (acc[i] << 1) | c; c = acc[i] >> 63; acc[i] = n; }
            const int u = (int)((out.code[s >> 3] >> (4 * (s & 7) + 2 * which)) & 3u) - 1;
            if (u == 1) { for (int i = 0; i < 3; i++) { if (++acc[i] != 0) break; } }
            else if (u == -1) { for (int i = 0; i < 3; i++) { if (acc[i]-- != 0) break; } }
        }
        (void)hi;
        const uint32_t* k = which ? k2 : k1;
        const uint64_t w0 = (uint64_t)k[0] | ((uint64_t)k[1] << 32), w1 = (uint64_t)k[2] | ((uint64_t)k[3] << 32), w2 = k[4];
        if (acc[0] != w0 || acc[1] != w1 || acc[2] != w2) return false;
    }
    return true;
}

template <class C>
struct GlvHost {
    using Fr = HostFp<typename C::Fr>;
    static void mul_n(const uint64_t* a, int na, const uint64_t* b, int nb, uint64_t* out) {
        for (int i = 0; i < na + nb; i++) out[i] = 0;
        for (int i = 0; i < na; i++) {
            unsigned __int128 carry = 0;
            for (int j = 0; j < nb; j++) {
                unsigned __int128 t = (unsigned __int128)a[i] * b[j] + out[i + j] + carry;
                out[i + j] = (uint64_t)t;
                carry = t >> 64;
            }
            out[i + nb] = (uint64_t)carry;
        }
    }
    // c = (k*g + 2^383) >> 384, k: 4 limbs, g: 5 limbs -> 3 limbs
    static void round_mul(const uint64_t k[4], const uint64_t g[5], uint64_t c[3]) {
        uint64_t p[9];
        mul_n(k, 4, g, 5, p);
        unsigned __int128 t = (unsigned __int128)p[5] + (1ull << 63);
        p[5] = (uint64_t)t;
        for (int i = 6; i < 9 && (t >> 64); i++) { t = (unsigned __int128)p[i] + 1; p[i] = (uint64_t)t; }
        c[0] = p[6]; c[1] = p[7]; c[2] = p[8];
    }
    static void sub6(uint64_t* a, const uint64_t* b) {    // a -= b (mod 2^384)
        unsigned __int128 borrow = 0;
        for (int i = 0; i < 6; i++) {
            unsigned __int128 t = (unsigned __int128)a[i] - b[i] - (uint64_t)borrow;
            a[i] = (uint64_t)t;
            borrow = (t >> 64) & 1;
        }
    }
    static bool to_mag(uint64_t v[6], uint32_t out[5], int& neg) {
        neg = (int)(v[5] >> 63);
        if (neg) {                                        // two's complement negate
            unsigned __int128 c = 1;
            for (int i = 0; i < 6; i++) { c += (uint64_t)~v[i]; v[i] = (uint64_t)c; c >>= 64; }
        }
        if (v[3] | v[4] | v[5] | (v[2] >> 32)) return false;
        out[0] = (uint32_t)v[0]; out[1] = (uint32_t)(v[0] >> 32); out[2] = (uint32_t)v[1]; out[3] = (uint32_t)(v[1] >> 32); out[4] = (uint32_t)v[2];
        return true;
    }
    static fe mag_to_fr(const uint32_t m[5], int neg) {
        fe t = Fr::zero();
        for (int i = 0; i < 5; i++) t.v[i] = m[i];
        t = Fr::to_mont(t);
        return neg ? Fr::neg(t) : t;
    }

    static bool split(const fe& kappa_mont, GlvSplit& out) {
        if constexpr (!C::HAS_GLV) {
            return false;
        } else {
            fe kc = Fr::from_mont(kappa_mont);
            uint64_t k[4];
            memcpy(k, kc.v, 32);
            uint64_t c1[3], c2[3];
            round_mul(k, C::GLV_G1, c1);
            round_mul(k, C::GLV_G2, c2);
            uint64_t k1[6] = {k[0], k[1], k[2], k[3], 0, 0}, k2[6], t[6];
            mul_n(c1, 3, C::GLV_A1, 3, t); sub6(k1, t);
            mul_n(c2, 3, C::GLV_A2, 3, t); sub6(k1, t);             // k1 = k - c1*a1 - c2*a2
            mul_n(c1, 3, C::GLV_B1N, 3, k2);                        // k2 = -c1*b1 - c2*b2 = c1*|b1| - c2*b2
            mul_n(c2, 3, C::GLV_B2, 3, t); sub6(k2, t);
            if (!to_mag(k1, out.k1, out.neg1) || !to_mag(k2, out.k2, out.neg2)) return false;
            // verify k1 + k2*lambda == kappa (mod r)
            fe lam;
            for (int i = 0; i < 8; i++) lam.v[i] = C::glv_lambda(i);
            fe chk = Fr::add(mag_to_fr(out.k1, out.neg1), Fr::mul(mag_to_fr(out.k2, out.neg2), lam));
            if (!Fr::eq(chk, kappa_mont)) return false;
            out.top = -1;
            for (int b = 159; b >= 0; b--)
                if (((out.k1[b >> 5] | out.k2[b >> 5]) >> (b & 31)) & 1u) { out.top = b; break; }
            return true;
        }
    }
};

}  // namespace bp

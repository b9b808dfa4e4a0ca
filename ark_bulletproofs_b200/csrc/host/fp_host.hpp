// Host-side 4 x 64-bit Montgomery arithmetic with the same static interface and the same
// memory format as the device `Fp<M>` (csrc/fp.cuh), so `SW<C, HostFp<..>>` gives the host a
// group law for the O(log n) serial tails that a GPU thread is 10x slower at: the Horner
// combination of window sums (256 dependent doublings), challenge inverses, Pedersen
// commitments of single scalars. Bulk work never runs here.
#pragma once
#include <cstdint>
#include <cstring>
#include "../fp.cuh"

namespace bp {

template <class M>
struct HostFp {
    using Mod = M;
    using el = fe;
    template <class C> static fe te_d2() { fe r; for (int i = 0; i < 8; i++) r.v[i] = C::d2(i); return r; }
    template <class C> static fe curve_b() { fe r; for (int i = 0; i < 8; i++) r.v[i] = C::b(i); return r; }
    typedef unsigned __int128 u128;
    static inline uint64_t ml(int i) { return (uint64_t)M::m(2 * i) | ((uint64_t)M::m(2 * i + 1) << 32); }
    static inline void get(const fe& a, uint64_t* o) { memcpy(o, a.v, 32); }
    static inline fe put(const uint64_t* o) { fe r; memcpy(r.v, o, 32); return r; }

    static inline bool geq_m(const uint64_t* t) {
        for (int i = 3; i >= 0; i--) {
            uint64_t mi = ml(i);
            if (t[i] != mi) return t[i] > mi;
        }
        return true;
    }
    static inline void sub_m(uint64_t* t) {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)t[i] - ml(i) - borrow;
            t[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    static inline fe add(const fe& a, const fe& b) {
        uint64_t x[4], y[4], t[4];
        get(a, x); get(b, y);
        uint64_t c = 0;
        for (int i = 0; i < 4; i++) {
            u128 s = (u128)x[i] + y[i] + c;
            t[i] = (uint64_t)s;
            c = (uint64_t)(s >> 64);
        }
        if (c || geq_m(t)) sub_m(t);
        return put(t);
    }
    static inline fe sub(const fe& a, const fe& b) {
        uint64_t x[4], y[4], t[4];
        get(a, x); get(b, y);
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)x[i] - y[i] - borrow;
            t[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            uint64_t c = 0;
            for (int i = 0; i < 4; i++) {
                u128 s = (u128)t[i] + ml(i) + c;
                t[i] = (uint64_t)s;
                c = (uint64_t)(s >> 64);
            }
        }
        return put(t);
    }
    static inline fe zero() { fe r; memset(r.v, 0, 32); return r; }
    static inline bool is_zero(const fe& a) {
        uint32_t o = 0;
        for (int i = 0; i < 8; i++) o |= a.v[i];
        return o == 0;
    }
    static inline bool eq(const fe& a, const fe& b) { return memcmp(a.v, b.v, 32) == 0; }
    static inline fe neg(const fe& a) { return is_zero(a) ? a : sub(zero(), a); }
    static inline fe dbl(const fe& a) { return add(a, a); }
    static inline fe add_l(const fe& a, const fe& b) { return add(a, b); }
    static inline fe sub_l(const fe& a, const fe& b) { return sub(a, b); }
    static inline fe dbl_l(const fe& a) { return add(a, a); }
    static inline fe norm(const fe& a) { return a; }
    static inline fe mul3(const fe& a) { return add(dbl(a), a); }
    static inline fe mul_small(const fe& a, int k) {
        fe r = zero(), p = a;
        for (int bit = 0; bit < 4; bit++) {
            if ((k >> bit) & 1) r = add(r, p);
            p = dbl(p);
        }
        return r;
    }
    // CIOS Montgomery multiplication, 4 x 64
    static inline fe mul(const fe& a, const fe& b) {
        uint64_t x[4], y[4];
        get(a, x); get(b, y);
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            uint64_t c = 0;
            for (int j = 0; j < 4; j++) {
                u128 s = (u128)x[j] * y[i] + t[j] + c;
                t[j] = (uint64_t)s;
                c = (uint64_t)(s >> 64);
            }
            u128 s = (u128)t[4] + c;
            t[4] = (uint64_t)s;
            t[5] = (uint64_t)(s >> 64);
            uint64_t q = t[0] * M::INV64;
            s = (u128)q * ml(0) + t[0];
            c = (uint64_t)(s >> 64);
            for (int j = 1; j < 4; j++) {
                s = (u128)q * ml(j) + t[j] + c;
                t[j - 1] = (uint64_t)s;
                c = (uint64_t)(s >> 64);
            }
            s = (u128)t[4] + c;
            t[3] = (uint64_t)s;
            t[4] = t[5] + (uint64_t)(s >> 64);
            t[5] = 0;
        }
        if (t[4] || geq_m(t)) sub_m(t);
        return put(t);
    }
    static inline fe sqr(const fe& a) { return mul(a, a); }
    static inline fe mul_sub(const fe& a, const fe& b, const fe& c, const fe& d) { return sub(mul(a, b), mul(c, d)); }
    static inline fe mul_add(const fe& a, const fe& b, const fe& c, const fe& d) { return add(mul(a, b), mul(c, d)); }
    static inline fe one() { fe r; for (int i = 0; i < 8; i++) r.v[i] = M::one(i); return r; }
    static inline fe r2() { fe r; for (int i = 0; i < 8; i++) r.v[i] = M::r2(i); return r; }
    static inline fe to_mont(const fe& a) { return mul(a, r2()); }
    static inline fe from_mont(const fe& a) { fe o = zero(); o.v[0] = 1; return mul(a, o); }
    static inline fe from_u32(uint32_t x) { fe o = zero(); o.v[0] = x; return to_mont(o); }
    static inline fe from_u64(uint64_t x) { fe o = zero(); o.v[0] = (uint32_t)x; o.v[1] = (uint32_t)(x >> 32); return to_mont(o); }
    static inline fe pow(const fe& a, const uint32_t* e) {
        fe r = one();
        bool started = false;
        for (int i = 7; i >= 0; i--)
            for (int bit = 31; bit >= 0; bit--) {
                if (started) r = sqr(r);
                if ((e[i] >> bit) & 1u) { r = started ? mul(r, a) : a; started = true; }
            }
        return r;
    }
    static inline fe inv(const fe& a) {
        uint32_t e[8];
        e[0] = M::m(0) - 2u;
        for (int i = 1; i < 8; i++) e[i] = M::m(i);
        return pow(a, e);
    }
    // canonical integer comparison helpers (ark-ff `Ord` compares canonical values)
    static inline bool canonical_lt(const fe& a, const fe& b) {   // both already canonical (non-Montgomery)
        for (int i = 7; i >= 0; i--)
            if (a.v[i] != b.v[i]) return a.v[i] < b.v[i];
        return false;
    }
};

}  // namespace bp

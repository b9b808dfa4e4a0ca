// Host build of the device math templates (fp.cuh / ec.cuh) for CPU unit tests.
// The carry primitives are emulated on the host (see fp.cuh), so the exact same
// algorithms that run in the kernels are checked here against the Python oracle.
#include <cstring>
#define BP_FP29_CHECK 1
#include "../../ark_bulletproofs_b200/csrc/ec.cuh"
#include "../../ark_bulletproofs_b200/csrc/host/fp_host.hpp"
#include "../../ark_bulletproofs_b200/csrc/experimental/fp29.cuh"
#include "../../ark_bulletproofs_b200/csrc/host/glv_host.hpp"
#include "../../ark_bulletproofs_b200/csrc/msm_sort.cuh"
#include "../../ark_bulletproofs_b200/csrc/host/workers.hpp"
#include <atomic>
using namespace bp;

// Fp<M>::mul_sparse where it exists (the device templates), the ordinary product for the host reference class
template <class F> static auto fp_mul_sparse(const fe& x, const fe& y) -> decltype(F::mul_sparse(x, y)) { return F::mul_sparse(x, y); }
template <class F, class... A> static fe fp_mul_sparse(const fe& x, const fe& y, A...) { return F::mul(x, y); }

template <class F> static int fp_op_t(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    fe x, y, r;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    switch (op) {
        case 0: r = F::mul(x, y); break;
        case 1: r = F::add(x, y); break;
        case 2: r = F::sub(x, y); break;
        case 3: r = F::inv(x); break;
        case 4: r = F::from_mont(x); break;
        case 5: r = F::to_mont(x); break;
        case 6: r = F::neg(x); break;
        case 7: r = F::sqr(x); break;
        case 8: r = fp_mul_sparse<F>(x, y); break;   // alternative reduction for sparse moduli (fp.cuh)
        default: return -1;
    }
    memcpy(out, r.v, 32);
    return 0;
}
extern "C" int hm_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    switch (field) {
        case 0: return fp_op_t<Fp<SecqFq>>(op, a, b, out);
        case 1: return fp_op_t<Fp<SecqFr>>(op, a, b, out);
        case 2: return fp_op_t<Fp<ZorroFq>>(op, a, b, out);
        case 3: return fp_op_t<Fp<Fp25519>>(op, a, b, out);
        case 4: return fp_op_t<Fp<Fr25519>>(op, a, b, out);
        case 10: return fp_op_t<HostFp<SecqFq>>(op, a, b, out);
        case 11: return fp_op_t<HostFp<SecqFr>>(op, a, b, out);
        case 12: return fp_op_t<HostFp<ZorroFq>>(op, a, b, out);
        case 13: return fp_op_t<HostFp<Fp25519>>(op, a, b, out);
        case 14: return fp_op_t<HostFp<Fr25519>>(op, a, b, out);
    }
    return -1;
}

// four-operand field operations: a*b - c*d and a*b + c*d with one reduction (Fp::mul2)
template <class F> static int fp_op4_t(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out) {
    fe x, y, z, w, r;
    memcpy(x.v, a, 32); memcpy(y.v, b, 32); memcpy(z.v, c, 32); memcpy(w.v, d, 32);
    switch (op) {
        case 0: r = F::mul_sub(x, y, z, w); break;
        case 1: r = F::mul_add(x, y, z, w); break;
        default: return -1;
    }
    memcpy(out, r.v, 32);
    return 0;
}
extern "C" int hm_fp_op4(int field, int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out) {
    switch (field) {
        case 0: return fp_op4_t<Fp<SecqFq>>(op, a, b, c, d, out);
        case 1: return fp_op4_t<Fp<SecqFr>>(op, a, b, c, d, out);
        case 2: return fp_op4_t<Fp<ZorroFq>>(op, a, b, c, d, out);
        case 3: return fp_op4_t<Fp<Fp25519>>(op, a, b, c, d, out);
        case 4: return fp_op4_t<Fp<Fr25519>>(op, a, b, c, d, out);
        case 10: return fp_op4_t<HostFp<SecqFq>>(op, a, b, c, d, out);
        case 11: return fp_op4_t<HostFp<SecqFr>>(op, a, b, c, d, out);
        case 12: return fp_op4_t<HostFp<ZorroFq>>(op, a, b, c, d, out);
        case 13: return fp_op4_t<HostFp<Fp25519>>(op, a, b, c, d, out);
        case 14: return fp_op4_t<HostFp<Fr25519>>(op, a, b, c, d, out);
    }
    return -1;
}

// points cross as affine (x,y) Montgomery, (0,0) = identity
template <class E> static int ec_op_t(int op, const uint32_t* p, const uint32_t* q, const uint32_t* s, uint32_t* out) {
    affine P, Q;
    memcpy(&P, p, 64);
    memcpy(&Q, q, 64);
    xyzz r;
    switch (op) {
        case 0: r = E::from_affine(P); E::madd(r, Q); break;                          // P + Q (mixed)
        case 1: { xyzz a = E::dbl_affine(P); xyzz b = E::from_affine(Q); b = E::dbl(b); r = a; E::add(r, b); break; }  // 2P + 2Q (full add on non-trivial Z)
        case 2: r = E::dbl(E::dbl_affine(P)); break;                                  // 4P
        case 3: r = E::mul_scalar(P, s); break;                                       // s*P
        case 4: { r = E::dbl_affine(P); E::madd(r, Q); break; }                       // 2P + Q
        case 5: { r = E::dbl_affine(P); xyzz b = E::dbl_affine(Q); E::add(r, b); break; }  // 2P + 2Q via dbl_affine
        case 6: r = E::mul_u32(E::dbl_affine(P), s[0]); break;                        // s0 * 2P
        default: return -1;
    }
    affine o = E::to_affine(r);
    memcpy(out, &o, 64);
    return E::on_curve(o) ? 0 : 1;
}
extern "C" int hm_ec_op(int curve, int op, const uint32_t* p, const uint32_t* q, const uint32_t* s, uint32_t* out) {
    switch (curve) {
        case 0: return ec_op_t<SW<Secq256k1>>(op, p, q, s, out);
        case 1: return ec_op_t<SW<Zorro>>(op, p, q, s, out);
        case 2: return ec_op_t<TE<Curve25519>>(op, p, q, s, out);
        case 12: return ec_op_t<TE<Curve25519, HostFp<Fp25519>>>(op, p, q, s, out);
        case 10: return ec_op_t<SW<Secq256k1, HostFp<SecqFq>>>(op, p, q, s, out);
        case 11: return ec_op_t<SW<Zorro, HostFp<ZorroFq>>>(op, p, q, s, out);
    }
    return -1;
}


// ---- Fp29 (balanced 9 x 29-bit limbs, Montgomery domain 2^261); values cross as canonical 8 x 32-bit integers ----
extern "C" long hm_fp29_violations() { return fp29_violations(); }
template <class M> static int fp29_op_t(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    using F = Fp29<M>;
    fe x, y;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    fl X = F::unpack(x), Y = F::unpack(y), R;
    switch (op) {
        case 0: R = F::mul(X, Y); break;
        case 1: R = F::add(X, Y); break;
        case 2: R = F::sub(X, Y); break;
        case 3: R = F::sqr(X); break;
        case 4: R = F::neg(X); break;
        case 5: R = F::mul3(X); break;
        // chains that exercise the loose bounds: ((x - y) - y + x) * (x + y + y) ; and ((x-y)^2 - 3x) * (y - x)
        case 6: R = F::mul(F::add(F::sub(F::sub_l(X, Y), Y), X), F::add(F::add_l(X, Y), Y)); break;
        case 7: R = F::mul(F::sub(F::sqr(F::sub(X, Y)), F::mul3(X)), F::sub(Y, X)); break;
        case 8: { fe o = F::to_storage(X); memcpy(out, o.v, 32); return F::is_zero(X) ? 1 : 0; }
        case 9: { bool e = F::eq(X, Y); fl d = F::sub_l(X, Y); fe o = F::pack_canonical(d); memcpy(out, o.v, 32); return e ? 1 : 0; }
        case 10: R = F::from_storage(x); break;                              // x * 2^5 as an integer (domain 2^256 -> 2^261)
        case 11: {                                                           // the MSM's pre-converted base format
            fe u = Fp<M>::add(x, F::bias_d());                               // caller passes the canonical integer U - D... see test
            R = F::load_biased(F::to_biased(u));
            break;
        }
        case 12: R = F::inv(X); break;
        case 13: R = F::mul_small(X, (int)(y.v[0] & 7)); break;
        // is_zero on multiples of m reached through arithmetic: (x - y) + (y - x), x*3 - x - x - x, and k*m literally
        case 14: { fl d = F::add_l(F::sub_l(X, Y), F::sub_l(Y, X)); return F::is_zero(d) ? 1 : 0; }
        case 15: {
            fl d = X;
            const int k = (int)(y.v[0] % 9) - 4;
            for (int i = 0; i < 9; i++) d.v[i] += k * M::mb(i);
            int z = F::is_zero(F::sub_l(d, X)) ? 1 : 0;                       // k*m, un-normalized
            fe o = F::pack_canonical(d);
            memcpy(out, o.v, 32);
            return z;
        }
        default: return -1;
    }
    fe o = F::pack_canonical(R);
    memcpy(out, o.v, 32);
    return 0;
}
extern "C" int hm_fp29_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    switch (field) {
        case 0: return fp29_op_t<SecqFq>(op, a, b, out);
        case 1: return fp29_op_t<SecqFr>(op, a, b, out);
        case 2: return fp29_op_t<ZorroFq>(op, a, b, out);
        case 3: return fp29_op_t<Fp25519>(op, a, b, out);
        case 4: return fp29_op_t<Fr25519>(op, a, b, out);
    }
    return -1;
}

// group law over Fp29: points cross as canonical affine (x,y) in the 2^261 domain, (0,0) = identity
template <class E> static int ec29_op_t(int op, const uint32_t* p, const uint32_t* q, const uint32_t* s, uint32_t* out) {
    using F = typename E::F;
    affine P, Q;
    memcpy(&P, p, 64);
    memcpy(&Q, q, 64);
    typename E::aff Pl{F::unpack(P.x), F::unpack(P.y)}, Ql{F::unpack(Q.x), F::unpack(Q.y)};
    typename E::ext r;
    switch (op) {
        case 0: r = E::from_affine(Pl); E::madd(r, Ql); break;
        case 1: { typename E::ext a = E::dbl_affine(Pl); typename E::ext b = E::from_affine(Ql); b = E::dbl(b); r = a; E::add(r, b); break; }
        case 2: r = E::dbl(E::dbl_affine(Pl)); break;
        case 3: r = E::mul_scalar(Pl, s); break;
        case 4: { r = E::dbl_affine(Pl); E::madd(r, Ql); break; }
        case 5: { r = E::dbl_affine(Pl); typename E::ext b = E::dbl_affine(Ql); E::add(r, b); break; }
        case 6: {                                                           // long chain: ((P + Q) + Q) ... 64 mixed additions, then doublings
            r = E::from_affine(Pl);
            for (int i = 0; i < 64; i++) E::madd(r, (i & 1) ? Pl : Ql);
            for (int i = 0; i < 8; i++) r = E::dbl(r);
            typename E::ext t = r;
            for (int i = 0; i < 8; i++) E::add(r, t);
            break;
        }
        case 7: { typename E::aff o = E::to_affine(E::mul_scalar(Pl, s)); r = E::from_affine(o); if (!E::on_curve(o)) return -2; break; }
        default: return -1;
    }
    // return the projective result as 4 canonical coordinates (128 B): the caller normalises
    fe o[4] = {F::pack_canonical(r.x), F::pack_canonical(r.y), F::pack_canonical(r.zz), F::pack_canonical(r.zzz)};
    memcpy(out, o, 128);
    return E::is_identity(r) ? 1 : 0;
}
extern "C" int hm_ec29_op(int curve, int op, const uint32_t* p, const uint32_t* q, const uint32_t* s, uint32_t* out) {
    switch (curve) {
        case 0: return ec29_op_t<SW<Secq256k1, Fp29<SecqFq>>>(op, p, q, s, out);
        case 1: return ec29_op_t<SW<Zorro, Fp29<ZorroFq>>>(op, p, q, s, out);
        case 2: return ec29_op_t<TE<Curve25519, Fp29<Fp25519>>>(op, p, q, s, out);
    }
    return -1;
}

// GLV split of a Montgomery scalar of secq256k1 (host/glv_host.hpp): out = k1[5] | k2[5] | neg1 | neg2 | top
extern "C" int hm_glv_split(const uint32_t* kappa_mont, uint32_t* out) {
    fe k;
    memcpy(k.v, kappa_mont, 32);
    GlvSplit s;
    if (!GlvHost<Secq256k1>::split(k, s)) return 0;
    memcpy(out, s.k1, 20);
    memcpy(out + 5, s.k2, 20);
    out[10] = (uint32_t)s.neg1; out[11] = (uint32_t)s.neg2; out[12] = (uint32_t)s.top;
    return 1;
}

// joint sparse form of two 160-bit magnitudes (host/glv_host.hpp); returns top, -1000 on failure
extern "C" int hm_jsf(const uint32_t* k1, const uint32_t* k2, uint32_t* code21) {
    JsfDigits d;
    if (!jsf_digits(k1, k2, d)) return -1000;
    memcpy(code21, d.code, sizeof(d.code));
    return d.top;
}

// planner of the MSM's bucket sort (csrc/msm_sort.cuh): out = {ok, low_bits, low_top, nb1, tiles, cap, tile}
extern "C" void hm_sort_plan(uint64_t n, int W, int cb, uint32_t* out) {
    SortPlan p = make_sort_plan((size_t)n, W, cb);
    out[0] = p.ok; out[1] = (uint32_t)p.low_bits; out[2] = (uint32_t)p.low_top; out[3] = p.nb1; out[4] = p.tiles;
    out[5] = SORT_BIN_CAP; out[6] = SORT_TS;
}

// persistent host workers (host/workers.hpp): `iters` batches of 1..8 jobs; returns 0 when every job ran exactly once
extern "C" int hm_workers_stress(int iters) {
    HostWorkers w(7);
    std::atomic<long> sum{0};
    long want = 0;
    for (int it = 0; it < iters; it++) {
        const int n = 1 + it % 8;
        std::function<void(int)> fn = [&](int j) { sum += (long)(j + 1) * (it + 1); };
        w.run(n, fn);
        want += (long)n * (n + 1) / 2 * (it + 1);
        if (sum.load() != want) return 1;              // run() returns only when the whole batch is done
    }
    return w.size() == 7 ? 0 : 2;
}

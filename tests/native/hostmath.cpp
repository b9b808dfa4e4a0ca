// Host build of the device math templates (fp.cuh / ec.cuh) for CPU unit tests.
// The carry primitives are emulated on the host (see fp.cuh), so the exact same
// algorithms that run in the kernels are checked here against the Python oracle.
#include <cstring>
#include "../../ark_bulletproofs_b200/csrc/ec.cuh"
#include "../../ark_bulletproofs_b200/csrc/host/fp_host.hpp"
using namespace bp;

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

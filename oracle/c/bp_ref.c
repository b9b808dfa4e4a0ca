/*
 * CPU oracle / CPU baseline (TEST INFRASTRUCTURE ONLY) -- C restatement of the arithmetic the
 * reference reaches through ark-ec / ark-ff 0.4 (un-vendored):
 *
 *   - Fp256 Montgomery arithmetic, 4 x 64-bit limbs, R = 2^256 (ark-ff MontBackend)
 *   - short-Weierstrass Jacobian group law (ark-ec short_weierstrass::Projective:
 *     add-2007-bl, madd-2007-bl, dbl-2009-l / dbl-2007-bl for a != 0)
 *   - VariableBaseMSM::msm as ark-ec 0.4.x implements it for curves with cheap negation
 *     (`msm_bigint_wnaf`): window c = 3 if n < 32 else ln_without_floor(n) + 2 with
 *     ln_without_floor(n) = log2(n) * 69 / 100, signed digits, 2^c buckets per window,
 *     per-window running sum, windows independent (parallel only with feature `parallel`,
 *     Cargo.toml:76), final Horner over windows.      [restated from the published crate]
 *   - the per-element generator fold of src/inner_product_proof.rs:139-156,216-225:
 *     a 2-point msm (c = 3) + into_affine per output point.
 *
 * PARITY UNPINNED (see oracle/bp_oracle.py): checked against the Python restatement in
 * tests/test_c_oracle.py. Only tests/, smoke() and bench.py's CPU legs may call this.
 *
 * Build: see oracle/Makefile  ->  oracle/_build/libbp_ref.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;
typedef struct { fe x, y; } aff;          /* (0,0) = identity */
typedef struct { fe x, y, z; } jac;       /* z == 0 = identity */

typedef struct {
    uint64_t m[4];      /* modulus */
    uint64_t inv;       /* -m^-1 mod 2^64 */
    fe one;             /* R mod m */
    fe r2;              /* R^2 mod m */
} field_t;

typedef struct {
    field_t fq, fr;
    int a;              /* curve coefficient a (small integer) */
} curve_t;

/* ---- constants ---------------------------------------------------- */
static void field_init(field_t* f, const uint64_t m[4]) {
    memcpy(f->m, m, 32);
    uint64_t inv = 1;
    for (int i = 0; i < 63; i++) { inv *= inv; inv *= m[0]; }   /* m^(2^63-1) = m^-1 mod 2^64 */
    f->inv = (uint64_t)(0 - inv);
    /* R mod m by 256 modular doublings of 1, R^2 by 256 more */
    uint64_t t[4] = {1, 0, 0, 0};
    for (int k = 0; k < 512; k++) {
        uint64_t c = 0, s[4];
        for (int i = 0; i < 4; i++) { uint64_t n = (t[i] << 1) | c; c = t[i] >> 63; t[i] = n; }
        uint64_t b = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - m[i] - b; s[i] = (uint64_t)d; b = (uint64_t)(d >> 64) & 1; }
        if (c || !b) memcpy(t, s, 32);
        if (k == 255) memcpy(f->one.v, t, 32);
    }
    memcpy(f->r2.v, t, 32);
}

static curve_t CURVES[2];
static int curves_ready = 0;
static void curves_init(void) {
    if (curves_ready) return;
    const uint64_t secq_q[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
    const uint64_t secq_r[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
    const uint64_t zorro_q[4] = {0x885F1923D3651021ULL, 0x69F40306A6210BEDULL, 0x0000000000000001ULL, 0x8000000000000000ULL};
    const uint64_t p25519[4] = {0xFFFFFFFFFFFFFFEDULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL};
    field_init(&CURVES[0].fq, secq_q); field_init(&CURVES[0].fr, secq_r); CURVES[0].a = 0;
    field_init(&CURVES[1].fq, zorro_q); field_init(&CURVES[1].fr, p25519); CURVES[1].a = 6;
    curves_ready = 1;
}

/* ---- field ------------------------------------------------------------------------------- */
static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }
static inline int geq(const uint64_t* t, const uint64_t* m) {
    for (int i = 3; i >= 0; i--) if (t[i] != m[i]) return t[i] > m[i];
    return 1;
}
static inline void fe_add(const field_t* f, fe* r, const fe* a, const fe* b) {
    uint64_t t[4], c = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)a->v[i] + b->v[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    if (c || geq(t, f->m)) { uint64_t bb = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - f->m[i] - bb; t[i] = (uint64_t)d; bb = (uint64_t)(d >> 64) & 1; } }
    memcpy(r->v, t, 32);
}
static inline void fe_sub(const field_t* f, fe* r, const fe* a, const fe* b) {
    uint64_t t[4], bb = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->v[i] - b->v[i] - bb; t[i] = (uint64_t)d; bb = (uint64_t)(d >> 64) & 1; }
    if (bb) { uint64_t c = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)t[i] + f->m[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } }
    memcpy(r->v, t, 32);
}
static inline void fe_neg(const field_t* f, fe* r, const fe* a) {
    if (fe_is_zero(a)) { *r = *a; return; }
    fe z; memset(&z, 0, 32); fe_sub(f, r, &z, a);
}
static inline void fe_mul(const field_t* f, fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) { u128 s = (u128)a->v[j] * b->v[i] + t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        u128 s = (u128)t[4] + c; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
        uint64_t q = t[0] * f->inv;
        s = (u128)q * f->m[0] + t[0]; c = (uint64_t)(s >> 64);
        for (int j = 1; j < 4; j++) { s = (u128)q * f->m[j] + t[j] + c; t[j - 1] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        s = (u128)t[4] + c; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64); t[5] = 0;
    }
    if (t[4] || geq(t, f->m)) { uint64_t bb = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - f->m[i] - bb; t[i] = (uint64_t)d; bb = (uint64_t)(d >> 64) & 1; } }
    memcpy(r->v, t, 32);
}
static inline void fe_sqr(const field_t* f, fe* r, const fe* a) { fe_mul(f, r, a, a); }
static inline void fe_dbl(const field_t* f, fe* r, const fe* a) { fe_add(f, r, a, a); }
static void fe_inv(const field_t* f, fe* r, const fe* a) {   /* Fermat */
    uint64_t e[4]; memcpy(e, f->m, 32); e[0] -= 2;
    fe acc = f->one;
    for (int i = 255; i >= 0; i--) {
        fe_sqr(f, &acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(f, &acc, &acc, a);
    }
    *r = acc;
}
static inline void fe_from_mont(const field_t* f, fe* r, const fe* a) { fe o; memset(&o, 0, 32); o.v[0] = 1; fe_mul(f, r, a, &o); }
static inline void fe_small(const field_t* f, fe* r, const fe* a, int k) {
    fe acc, p = *a; memset(&acc, 0, 32);
    for (int b = 0; b < 4; b++) { if ((k >> b) & 1) fe_add(f, &acc, &acc, &p); fe_dbl(f, &p, &p); }
    *r = acc;
}

/* ---- Jacobian group law (ark-ec short_weierstrass::Projective) ----------------------------- */
static inline int aff_is_id(const aff* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void jac_set_id(jac* p) { memset(p, 0, sizeof(*p)); }
static void jac_dbl(const curve_t* cv, jac* r, const jac* p) {
    const field_t* f = &cv->fq;
    if (fe_is_zero(&p->z) || fe_is_zero(&p->y)) { jac_set_id(r); return; }
    fe A, B, C, D, E, F, t, zz;
    fe_sqr(f, &A, &p->x); fe_sqr(f, &B, &p->y); fe_sqr(f, &C, &B);
    fe_add(f, &t, &p->x, &B); fe_sqr(f, &t, &t); fe_sub(f, &t, &t, &A); fe_sub(f, &t, &t, &C); fe_dbl(f, &D, &t);
    fe_dbl(f, &E, &A); fe_add(f, &E, &E, &A);
    if (cv->a) { fe_sqr(f, &zz, &p->z); fe_sqr(f, &zz, &zz); fe_small(f, &zz, &zz, cv->a); fe_add(f, &E, &E, &zz); }
    fe_sqr(f, &F, &E);
    fe z3; fe_mul(f, &z3, &p->y, &p->z); fe_dbl(f, &z3, &z3);
    fe x3; fe_sub(f, &x3, &F, &D); fe_sub(f, &x3, &x3, &D);
    fe c8; fe_dbl(f, &c8, &C); fe_dbl(f, &c8, &c8); fe_dbl(f, &c8, &c8);
    fe y3; fe_sub(f, &y3, &D, &x3); fe_mul(f, &y3, &E, &y3); fe_sub(f, &y3, &y3, &c8);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_madd(const curve_t* cv, jac* r, const jac* p, const aff* q) {   /* madd-2007-bl */
    const field_t* f = &cv->fq;
    if (aff_is_id(q)) { *r = *p; return; }
    if (fe_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = f->one; return; }
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t;
    fe_sqr(f, &z1z1, &p->z); fe_mul(f, &u2, &q->x, &z1z1);
    fe_mul(f, &s2, &q->y, &p->z); fe_mul(f, &s2, &s2, &z1z1);
    if (fe_eq(&u2, &p->x)) {
        if (fe_eq(&s2, &p->y)) { jac_dbl(cv, r, p); return; }
        jac_set_id(r); return;
    }
    fe_sub(f, &h, &u2, &p->x); fe_sqr(f, &hh, &h); fe_dbl(f, &i, &hh); fe_dbl(f, &i, &i);
    fe_mul(f, &j, &h, &i); fe_sub(f, &rr, &s2, &p->y); fe_dbl(f, &rr, &rr); fe_mul(f, &v, &p->x, &i);
    fe x3; fe_sqr(f, &x3, &rr); fe_sub(f, &x3, &x3, &j); fe_sub(f, &x3, &x3, &v); fe_sub(f, &x3, &x3, &v);
    fe y3; fe_sub(f, &y3, &v, &x3); fe_mul(f, &y3, &rr, &y3); fe_mul(f, &t, &p->y, &j); fe_dbl(f, &t, &t); fe_sub(f, &y3, &y3, &t);
    fe z3; fe_add(f, &z3, &p->z, &h); fe_sqr(f, &z3, &z3); fe_sub(f, &z3, &z3, &z1z1); fe_sub(f, &z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(const curve_t* cv, jac* r, const jac* p, const jac* q) {   /* add-2007-bl */
    const field_t* f = &cv->fq;
    if (fe_is_zero(&p->z)) { *r = *q; return; }
    if (fe_is_zero(&q->z)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fe_sqr(f, &z1z1, &p->z); fe_sqr(f, &z2z2, &q->z);
    fe_mul(f, &u1, &p->x, &z2z2); fe_mul(f, &u2, &q->x, &z1z1);
    fe_mul(f, &s1, &p->y, &q->z); fe_mul(f, &s1, &s1, &z2z2);
    fe_mul(f, &s2, &q->y, &p->z); fe_mul(f, &s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { jac_dbl(cv, r, p); return; }
        jac_set_id(r); return;
    }
    fe_sub(f, &h, &u2, &u1); fe_dbl(f, &i, &h); fe_sqr(f, &i, &i); fe_mul(f, &j, &h, &i);
    fe_sub(f, &rr, &s2, &s1); fe_dbl(f, &rr, &rr); fe_mul(f, &v, &u1, &i);
    fe x3; fe_sqr(f, &x3, &rr); fe_sub(f, &x3, &x3, &j); fe_sub(f, &x3, &x3, &v); fe_sub(f, &x3, &x3, &v);
    fe y3; fe_sub(f, &y3, &v, &x3); fe_mul(f, &y3, &rr, &y3); fe_mul(f, &t, &s1, &j); fe_dbl(f, &t, &t); fe_sub(f, &y3, &y3, &t);
    fe z3; fe_add(f, &z3, &p->z, &q->z); fe_sqr(f, &z3, &z3); fe_sub(f, &z3, &z3, &z1z1); fe_sub(f, &z3, &z3, &z2z2); fe_mul(f, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_to_aff(const curve_t* cv, aff* r, const jac* p) {   /* into_affine: one inversion */
    const field_t* f = &cv->fq;
    if (fe_is_zero(&p->z)) { memset(r, 0, sizeof(*r)); return; }
    fe zi, zi2, zi3;
    fe_inv(f, &zi, &p->z); fe_sqr(f, &zi2, &zi); fe_mul(f, &zi3, &zi2, &zi);
    fe_mul(f, &r->x, &p->x, &zi2); fe_mul(f, &r->y, &p->y, &zi3);
}

/* ---- ark-ec 0.4 msm_bigint_wnaf ------------------------------------------------------------- */
static int ark_window(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;          /* floor(log2 n) */
    if (((size_t)1 << lg) < n) lg++;                     /* ark_std::log2 = ceil */
    return lg * 69 / 100 + 2;
}
/* make_digits: signed c-bit digits, little-endian window order */
static void make_digits(const uint64_t s[4], int c, int ndig, int64_t* out) {
    uint64_t carry = 0;
    const uint64_t radix = 1ull << c, mask = radix - 1;
    for (int i = 0; i < ndig; i++) {
        int bit = i * c, limb = bit >> 6, sh = bit & 63;
        uint64_t d = 0;
        if (limb < 4) {
            d = s[limb] >> sh;
            if (sh + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
        }
        d = (d & mask) + carry;
        carry = (d + radix / 2) >> c;
        int64_t sd = (int64_t)d - (int64_t)(carry << c);
        if (i == ndig - 1) sd += (int64_t)(carry << c);   /* top digit keeps the carry */
        out[i] = sd;
    }
}

static void msm_core(const curve_t* cv, const aff* bases, const fe* scalars_mont, size_t n, jac* out, int threads) {
    const field_t* fr = &cv->fr;
    jac_set_id(out);
    if (n == 0) return;
    int c = ark_window(n);
    int num_bits = 256;
    { int top = 255; while (top > 0 && !((fr->m[top >> 6] >> (top & 63)) & 1)) top--; num_bits = top + 1; }
    int ndig = (num_bits + c - 1) / c;
    int64_t* digits = (int64_t*)malloc(sizeof(int64_t) * n * (size_t)ndig);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        fe s; fe_from_mont(fr, &s, &scalars_mont[i]);   /* into_bigint */
        make_digits(s.v, c, ndig, digits + i * ndig);
    }
    jac* wsum = (jac*)malloc(sizeof(jac) * ndig);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int w = 0; w < ndig; w++) {
        size_t nbk = (size_t)1 << c;
        jac* buckets = (jac*)calloc(nbk, sizeof(jac));
        for (size_t i = 0; i < n; i++) {
            int64_t d = digits[i * ndig + w];
            if (d > 0) jac_madd(cv, &buckets[d - 1], &buckets[d - 1], &bases[i]);
            else if (d < 0) { aff nb = bases[i]; fe_neg(&cv->fq, &nb.y, &nb.y); jac_madd(cv, &buckets[-d - 1], &buckets[-d - 1], &nb); }
        }
        jac run, res; jac_set_id(&run); jac_set_id(&res);
        for (size_t b = nbk; b-- > 0;) { jac_add(cv, &run, &run, &buckets[b]); jac_add(cv, &res, &res, &run); }
        wsum[w] = res;
        free(buckets);
    }
    jac total; jac_set_id(&total);
    for (int w = ndig - 1; w >= 1; w--) {
        jac_add(cv, &total, &total, &wsum[w]);
        for (int k = 0; k < c; k++) jac_dbl(cv, &total, &total);
    }
    jac_add(cv, out, &wsum[0], &total);
    free(wsum);
    free(digits);
}

/* G::Group::msm(bases, scalars).into_affine(); threads = 1 is the crate's default build */
int ref_msm(int curve, const uint8_t* bases_xy, const uint8_t* scalars_mont, size_t n, uint8_t out_xy[64], int threads) {
    curves_init();
    if (curve < 0 || curve > 1) return -1;
    const curve_t* cv = &CURVES[curve];
    jac r;
    if (threads < 1) threads = 1;
    msm_core(cv, (const aff*)bases_xy, (const fe*)scalars_mont, n, &r, threads);
    aff a; jac_to_aff(cv, &a, &r);
    memcpy(out_xy, &a, 64);
    return 0;
}

/* The reference's generator fold for one IPA round (src/inner_product_proof.rs:216-225):
 *   G_L[i] = msm([G_L[i], G_R[i]], [sL, sR]).into_affine()   for i < h, in place. */
int ref_fold_points(int curve, uint8_t* pts_xy, size_t h, const uint8_t* sL_mont, const uint8_t* sR_mont, int threads) {
    curves_init();
    if (curve < 0 || curve > 1) return -1;
    const curve_t* cv = &CURVES[curve];
    aff* P = (aff*)pts_xy;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < h; i++) {
        aff b2[2] = {P[i], P[h + i]};
        fe s2[2]; memcpy(&s2[0], sL_mont, 32); memcpy(&s2[1], sR_mont, 32);
        jac r; msm_core(cv, b2, s2, 2, &r, 1);
        jac_to_aff(cv, &P[i], &r);
    }
    return 0;
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Synthetic workload for the CPU arm: out[i] = (start + i + 1) * G as affine Montgomery pairs
 * (same points as the GPU side's bp_synth_points_device). g_xy = generator (Montgomery). */
int ref_synth_points(int curve, const uint8_t* g_xy, uint8_t* out_xy, size_t n, uint64_t start) {
    curves_init();
    if (curve < 0 || curve > 1) return -1;
    const curve_t* cv = &CURVES[curve];
    const field_t* f = &cv->fq;
    aff g; memcpy(&g, g_xy, 64);
    /* (start+1)*G by double-and-add */
    jac acc; jac_set_id(&acc);
    uint64_t k = start + 1;
    for (int bit = 63; bit >= 0; bit--) { jac_dbl(cv, &acc, &acc); if ((k >> bit) & 1) jac_madd(cv, &acc, &acc, &g); }
    const size_t CH = 1024;
    jac* buf = (jac*)malloc(sizeof(jac) * CH);
    fe* pre = (fe*)malloc(sizeof(fe) * CH);
    aff* out = (aff*)out_xy;
    for (size_t base = 0; base < n; base += CH) {
        size_t m = n - base < CH ? n - base : CH;
        for (size_t i = 0; i < m; i++) { buf[i] = acc; jac_madd(cv, &acc, &acc, &g); }
        /* batch inversion of z (none is zero: the points are small multiples of G) */
        fe run = f->one;
        for (size_t i = 0; i < m; i++) { pre[i] = run; fe_mul(f, &run, &run, &buf[i].z); }
        fe inv; fe_inv(f, &inv, &run);
        for (size_t i = m; i-- > 0;) {
            fe zi, zi2, zi3; fe_mul(f, &zi, &inv, &pre[i]); fe_mul(f, &inv, &inv, &buf[i].z);
            fe_sqr(f, &zi2, &zi); fe_mul(f, &zi3, &zi2, &zi);
            fe_mul(f, &out[base + i].x, &buf[i].x, &zi2); fe_mul(f, &out[base + i].y, &buf[i].y, &zi3);
        }
    }
    free(buf); free(pre);
    return 0;
}

/* ---- per-element generator fold with per-element scalars (first IPA round, src/inner_product_proof.rs:139-156):
 *   out[i] = msm([L[i], R[i]], [sL[i], sR[i]]).into_affine() */
int ref_fold_points_v(int curve, const uint8_t* L_xy, const uint8_t* R_xy, size_t h, const uint8_t* sL_mont, const uint8_t* sR_mont,
                      uint8_t* out_xy, int threads) {
    curves_init();
    if (curve < 0 || curve > 1) return -1;
    const curve_t* cv = &CURVES[curve];
    const aff *L = (const aff*)L_xy, *R = (const aff*)R_xy;
    aff* O = (aff*)out_xy;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < h; i++) {
        aff b2[2] = {L[i], R[i]};
        fe s2[2]; memcpy(&s2[0], sL_mont + 32 * i, 32); memcpy(&s2[1], sR_mont + 32 * i, 32);
        jac r; msm_core(cv, b2, s2, 2, &r, 1);
        jac_to_aff(cv, &O[i], &r);
    }
    return 0;
}

/* ---- merlin TranscriptRng draws (merlin 3.0 src/transcript.rs TranscriptRng::fill_bytes over strobe.rs; restated from the
 * published crate, pinned by the public Merlin KAT through oracle/bp_oracle.py's Strobe128, against which
 * tests/test_c_oracle.py checks this function): the prover's 2n blinding scalars are 8n dependent Keccak-f[1600]
 * permutations, far too slow in pure Python at 2^16..2^20 multipliers. ------------------------------------------------ */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t a[25]) {
    static const int rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], rot[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= KRC[rnd];
    }
}
typedef struct { uint8_t st[200]; int pos, pos_begin, cur_flags; } strobe_t;
enum { SR = 166, FI = 1, FA = 2, FC = 4, FT = 8, FM = 16, FK = 32 };
static void strobe_run_f(strobe_t* s) {
    s->st[s->pos] ^= (uint8_t)s->pos_begin;
    s->st[s->pos + 1] ^= 0x04;
    s->st[SR + 1] ^= 0x80;
    uint64_t lanes[25];
    memcpy(lanes, s->st, 200);          /* little-endian host */
    keccak_f(lanes);
    memcpy(s->st, lanes, 200);
    s->pos = 0; s->pos_begin = 0;
}
static void strobe_absorb(strobe_t* s, const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) { s->st[s->pos++] ^= d[i]; if (s->pos == SR) strobe_run_f(s); }
}
static void strobe_squeeze(strobe_t* s, uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) { d[i] = s->st[s->pos]; s->st[s->pos++] = 0; if (s->pos == SR) strobe_run_f(s); }
}
static void strobe_begin_op(strobe_t* s, int flags) {
    uint8_t hdr[2] = {(uint8_t)s->pos_begin, (uint8_t)flags};
    s->pos_begin = s->pos + 1;
    s->cur_flags = flags;
    strobe_absorb(s, hdr, 2);
    if ((flags & (FC | FK)) && s->pos != 0) strobe_run_f(s);
}
static uint64_t trng_next_u64(strobe_t* s) {      /* fill_bytes(8): meta_ad(LE32(8), false); prf(8, false) */
    const uint8_t len[4] = {8, 0, 0, 0};
    strobe_begin_op(s, FM | FA);
    strobe_absorb(s, len, 4);
    strobe_begin_op(s, FI | FA | FC);
    uint8_t out[8];
    strobe_squeeze(s, out, 8);
    uint64_t v; memcpy(&v, out, 8);
    return v;
}
/* `count` draws of ark-ff Fp::rand (4 x next_u64, top bits shaved to the modulus' bit length, rejected unless < m);
 * out = the accepted raw limbs, i.e. the Montgomery representation. state = strobe bytes, io = {pos, pos_begin, cur_flags}. */
int ref_trng_scalars(uint8_t state[200], int io[3], const uint8_t modulus_le[32], size_t count, uint8_t* out_raw) {
    strobe_t s;
    memcpy(s.st, state, 200); s.pos = io[0]; s.pos_begin = io[1]; s.cur_flags = io[2];
    uint64_t m[4]; memcpy(m, modulus_le, 32);
    int bits = 256;
    while (bits > 0 && !((m[(bits - 1) / 64] >> ((bits - 1) % 64)) & 1)) bits--;
    const uint64_t mask = bits >= 256 ? ~0ULL : (~0ULL >> (256 - bits));
    for (size_t i = 0; i < count;) {
        uint64_t l[4];
        for (int k = 0; k < 4; k++) l[k] = trng_next_u64(&s);
        l[3] &= mask;
        if (!geq(l, m)) { memcpy(out_raw + 32 * i, l, 32); i++; }
    }
    memcpy(state, s.st, 200); io[0] = s.pos; io[1] = s.pos_begin; io[2] = s.cur_flags;
    return 0;
}

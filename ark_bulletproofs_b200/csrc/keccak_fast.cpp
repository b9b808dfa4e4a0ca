// AVX-512 Keccak-f[1600] for the host-side Fiat-Shamir chain (compiled by g++, no device code).
//
// Why it exists: merlin's TranscriptRng runs one Keccak-f per next_u64 and ark-ff's Fp::rand draws four of
// them per scalar, so the 2n blinding scalars s_L, s_R of Prover::prove (src/r1cs/prover.rs:510-513,599-602)
// cost 8n dependent permutations on one host core - the serial floor of every byte-identical prover
// (SURVEY.md 0.5). This file shortens that chain: the state lives in five zmm registers (one 5-lane plane
// each); theta is two 3-way XORs (vpternlogq) and two lane rotations, rho a per-lane vprolvq, pi a lane
// permutation that leaves register x holding column x so chi is one vpternlogq per register, and a 5x5
// transpose (unpack + vpermt2q) restores the plane layout. The chain is latency bound (one round feeds the
// next), so the dependent path is kept short: three of the five fifth-column elements ride through chi in the
// spare lanes 5..7 of column 0 (the pi permutes of columns 0..2 reach into a second plane), the other two are
// blended in after the permutes, and iota is applied to plane 0 at the start of the next round.
// The RNG loop keeps the state in registers between draws. Installed at load time only when the CPU reports AVX-512F/VL; otherwise the scalar code
// in host/merlin.hpp stays in place (same results: tests/test_host_layer.py runs both).
#include <immintrin.h>
#include <cstddef>
#include <cstdint>
#include "host/merlin.hpp"

namespace bp {
namespace {

#define BP_AVX512 __attribute__((target("avx512f,avx512vl"), always_inline)) inline

struct Planes { __m512i p0, p1, p2, p3, p4; };

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

BP_AVX512 void load(Planes& s, const uint64_t* A) {
    const __mmask8 m5 = 0x1F;
    s.p0 = _mm512_maskz_loadu_epi64(m5, A); s.p1 = _mm512_maskz_loadu_epi64(m5, A + 5); s.p2 = _mm512_maskz_loadu_epi64(m5, A + 10);
    s.p3 = _mm512_maskz_loadu_epi64(m5, A + 15); s.p4 = _mm512_maskz_loadu_epi64(m5, A + 20);
}
BP_AVX512 void store(const Planes& s, uint64_t* A) {
    const __mmask8 m5 = 0x1F;
    _mm512_mask_storeu_epi64(A, m5, s.p0); _mm512_mask_storeu_epi64(A + 5, m5, s.p1); _mm512_mask_storeu_epi64(A + 10, m5, s.p2);
    _mm512_mask_storeu_epi64(A + 15, m5, s.p3); _mm512_mask_storeu_epi64(A + 20, m5, s.p4);
}

BP_AVX512 void permute(Planes& s) {
    __m512i p0 = s.p0, p1 = s.p1, p2 = s.p2, p3 = s.p3, p4 = s.p4;
    const __m512i thPrev = _mm512_setr_epi64(4, 0, 1, 2, 3, 5, 6, 7), thNext = _mm512_setr_epi64(1, 2, 3, 4, 0, 5, 6, 7);
    // rho offsets of A[x + 5y], one vector per plane y
    const __m512i rho0 = _mm512_setr_epi64(0, 1, 62, 28, 27, 0, 0, 0), rho1 = _mm512_setr_epi64(36, 44, 6, 55, 20, 0, 0, 0),
                  rho2 = _mm512_setr_epi64(3, 10, 43, 25, 39, 0, 0, 0), rho3 = _mm512_setr_epi64(41, 45, 15, 21, 8, 0, 0, 0),
                  rho4 = _mm512_setr_epi64(18, 2, 61, 56, 14, 0, 0, 0);
    // pi: B[y, (2x+3y)%5] = A[x, y]  =>  lane j of plane y takes x = (3j + y) % 5; afterwards register y is column y of B.
    // Lanes 5..7 of the first three columns carry a second copy of rows 1, 0, 3 of columns 4, 0, 1 (vpermt2q reaches
    // into a second plane at no extra cost), so chi on column 0 also yields A'[4,1], A'[4,0], A'[4,3] in its spare
    // lanes -- exactly where the transpose below picks up the fifth element of planes 1, 0 and 3.
    const __m512i pi0x = _mm512_setr_epi64(0, 3, 1, 4, 2, 8 + 2, 8 + 4, 8 + 3);   // + B[4,1], B[4,0], B[4,3] from plane 4 (pi4 = 4,2,0,3,1)
    const __m512i pi1x = _mm512_setr_epi64(1, 4, 2, 0, 3, 8 + 3, 8 + 0, 8 + 4);   // + B[0,1], B[0,0], B[0,3] from plane 0 (pi0 = 0,3,1,4,2)
    const __m512i pi2x = _mm512_setr_epi64(2, 0, 3, 1, 4, 8 + 4, 8 + 1, 8 + 0);   // + B[1,1], B[1,0], B[1,3] from plane 1 (pi1 = 1,4,2,0,3)
    const __m512i pi3 = _mm512_setr_epi64(3, 1, 4, 2, 0, 5, 6, 7), pi4 = _mm512_setr_epi64(4, 2, 0, 3, 1, 5, 6, 7);
    // transpose indices: lanes 0..3 from the unpacked column pairs, lane 4 from the spare lanes where available
    const __m512i trA = _mm512_setr_epi64(0, 1, 8, 9, 6, 5, 5, 5);     // plane 0: t0[0,1], t2[0,1], t0[6] = n0[6] = A'[4,0]
    const __m512i trB = _mm512_setr_epi64(0, 1, 8, 9, 4, 5, 5, 5);     // plane 1: t1[0,1], t3[0,1], t1[4] = n0[5] = A'[4,1]
    const __m512i trC = _mm512_setr_epi64(2, 3, 10, 11, 5, 5, 5, 5);   // plane 2: lane 4 inserted below
    const __m512i trD = _mm512_setr_epi64(2, 3, 10, 11, 6, 5, 5, 5);   // plane 3: t1[6] = n0[7] = A'[4,3]
    const __m512i trE = _mm512_setr_epi64(4, 5, 12, 13, 5, 5, 5, 5);   // plane 4: lane 4 inserted below
    const __m512i s2 = _mm512_set1_epi64(2);
    for (int r = 0; r < 24; r++) {
        // iota of the previous round is applied here, to plane 0 only: after chi it would delay all five planes
        if (r) p0 = _mm512_xor_si512(p0, _mm512_maskz_set1_epi64(1, (long long)RC[r - 1]));
        __m512i c = _mm512_ternarylogic_epi64(_mm512_ternarylogic_epi64(p1, p3, p0, 0x96), p2, p4, 0x96);   // planes 2, 4 arrive last
        __m512i dp = _mm512_permutexvar_epi64(thPrev, c);
        __m512i dn = _mm512_rol_epi64(_mm512_permutexvar_epi64(thNext, c), 1);
        p0 = _mm512_rolv_epi64(_mm512_ternarylogic_epi64(p0, dp, dn, 0x96), rho0);
        p1 = _mm512_rolv_epi64(_mm512_ternarylogic_epi64(p1, dp, dn, 0x96), rho1);
        p2 = _mm512_rolv_epi64(_mm512_ternarylogic_epi64(p2, dp, dn, 0x96), rho2);
        p3 = _mm512_rolv_epi64(_mm512_ternarylogic_epi64(p3, dp, dn, 0x96), rho3);
        p4 = _mm512_rolv_epi64(_mm512_ternarylogic_epi64(p4, dp, dn, 0x96), rho4);
        __m512i b0 = _mm512_permutex2var_epi64(p0, pi0x, p4), b1 = _mm512_permutex2var_epi64(p1, pi1x, p0),
                b2 = _mm512_permutex2var_epi64(p2, pi2x, p1), b3 = _mm512_permutexvar_epi64(pi3, p3), b4 = _mm512_permutexvar_epi64(pi4, p4);
        // chi along x: 0xD2 = a ^ (~b & c); register x = column x, lane = row
        __m512i n0 = _mm512_ternarylogic_epi64(b0, b1, b2, 0xD2), n1 = _mm512_ternarylogic_epi64(b1, b2, b3, 0xD2),
                n2 = _mm512_ternarylogic_epi64(b2, b3, b4, 0xD2), n3 = _mm512_ternarylogic_epi64(b3, b4, b0, 0xD2),
                n4 = _mm512_ternarylogic_epi64(b4, b0, b1, 0xD2);
        // columns -> planes
        __m512i t0 = _mm512_unpacklo_epi64(n0, n1), t1 = _mm512_unpackhi_epi64(n0, n1), t2 = _mm512_unpacklo_epi64(n2, n3),
                t3 = _mm512_unpackhi_epi64(n2, n3);
        p0 = _mm512_permutex2var_epi64(t0, trA, t2);
        p1 = _mm512_permutex2var_epi64(t1, trB, t3);
        p3 = _mm512_permutex2var_epi64(t1, trD, t3);
        // planes 2 and 4 take their fifth element from column 4 itself; a 1-cycle blend after the 3-cycle permutes
        p2 = _mm512_mask_blend_epi64(0x10, _mm512_permutex2var_epi64(t0, trC, t2), _mm512_permutexvar_epi64(s2, n4));
        p4 = _mm512_mask_blend_epi64(0x10, _mm512_permutex2var_epi64(t0, trE, t2), n4);       // n4[4] is already in lane 4
    }
    p0 = _mm512_xor_si512(p0, _mm512_maskz_set1_epi64(1, (long long)RC[23]));
    s.p0 = p0; s.p1 = p1; s.p2 = p2; s.p3 = p3; s.p4 = p4;
}

__attribute__((target("avx512f,avx512vl"))) void keccak_f1600_avx512(uint64_t* A) {
    Planes s;
    load(s, A);
    permute(s);
    store(s, A);
}

// `count` steady-state TranscriptRng::next_u64 draws (see TranscriptRng::draw_u64 in host/merlin.hpp for the
// byte-level derivation of the per-draw constants), state resident in registers.
__attribute__((target("avx512f,avx512vl"))) void rng_draw_avx512(uint64_t* A, uint64_t* out, size_t count) {
    Planes s;
    load(s, A);
    const __m512i in0 = _mm512_setr_epi64(0, (long long)TranscriptRng::DRAW_LANE1, (long long)TranscriptRng::DRAW_LANE2, 0, 0, 0, 0, 0);
    const __m512i in4 = _mm512_setr_epi64((long long)0x8000000000000000ULL, 0, 0, 0, 0, 0, 0, 0);
    for (size_t i = 0; i < count; i++) {
        s.p0 = _mm512_xor_si512(s.p0, in0);
        s.p4 = _mm512_xor_si512(s.p4, in4);
        permute(s);
        out[i] = (uint64_t)_mm_cvtsi128_si64(_mm512_castsi512_si128(s.p0));
        s.p0 = _mm512_maskz_mov_epi64(0x1E, s.p0);   // squeeze zeroes what it returns
    }
    store(s, A);
}

struct Install {
    Install() {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) {
            g_keccak = keccak_f1600_avx512;
            g_rng_draw = rng_draw_avx512;
        }
    }
} install;

}  // namespace

// 0 = scalar, 1 = AVX-512; `which` < 0 only queries. Lets the tests run both implementations.
int keccak_select(int which) {
    if (which == 0) { g_keccak = keccak_f1600_scalar; g_rng_draw = nullptr; }
    if (which == 1) {
        __builtin_cpu_init();
        if (!(__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl"))) return -1;
        g_keccak = keccak_f1600_avx512; g_rng_draw = rng_draw_avx512;
    }
    return g_keccak == keccak_f1600_avx512 ? 1 : 0;
}

}  // namespace bp

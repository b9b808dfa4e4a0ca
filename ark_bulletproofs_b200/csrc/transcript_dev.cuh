// Device-side Fiat-Shamir transcript: Keccak-f[1600], STROBE-128 as merlin 3.0 drives it,
// merlin::Transcript::append_message / challenge_bytes, and TranscriptProtocol::challenge_scalar
// (src/transcript.rs:45-101: 32 challenge bytes -> ChaCha20Rng::from_seed -> ScalarField::rand) -- one transcript per
// thread, so that the many independent transcripts of a batch verification (src/r1cs/verifier.rs:604-691) derive their
// inner-product challenges (src/inner_product_proof.rs:266-277) in one launch.
//
// A single transcript is slower here than on the host (one thread needs ~10 us per permutation where a host core needs
// 0.3 us), so the prover and single verifications keep the host transcript; the device engine pays
// when there are hundreds of transcripts. The host hands over its STROBE state after the challenge `w`
// (verifier.rs:459); the kernel appends the IPA domain separator, every (L_j, R_j) pair in ark-serialize's uncompressed
// form, draws the u_j, inverts them with Montgomery's trick and draws `r` from a clone (verifier.rs:516-519).
// Checked challenge by challenge against the oracle (tests/test_r1cs_gpu.py::test_device_transcript_*).
#pragma once
#include "gens_kernels.cuh"

namespace bp {

__device__ __forceinline__ uint64_t dk_rotl(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }

// Keccak-f[1600] on 25 lanes held by one thread
__device__ inline void keccak_f1600_dev(uint64_t* A) {
    const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], d[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ dk_rotl(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; i++) A[i] ^= d[i % 5];
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) {
                const int rot = ROT[x + 5 * y];
                b[y + 5 * ((2 * x + 3 * y) % 5)] = rot ? dk_rotl(A[x + 5 * y], rot) : A[x + 5 * y];
            }
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) A[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        A[0] ^= RC[r];
    }
}

// STROBE-128 (merlin/src/strobe.rs); the 200 state bytes live in 25 little-endian lanes
struct DevStrobe {
    uint64_t st[25];
    uint32_t pos, pos_begin;
    static constexpr uint32_t R = 166;
    enum : uint32_t { FI = 1, FA = 2, FC = 4, FT = 8, FM = 16, FK = 32 };
    __device__ __forceinline__ void xor_byte(uint32_t at, uint32_t b) { st[at >> 3] ^= (uint64_t)b << (8 * (at & 7)); }
    __device__ void run_f() {
        xor_byte(pos, pos_begin);
        xor_byte(pos + 1, 0x04);
        xor_byte(R + 1, 0x80);
        keccak_f1600_dev(st);
        pos = 0;
        pos_begin = 0;
    }
    __device__ void absorb_byte(uint32_t b) {
        xor_byte(pos, b);
        if (++pos == R) run_f();
    }
    __device__ uint32_t squeeze_byte() {
        const uint32_t sh = 8 * (pos & 7);
        const uint32_t b = (uint32_t)(st[pos >> 3] >> sh) & 0xFFu;
        st[pos >> 3] &= ~((uint64_t)0xFF << sh);
        if (++pos == R) run_f();
        return b;
    }
    __device__ void begin_op(uint32_t flags) {
        const uint32_t old_begin = pos_begin;
        pos_begin = pos + 1;
        absorb_byte(old_begin);
        absorb_byte(flags);
        if ((flags & (FC | FK)) && pos != 0) run_f();
    }
    // append_message(label, msg): meta_ad(label) ; meta_ad(LE32(len), more) ; ad(msg)
    __device__ void append_begin(const char* label, int llen, uint32_t msg_len) {
        begin_op(FM | FA);
        for (int i = 0; i < llen; i++) absorb_byte((uint8_t)label[i]);
        for (int i = 0; i < 4; i++) absorb_byte((msg_len >> (8 * i)) & 0xFFu);
        begin_op(FA);
    }
    // challenge_bytes(label, 32) -> 8 little-endian words
    __device__ void challenge32(const char* label, int llen, uint32_t out[8]) {
        begin_op(FM | FA);
        for (int i = 0; i < llen; i++) absorb_byte((uint8_t)label[i]);
        absorb_byte(32); absorb_byte(0); absorb_byte(0); absorb_byte(0);
        begin_op(FI | FA | FC);
#pragma unroll 1
        for (int w = 0; w < 8; w++) {
            uint32_t v = 0;
            for (int k = 0; k < 4; k++) v |= squeeze_byte() << (8 * k);
            out[w] = v;
        }
    }
};

// ScalarField::rand(&mut ChaCha20Rng::from_seed(seed)) (ark-ff Fp::rand: 4 x next_u64, shave, accept iff < r); the raw
// limbs are the Montgomery representation
template <class C>
__device__ fe challenge_scalar_dev(DevStrobe& s, const char* label, int llen) {
    uint32_t key[8];
    s.challenge32(label, llen, key);
    uint32_t blk[16];
    fe out;
#pragma unroll 1
    for (uint32_t t = 0; t < 256; t++) {
        if ((t & 1u) == 0) chacha20_block_dev(key, t >> 1, blk);      // try t reads words [8t, 8t + 8)
        const uint32_t* w = blk + 8 * (t & 1u);
#pragma unroll
        for (int k = 0; k < 8; k++) out.v[k] = w[k];
        if (C::Fr::BITS < 256) out.v[7] &= 0xFFFFFFFFu >> (256 - C::Fr::BITS);
        bool geq = true;
        for (int k = 7; k >= 0; k--) {
            const uint32_t mk = C::Fr::m(k);
            if (out.v[k] != mk) { geq = out.v[k] > mk; break; }
        }
        if (!geq) break;
    }
    return out;
}

// append_point(label, P): ark-serialize uncompressed form (src/transcript.rs:75-79) of a Montgomery affine point
template <class C>
__device__ void append_point_dev(DevStrobe& s, const char* label, int llen, const affine& p) {
    using F = Fp<typename C::Fq>;
    const bool te = C::KIND == 1;
    s.append_begin(label, llen, te ? 64u : 65u);
    const fe xc = F::from_mont(p.x), yc = F::from_mont(p.y);
    for (int k = 0; k < 8; k++)
        for (int b = 0; b < 4; b++) s.absorb_byte((xc.v[k] >> (8 * b)) & 0xFFu);
    for (int k = 0; k < 8; k++)
        for (int b = 0; b < 4; b++) s.absorb_byte((yc.v[k] >> (8 * b)) & 0xFFu);
    if (!te) {
        // SWFlags of the y coordinate: bit 7 set iff y > -y as canonical integers
        const fe nc = F::from_mont(F::neg(p.y));
        bool larger = false;
        for (int k = 7; k >= 0; k--)
            if (yc.v[k] != nc.v[k]) { larger = yc.v[k] > nc.v[k]; break; }
        s.absorb_byte(larger ? 0x80u : 0u);
    }
}

struct DevTranscriptIn {
    uint64_t st[25];          // STROBE state after the challenge `w`
    uint32_t pos, pos_begin;
    uint32_t lg_n;            // rounds
    uint32_t pt_off;          // first L of this proof in the point array: L_0..L_{k-1}, R_0..R_{k-1}
    uint64_t padded_n;
};

// One thread per proof. out_u[p * 32 + j] = u_j, out_uinv = their inverses (0 stays 0), out_r[p] = the challenge `r`
// of the cloned transcript, status[p] = 1 when an L_j / R_j is the identity (validate_and_append_point -> VerificationError).
template <class C>
__global__ void __launch_bounds__(64) verifier_ipa_challenges_kernel(const DevTranscriptIn* __restrict__ in, const affine* __restrict__ pts, size_t count,
                                                                     fe* __restrict__ out_u, fe* __restrict__ out_uinv, fe* __restrict__ out_r,
                                                                     uint8_t* __restrict__ status) {
    using E = GroupLaw<C>;
    using Fr = Fp<typename C::Fr>;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    DevStrobe s;
    for (int i = 0; i < 25; i++) s.st[i] = in[p].st[i];
    s.pos = in[p].pos;
    s.pos_begin = in[p].pos_begin;
    const uint32_t k = in[p].lg_n;
    const affine* L = pts + in[p].pt_off;
    const affine* R = L + k;
    // innerproduct_domain_sep(n): append_message("dom-sep", "ipp v1"); append_u64("n", n)
    s.append_begin("dom-sep", 7, 6);
    { const char* m = "ipp v1"; for (int i = 0; i < 6; i++) s.absorb_byte((uint8_t)m[i]); }
    s.append_begin("n", 1, 8);
    for (int i = 0; i < 8; i++) s.absorb_byte((uint32_t)(in[p].padded_n >> (8 * i)) & 0xFFu);
    fe* u = out_u + p * 32;
    fe* ui = out_uinv + p * 32;
    uint8_t bad = 0;
    fe run = Fr::one();
#pragma unroll 1
    for (uint32_t j = 0; j < k; j++) {
        const affine l = ld_affine(L + j), r = ld_affine(R + j);
        if (E::is_identity(l) || E::is_identity(r)) { bad = 1; break; }
        append_point_dev<C>(s, "L", 1, l);
        append_point_dev<C>(s, "R", 1, r);
        const fe c = challenge_scalar_dev<C>(s, "u", 1);
        st_fe(u + j, c);
        st_fe(ui + j, run);                       // prefix product of the non-zero challenges before j
        if (!Fr::is_zero(c)) run = Fr::mul(run, c);
    }
    status[p] = bad;
    if (bad) return;
    // ark_ff::batch_inversion: one inversion, zeros stay zeros
    fe inv = Fr::inv(run);
#pragma unroll 1
    for (uint32_t j = k; j-- > 0;) {
        const fe c = ld_fe_rw(u + j);
        if (Fr::is_zero(c)) { st_fe(ui + j, c); continue; }
        const fe pre = ld_fe_rw(ui + j);
        st_fe(ui + j, Fr::mul(inv, pre));
        inv = Fr::mul(inv, c);
    }
    // r is drawn from a clone: the state `s` is simply not used afterwards
    st_fe(out_r + p, challenge_scalar_dev<C>(s, "r", 1));
}

}  // namespace bp

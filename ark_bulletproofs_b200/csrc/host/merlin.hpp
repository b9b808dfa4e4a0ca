// Host-side Fiat-Shamir machinery, restated bit-exactly:
//   Keccak-f[1600], STROBE-128 as used by merlin 3.0 (strobe.rs), merlin::Transcript,
//   merlin's TranscriptRng (one permutation per next_u64), ChaCha20Rng (rand_chacha 0.3),
//   SHA3-512 (sha3 0.10) for src/generators.rs:47-93.
// The reference drives these from src/transcript.rs:45-101 and src/r1cs/prover.rs:483-513.
// They are inherently serial and stay on the host ("boundary").
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace bp {

// ---- Keccak-f[1600] ---------------------------------------------------------------------
static inline uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }

static inline void keccak_f1600_scalar(uint64_t* A) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    uint64_t a00 = A[0], a01 = A[1], a02 = A[2], a03 = A[3], a04 = A[4], a05 = A[5], a06 = A[6], a07 = A[7], a08 = A[8], a09 = A[9],
             a10 = A[10], a11 = A[11], a12 = A[12], a13 = A[13], a14 = A[14], a15 = A[15], a16 = A[16], a17 = A[17], a18 = A[18],
             a19 = A[19], a20 = A[20], a21 = A[21], a22 = A[22], a23 = A[23], a24 = A[24];
    for (int r = 0; r < 24; r++) {
        // theta
        uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                 c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        // rho + pi  (B[y + 5*((2x+3y)%5)] = rot(A[x+5y]))
        uint64_t b00 = a00, b10 = rotl64(a01, 1), b20 = rotl64(a02, 62), b05 = rotl64(a03, 28), b15 = rotl64(a04, 27);
        uint64_t b16 = rotl64(a05, 36), b01 = rotl64(a06, 44), b11 = rotl64(a07, 6), b21 = rotl64(a08, 55), b06 = rotl64(a09, 20);
        uint64_t b07 = rotl64(a10, 3), b17 = rotl64(a11, 10), b02 = rotl64(a12, 43), b12 = rotl64(a13, 25), b22 = rotl64(a14, 39);
        uint64_t b23 = rotl64(a15, 41), b08 = rotl64(a16, 45), b18 = rotl64(a17, 15), b03 = rotl64(a18, 21), b13 = rotl64(a19, 8);
        uint64_t b14 = rotl64(a20, 18), b24 = rotl64(a21, 2), b09 = rotl64(a22, 61), b19 = rotl64(a23, 56), b04 = rotl64(a24, 14);
        // chi
        a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
        a00 ^= RC[r];
    }
    A[0] = a00; A[1] = a01; A[2] = a02; A[3] = a03; A[4] = a04; A[5] = a05; A[6] = a06; A[7] = a07; A[8] = a08; A[9] = a09;
    A[10] = a10; A[11] = a11; A[12] = a12; A[13] = a13; A[14] = a14; A[15] = a15; A[16] = a16; A[17] = a17; A[18] = a18; A[19] = a19;
    A[20] = a20; A[21] = a21; A[22] = a22; A[23] = a23; A[24] = a24;
}

// The permutation actually used. csrc/keccak_fast.cpp replaces both pointers at load time with AVX-512
// versions when the CPU has them (bp::keccak_select switches for the tests); translation units that do not
// link that file (tests/native) keep the scalar code.
using keccak_fn = void (*)(uint64_t*);
using rng_draw_fn = void (*)(uint64_t* state, uint64_t* out, size_t count);
inline keccak_fn g_keccak = keccak_f1600_scalar;
inline rng_draw_fn g_rng_draw = nullptr;
static inline void keccak_f1600(uint64_t* A) { g_keccak(A); }
int keccak_select(int which);

// ---- SHA3-512 ---------------------------------------------------------------------------
static inline void sha3_512(const uint8_t* data, size_t len, uint8_t out[64]) {
    const size_t rate = 72;
    alignas(8) uint8_t st[200];
    memset(st, 0, 200);
    while (len >= rate) {
        for (size_t i = 0; i < rate; i++) st[i] ^= data[i];
        keccak_f1600(reinterpret_cast<uint64_t*>(st));
        data += rate;
        len -= rate;
    }
    for (size_t i = 0; i < len; i++) st[i] ^= data[i];
    st[len] ^= 0x06;
    st[rate - 1] ^= 0x80;
    keccak_f1600(reinterpret_cast<uint64_t*>(st));
    memcpy(out, st, 64);
}

// ---- STROBE-128 (merlin/src/strobe.rs) --------------------------------------------------------
struct Strobe128 {
    static constexpr uint8_t R = 166;
    static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
    alignas(8) uint8_t state[200];
    uint8_t pos = 0, pos_begin = 0, cur_flags = 0;

    Strobe128() { memset(state, 0, 200); }
    explicit Strobe128(const uint8_t* label, size_t len) {
        memset(state, 0, 200);
        const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(state, init, 6);
        memcpy(state + 6, "STROBEv1.0.2", 12);
        keccak_f1600(reinterpret_cast<uint64_t*>(state));
        meta_ad(label, len, false);
    }
    void run_f() {
        state[pos] ^= pos_begin;
        state[pos + 1] ^= 0x04;
        state[R + 1] ^= 0x80;
        keccak_f1600(reinterpret_cast<uint64_t*>(state));
        pos = 0;
        pos_begin = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            state[pos] ^= d[i];
            if (++pos == R) run_f();
        }
    }
    void overwrite(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            state[pos] = d[i];
            if (++pos == R) run_f();
        }
    }
    void squeeze(uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            d[i] = state[pos];
            state[pos] = 0;
            if (++pos == R) run_f();
        }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;   // merlin asserts cur_flags == flags
        uint8_t old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
    void ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
    void prf(uint8_t* d, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(d, n); }
    void key(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A | FLAG_C, more); overwrite(d, n); }
};

static inline void le32(uint32_t v, uint8_t out[4]) { out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16); out[3] = (uint8_t)(v >> 24); }

// ---- RNG interface (rand_core::RngCore) -----------------------------------------------------
struct Rng {
    virtual ~Rng() {}
    virtual uint32_t next_u32() = 0;
    virtual uint64_t next_u64() = 0;
    virtual void fill_bytes(uint8_t* out, size_t n) = 0;
    // `count` consecutive next_u64 results (what ark-ff's BigInt::rand consumes); overridable for bulk draws
    virtual void draw_u64(uint64_t* out, size_t count) { for (size_t i = 0; i < count; i++) out[i] = next_u64(); }
};

// ChaCha20Rng of rand_chacha 0.3: 64-bit block counter (words 12,13), stream id 0 (words 14,15)
struct ChaCha20Rng : Rng {
    uint32_t key[8];
    uint64_t counter = 0;
    uint32_t buf[16];
    int idx = 16;
    uint64_t words_used = 0;
    explicit ChaCha20Rng(const uint8_t seed[32]) { memcpy(key, seed, 32); }
    static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    static void block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
        uint32_t s[16] = {0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                          (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
        uint32_t w[16];
        memcpy(w, s, 64);
#define BP_QR(a, b, c, d)                                                                                \
    w[a] += w[b]; w[d] = rotl(w[d] ^ w[a], 16); w[c] += w[d]; w[b] = rotl(w[b] ^ w[c], 12);              \
    w[a] += w[b]; w[d] = rotl(w[d] ^ w[a], 8);  w[c] += w[d]; w[b] = rotl(w[b] ^ w[c], 7);
        for (int i = 0; i < 10; i++) {
            BP_QR(0, 4, 8, 12) BP_QR(1, 5, 9, 13) BP_QR(2, 6, 10, 14) BP_QR(3, 7, 11, 15)
            BP_QR(0, 5, 10, 15) BP_QR(1, 6, 11, 12) BP_QR(2, 7, 8, 13) BP_QR(3, 4, 9, 14)
        }
#undef BP_QR
        for (int i = 0; i < 16; i++) out[i] = w[i] + s[i];
    }
    // position the stream at an absolute 32-bit word index (the keystream is seekable)
    void seek_word(uint64_t word) {
        counter = word / 16;
        block(key, counter, buf);
        counter++;
        idx = (int)(word % 16);
    }
    uint32_t next_u32() override {
        if (idx == 16) { block(key, counter++, buf); idx = 0; }
        words_used++;
        return buf[idx++];
    }
    uint64_t next_u64() override {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return lo | (hi << 32);
    }
    void fill_bytes(uint8_t* out, size_t n) override {
        while (n) {
            uint32_t w = next_u32();
            size_t k = n < 4 ? n : 4;
            memcpy(out, &w, k);
            out += k;
            n -= k;
        }
    }
};

// ---- merlin::Transcript ----------------------------------------------------------------------
struct TranscriptRng : Rng {
    Strobe128 strobe;
    void fill_bytes(uint8_t* out, size_t n) override {
        uint8_t l[4];
        le32((uint32_t)n, l);
        strobe.meta_ad(l, 4, false);
        strobe.prf(out, n, false);
    }
    uint32_t next_u32() override { uint32_t v; fill_bytes(reinterpret_cast<uint8_t*>(&v), 4); return v; }
    uint64_t next_u64() override { uint64_t v; draw_u64(&v, 1); return v; }
    // Steady state of consecutive next_u64 calls. After a prf squeeze of 8 bytes the STROBE state has
    // pos = 8, pos_begin = 0 and state[0..8] = 0; the next fill_bytes(8) then absorbs, at fixed offsets,
    //   meta_ad header [old_begin = 0, M|A = 0x12] at 8..9, the length LE32(8) at 10..13,
    //   prf header [old_begin = 9, I|A|C = 7] at 14..15, and run_f XORs pos_begin = 15 at 16, 0x04 at 17, 0x80 at 167
    // i.e. three 64-bit lane XORs, one permutation, and lane 0 is the output (then zeroed).
    static constexpr uint64_t DRAW_LANE1 = 0x0709000000081200ULL, DRAW_LANE2 = 0x040FULL;
    void draw_u64(uint64_t* out, size_t count) override {
        while (count) {
            if (strobe.pos == 8 && strobe.pos_begin == 0) {
                uint64_t* A = reinterpret_cast<uint64_t*>(strobe.state);
                if (g_rng_draw) { g_rng_draw(A, out, count); return; }
                for (size_t i = 0; i < count; i++) {
                    A[1] ^= DRAW_LANE1; A[2] ^= DRAW_LANE2; A[20] ^= 0x8000000000000000ULL;
                    keccak_f1600(A);
                    out[i] = A[0];
                    A[0] = 0;
                }
                return;
            }
            fill_bytes(reinterpret_cast<uint8_t*>(out), 8);
            out++; count--;
        }
    }
};

struct Transcript {
    Strobe128 strobe;
    Transcript() {}
    Transcript(const uint8_t* label, size_t len) : strobe(reinterpret_cast<const uint8_t*>("Merlin v1.0"), 11) {
        append_message("dom-sep", label, len);
    }
    explicit Transcript(const char* label) : Transcript(reinterpret_cast<const uint8_t*>(label), strlen(label)) {}
    void append_message(const char* label, const uint8_t* msg, size_t len) { append_message_l((const uint8_t*)label, strlen(label), msg, len); }
    void append_message_l(const uint8_t* label, size_t llen, const uint8_t* msg, size_t len) {
        uint8_t l[4];
        le32((uint32_t)len, l);
        strobe.meta_ad(label, llen, false);
        strobe.meta_ad(l, 4, true);
        strobe.ad(msg, len, false);
    }
    void append_u64(const char* label, uint64_t x) {
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
        append_message(label, b, 8);
    }
    void challenge_bytes(const char* label, uint8_t* out, size_t n) { challenge_bytes_l((const uint8_t*)label, strlen(label), out, n); }
    void challenge_bytes_l(const uint8_t* label, size_t llen, uint8_t* out, size_t n) {
        uint8_t l[4];
        le32((uint32_t)n, l);
        strobe.meta_ad(label, llen, false);
        strobe.meta_ad(l, 4, true);
        strobe.prf(out, n, false);
    }
    // build_rng().rekey_with_witness_bytes(label, w)...finalize(rng)   (merlin transcript.rs)
    TranscriptRng make_rng(const char* label, const std::vector<std::vector<uint8_t>>& witnesses, Rng& external) const {
        TranscriptRng r;
        r.strobe = strobe;
        for (auto& w : witnesses) {
            uint8_t l[4];
            le32((uint32_t)w.size(), l);
            r.strobe.meta_ad((const uint8_t*)label, strlen(label), false);
            r.strobe.meta_ad(l, 4, true);
            r.strobe.key(w.data(), w.size(), false);
        }
        uint8_t rb[32];
        external.fill_bytes(rb, 32);
        r.strobe.meta_ad((const uint8_t*)"rng", 3, false);
        r.strobe.key(rb, 32, false);
        return r;
    }
};

}  // namespace bp

// Host generation of PedersenGens / BulletproofGens (src/generators.rs:47-121,196-221).
// The ChaCha20 keystream is seekable, and on secq256k1 every `Affine::rand` attempt consumes
// exactly 9 words (x is never rejected because q ~ 2^256), so attempts are
// evaluated by a thread pool and accepted in stream order; zorro's x-draw rejects with p ~ 1/2
// (data-dependent stream position) and is generated serially.
#pragma once
#include <atomic>
#include <thread>
#include "curve_host.hpp"

namespace bp {

template <class C>
struct GensHost {
    using HC = HostCurve<C>;

    static void chain_seed(const uint8_t* label, size_t llen, uint8_t seed[32]) {   // generators.rs:78-93
        std::vector<uint8_t> buf(15 + llen);
        memcpy(buf.data(), "GeneratorsChain", 15);
        memcpy(buf.data() + 15, label, llen);
        uint8_t h[64];
        sha3_512(buf.data(), buf.size(), h);
        memcpy(seed, h, 32);
    }

    // first `count` points of GeneratorsChain(label)
    static void chain(const uint8_t* label, size_t llen, size_t count, affine* out) {
        uint8_t seed[32];
        chain_seed(label, llen, seed);
        const bool fixed_stride = C::Fq::m(7) == 0xFFFFFFFFu && C::Fq::m(6) == 0xFFFFFFFFu && C::Fq::m(5) == 0xFFFFFFFFu;
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0) nthreads = 1;
        if (!fixed_stride || count < 256 || nthreads == 1) {
            ChaCha20Rng rng(seed);
            for (size_t i = 0; i < count; i++) out[i] = HC::affine_rand(rng);
            return;
        }
        // parallel attempts: attempt j reads words [9j, 9j+9). A raw x >= q (probability 2^-128)
        // would shift the stream; it is detected and handled by falling back to the serial path.
        size_t produced = 0;
        uint64_t attempt0 = 0;
        std::vector<affine> pts;
        std::vector<uint8_t> ok;
        while (produced < count) {
            size_t want = (count - produced) * 2 + 64;
            pts.assign(want, affine());
            ok.assign(want, 0);
            std::atomic<bool> irregular(false);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; t++) {
                th.emplace_back([&, t] {
                    size_t lo = want * t / nthreads, hi = want * (t + 1) / nthreads;
                    ChaCha20Rng rng(seed);
                    for (size_t j = lo; j < hi; j++) {
                        rng.seek_word((attempt0 + j) * 9);
                        uint64_t l[4];
                        for (int k = 0; k < 4; k++) l[k] = rng.next_u64();
                        if (HC::Fq::geq_m(l)) { irregular = true; return; }
                        bool greatest = (rng.next_u32() >> 31) & 1;
                        affine p;
                        if (HC::point_from_x(HC::Fq::put(l), greatest, p)) { pts[j] = p; ok[j] = 1; }
                    }
                });
            }
            for (auto& x : th) x.join();
            if (irregular) {   // astronomically unlikely; redo everything serially
                ChaCha20Rng rng(seed);
                for (size_t i = 0; i < count; i++) out[i] = HC::affine_rand(rng);
                return;
            }
            for (size_t j = 0; j < want && produced < count; j++)
                if (ok[j]) out[produced++] = pts[j];
            attempt0 += want;
        }
    }

    // BulletproofGens::new(capacity, 1): party 0, labels 'G'||LE32(0), 'H'||LE32(0)  (generators.rs:196-221)
    static void bulletproof_gens(size_t capacity, affine* G, affine* H) {
        uint8_t label[5] = {'G', 0, 0, 0, 0};
        chain(label, 5, capacity, G);
        label[0] = 'H';
        chain(label, 5, capacity, H);
    }

    // PedersenGens::default()  (generators.rs:47-66)
    static void pedersen_default(affine& B, affine& B_blinding) {
        B = HC::generator();
        uint8_t ser[65] = {0};
        HC::point_uncompressed(B, ser);
        uint8_t h[64];
        sha3_512(ser, HC::POINT_UNCOMPRESSED, h);
        ChaCha20Rng rng(h);
        B_blinding = HC::affine_rand(rng);
    }
};

}  // namespace bp

// Handle types and the per-curve function table behind the C ABI.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>
#include "host/merlin.hpp"

struct bp_ctx;
namespace bp {
struct GensDev;
struct ConstraintSystemBase;
struct Variable;

struct CurveApi {
    int (*gens_generate_host)(size_t cap, uint8_t* G, uint8_t* H, uint8_t* B, uint8_t* Bb);
    int (*gens_create)(bp_ctx*, size_t cap, GensDev** out);
    int (*gens_from_points)(bp_ctx*, const uint8_t* B, const uint8_t* Bb, const uint8_t* G, const uint8_t* H, size_t cap, GensDev** out);
    int (*pedersen_commit)(const GensDev*, const uint8_t* v, const uint8_t* blind, uint8_t* out);
    int (*challenge_scalar)(Transcript*, const char* label, uint8_t* out);
    int (*rng_scalar)(Rng*, uint8_t* out);
    int (*scalar_to_bytes)(const uint8_t* mont, uint8_t* out);
    int (*scalar_from_bytes)(const uint8_t* in, uint8_t* mont);
    int (*point_compress)(const uint8_t* xy, uint8_t* out);
    int (*point_uncompressed)(const uint8_t* xy, uint8_t* out);
    int (*point_decompress)(const uint8_t* in, uint8_t* xy);
    void* (*prover_new)(bp_ctx*, const GensDev*, Transcript*);
    void (*prover_free)(void*);
    ConstraintSystemBase* (*prover_cs)(void*);
    int (*prover_commit)(void*, const uint8_t* v, const uint8_t* blind, uint8_t* out_V, Variable* var);
    int (*prover_commit_batch)(void*, const uint8_t* v, const uint8_t* blind, size_t m, uint8_t* out_V, Variable* vars);
    int (*prover_prove)(void*, Rng*, void** out_proof);
    void* (*verifier_new)(bp_ctx*, Transcript*);
    void (*verifier_free)(void*);
    ConstraintSystemBase* (*verifier_cs)(void*);
    int (*verifier_commit)(void*, const uint8_t* V, Variable* var);
    int (*verifier_verify)(void*, const void* proof, const GensDev*);
    int (*batch_verify)(bp_ctx*, Rng*, void** verifiers, const void** proofs, size_t n, const GensDev*);
    int (*batch_verify_partial)(bp_ctx*, const uint8_t* alphas, void** verifiers, const void** proofs, size_t n, const GensDev*, uint8_t* out_xy,
                                int* out_identity);
    void (*proof_free)(void*);
    int (*proof_to_bytes)(const void*, std::vector<uint8_t>& out);
    int (*proof_from_bytes)(const uint8_t*, size_t, void** out);
    void* (*proof_clone)(const void*);
    int (*proof_field)(void*, int which, uint8_t* buf, int set);
    size_t (*proof_rounds)(const void*);
    int (*chain_circuit)(ConstraintSystemBase*, const Variable* v0, size_t n, const uint8_t* ks, const uint8_t* x0);
    int (*ipa_create_host)(bp_ctx*, Transcript*, const uint8_t* Q, const uint8_t* Gf, const uint8_t* Hf, const uint8_t* G, const uint8_t* H,
                           const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out_L, uint8_t* out_R, uint8_t* out_a, uint8_t* out_b);
    int (*ipa_verify_host)(bp_ctx*, Transcript*, size_t n, const uint8_t* L, const uint8_t* R, const uint8_t* a, const uint8_t* b, const uint8_t* Gf,
                           const uint8_t* Hf, const uint8_t* P, const uint8_t* Q, const uint8_t* G, const uint8_t* H);
    int (*rng_scalars)(Rng*, size_t n, uint8_t* out);
    int (*shuffle_gadget)(ConstraintSystemBase*, const Variable* x, const Variable* y, size_t k);
    int (*proofs_from_bytes_batch)(bp_ctx*, const uint8_t* const* bufs, const size_t* lens, size_t n, void** out, int* status);
    int (*ipa_challenges_device)(bp_ctx*, const Transcript*, uint64_t padded_n, const uint8_t* L, const uint8_t* R, size_t lg_n, uint8_t* out_u,
                                 uint8_t* out_uinv, uint8_t* out_r, int* status);
};

const CurveApi* curve_api_secq();
const CurveApi* curve_api_zorro();
const CurveApi* curve_api_c25519();
inline const CurveApi* curve_api(int curve) { return curve == 0 ? curve_api_secq() : curve == 1 ? curve_api_zorro() : curve == 2 ? curve_api_c25519() : nullptr; }
}  // namespace bp

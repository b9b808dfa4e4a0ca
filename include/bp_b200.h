/*
 * bp_b200.h -- C ABI of the B200-native hot path of ark-bulletproofs.
 *
 * The reference (FindoraNetwork/ark-bulletproofs v4.1.1) is a generic Rust crate with no
 * FFI seam; its hot path reaches arithmetic through static trait calls.
 * Each entry point below names the reference interface it replaces. All buffers are plain
 * host (or, for *_device, CUDA device) pointers owned by the caller; nothing is retained
 * after return except generator tables uploaded into the context.
 *
 * Data formats (identical to ark-ff / ark-ec in-memory values):
 *   scalar : 32 bytes, 4 x u64 little-endian limbs, Montgomery form (value * 2^256 mod r)
 *   point  : 64 bytes, affine x || y, each a base-field element in the format above;
 *            the identity is encoded as x = y = 0 (never a curve point when b != 0)
 *
 * Every function returns BP_OK (0) or a negative error code; no exceptions cross the ABI.
 */
#ifndef BP_B200_H
#define BP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bp_ctx bp_ctx;

enum bp_curve {
    BP_CURVE_SECQ256K1 = 0, /* ark_secq256k1::Affine   (tests/r1cs_secq256k1.rs:5) */
    BP_CURVE_ZORRO = 1,     /* curve::zorro::G1Affine  (src/curve/zorro/g1.rs:23-46) */
    BP_CURVE_CURVE25519 = 2 /* ark_curve25519          (tests/r1cs_curve25519.rs:4) */
};

enum bp_status {
    BP_OK = 0,
    BP_ERR_ARG = -1,      /* null pointer / bad enum */
    BP_ERR_LEN = -2,      /* length mismatch: the reference panics at msm(..).unwrap() */
    BP_ERR_POW2 = -3,     /* n not a power of two: assert at src/inner_product_proof.rs:66 */
    BP_ERR_GENS = -4,     /* R1CSError::InvalidGeneratorsLength (src/r1cs/prover.rs:499-501,577-579) */
    BP_ERR_CUDA = -5,     /* CUDA runtime failure (see bp_last_error) */
    BP_ERR_NOGPU = -6,    /* no CUDA device: there is deliberately no CPU fallback */
    BP_ERR_VERIFY = -7,   /* R1CSError::VerificationError (src/r1cs/verifier.rs:595-597) */
    BP_ERR_FORMAT = -8,   /* R1CSError::FormatError (src/r1cs/proof.rs:83-91) */
    BP_ERR_MISSING = -9,  /* R1CSError::MissingAssignment (src/r1cs/prover.rs:139,170) */
    BP_ERR_UNSUPPORTED = -10,
    BP_ERR_INTERNAL = -11 /* a C++ exception (e.g. host allocation failure) was caught at the boundary: nothing unwinds across the ABI */
};

/* ---- context ------------------------------------------------------------------------ */
/* One context = one curve on one GPU with its own stream and scratch arena. A context is
 * used by one host thread at a time. Replaces the implicit "process-wide CPU" of the crate. */
int bp_ctx_create(int curve, int device, bp_ctx** out);
void bp_ctx_destroy(bp_ctx* ctx);
const char* bp_last_error(const bp_ctx* ctx);
/* The cudaStream_t all work of this context is enqueued on (for event timing by callers). */
void* bp_ctx_stream(bp_ctx* ctx);
int bp_ctx_sync(bp_ctx* ctx);
/* Number of kernels this context launched since creation (bench.py's gpu_launches). */
uint64_t bp_ctx_launch_count(const bp_ctx* ctx);

/* ---- variable-base MSM -----------------------------------------------------------------
 * Replaces `G::Group::msm(&[G], &[G::ScalarField]) -> G::Group` followed by `.into_affine()`
 * (ark-ec VariableBaseMSM; the 17 call sites of SURVEY.md 8(a) row a1, e.g.
 * src/inner_product_proof.rs:104,124  src/r1cs/prover.rs:516-559  src/r1cs/verifier.rs:574,685).
 * out_xy receives the affine sum; *out_is_identity = 1 when the sum is the identity. */
int bp_msm(bp_ctx* ctx, const uint8_t* bases_xy, const uint8_t* scalars, size_t n, uint8_t out_xy[64],
           int* out_is_identity);
/* Same, with bases and scalars already resident in this context's GPU memory. */
int bp_msm_device(bp_ctx* ctx, const void* d_bases_xy, const void* d_scalars, size_t n, uint8_t out_xy[64],
                  int* out_is_identity);
/* Bases resident on the GPU, scalars in host memory: every large MSM of the protocol is over the fixed generator tables
 * (src/generators.rs:150-183), so a caller uploads them once (bp_bases_upload) and each bp_msm_bases call moves only the
 * 32-byte scalars -- streamed in growing chunks while the chunks already there are accumulated, like bp_msm. Computes
 * sum_i scalars[i] * bases[offset + i]. */
typedef struct bp_bases bp_bases;
int bp_bases_upload(bp_ctx* ctx, const uint8_t* bases_xy, size_t n, bp_bases** out);
void bp_bases_free(bp_bases* b);
const void* bp_bases_device_ptr(const bp_bases* b);
int bp_msm_bases(bp_ctx* ctx, const bp_bases* bases, size_t offset, const uint8_t* scalars, size_t n, uint8_t out_xy[64],
                 int* out_is_identity);
/* Per-phase device timing of the last MSM (cudaEvents on the context's stream), for the roofline
 * numbers in bench.py: phase_ms[0..4] = digits, sort, accumulate, partial reduction, bucket
 * reduction + window sums; phase_ms[5] = 1 if the pairs were ordered by the pipeline's own bucket sort (msm_sort.cuh),
 * 0 for the library radix sort; *c / *windows / *entries describe the Pippenger plan that ran. */
int bp_ctx_set_timing(bp_ctx* ctx, int enable);
int bp_msm_last_phases(const bp_ctx* ctx, float phase_ms[8], int* c, int* windows, uint64_t* entries);
/* Host wall-clock split (ms) of the last bp_prover_prove / bp_verifier_verify on this context:
 * 0 rng (TranscriptRng draws), 1 vector commitments, 2 flatten constraints, 3 l/r/t vector kernels,
 * 4 T commitments, 5 IPA total, 6 IPA MSMs, 7 IPA folds (only with timing enabled), 8 IPA host
 * (transcript, challenge inverse), 9 verification scalars, 10 mega-MSM, 11 uploads, 12 clearing of the secrets. */
int bp_ctx_last_stage_ms(const bp_ctx* ctx, double out[16]);
/* ---- multi-GPU: one context per process and GPU --------------------------------------------
 * `fn` must gather `bytes` bytes from every rank into recv[world * bytes] in rank order (an NCCL or gloo
 * all-gather issued by the host program) and return 0. After this call:
 *   - bp_gens_create keeps only the generators G_i, H_i with i = rank (mod world) on this GPU;
 *   - every MSM over them inside prove / verify / InnerProductProof::create runs on the local shard and the
 *     64-byte partial points are all-gathered and added on every rank (the reference's G::Group::msm call sites,
 *     inner_product_proof.rs:104,124,187,202; prover.rs:516-559,607-648; verifier.rs:574);
 *   - the IPA folds its shard of G and H locally (index i and its partner i + n/2 share a rank while n >= 2*world);
 *   - scalars, transcript and TranscriptRng are replicated, so all ranks return the same proof bytes / verdict.
 * world must be a power of two that divides every generator count used. */
typedef int (*bp_allgather_fn)(void* user, const void* send, void* recv, size_t bytes);
int bp_ctx_set_collective(bp_ctx* ctx, int rank, int world, bp_allgather_fn fn, void* user);
/* The same multi-GPU mode with the exchange owned by the library ("sum the per-GPU partial points with a
 * tiny NCCL all-gather over NVLink"): rank 0 calls bp_nccl_unique_id and hands the 128 bytes to every rank (any
 * out-of-band channel); every rank then calls bp_ctx_init_nccl (collective: it returns once all `world` ranks have
 * joined). From then on each sharded MSM ends with one ncclAllGather of the 64-byte partial points on the context's
 * stream -- no callback into the host program. NCCL is reached through dlopen("libnccl.so.2"), so inside a torch process
 * it is the copy torch has loaded; BP_ERR_UNSUPPORTED when no NCCL can be found. Replaces rayon's in-process reduction of
 * the reference's `parallel` feature (Cargo.toml:76) across GPUs. */
int bp_nccl_unique_id(uint8_t out[128]);
int bp_ctx_init_nccl(bp_ctx* ctx, int rank, int world, const uint8_t unique_id[128]);
/* bp_msm over host buffers of more than 1.5x `points` points (default 2^21) is streamed: the input is copied chunk by
 * chunk (first chunk points/8, each next one 1.5x larger, at most 2x points) while the kernels of the chunks already
 * on the device run, all chunks adding into one bucket array. Exposed for tests and tuning. */
int bp_msm_set_chunk(bp_ctx* ctx, size_t points);
/* IPA rounds of length n <= this threshold do not fold the generators; their L/R are MSMs over the last
 * folded stage with challenge-expanded scalars (same L, R, a, b). 0 = always fold. Default 2^13 (measured optimum with the GLV fold). */
int bp_ipa_set_nofold_threshold(bp_ctx* ctx, size_t n);
/* The R1CS prover's IPA factor vectors (src/r1cs/prover.rs:781-789) are geometric with one jump: G_f = [1]*n1 ++
 * [u]*(n2+pad), H_f[i] = y^-i * G_f[i]. Partners i and i + n/2 of a fold therefore differ by one of two scalars shared
 * by all threads, and every generator fold (first round included) is one scalar multiplication per output instead of the
 * joint double-and-add of inner_product_proof.rs:143-155. Same L, R, a, b. Default on; 0 restores the general first
 * round, which bp_ipa_create always uses for its arbitrary factor vectors (tests run both). */
int bp_ipa_set_geometric(bp_ctx* ctx, int enable);
/* secq256k1 has the endomorphism (x, y) -> (beta*x, y) = lambda*(x, y): the uniform fold scalar of each IPA round is
 * split as k1 + k2*lambda with 129-bit halves, halving the double-and-add chain of the generator fold
 * (inner_product_proof.rs:216-225). The two halves are recoded in joint sparse form, so the chain adds one of
 * {+-P, +-phi(P), +-(P + phi(P)), +-(P - phi(P))} (affine, one shared inversion per thread block) in half of its steps.
 * Same folded generators, hence the same L, R. Default 1; 2 = GLV with plain binary digits; 0 = plain 256-step fold. */
int bp_ipa_set_glv(bp_ctx* ctx, int enable);
/* bp_prover_commit_batch (m x PedersenGens::commit, src/generators.rs:39-44) uses a fixed-base table of B and B_blinding
 * (2 x 32 byte-windows x 256 multiples, built on first use per generator set): at most 64 mixed additions per
 * commitment instead of a 256-step double-and-add. Default on; 0 = the double-and-add kernel (tests compare both). */
int bp_pedersen_set_table(bp_ctx* ctx, int enable);
/* BulletproofGens::new (src/generators.rs:174-221): on secq256k1 every `G::rand` attempt of the GeneratorsChain reads
 * exactly 9 ChaCha20 words, so bp_gens_create evaluates the attempts on the GPU (seekable keystream, Tonelli-Shanks,
 * stream-order compaction) for capacities >= 256. Same points as the host generator (bp_gens_generate_host), which
 * remains the path for zorro / curve25519. Default on; 0 forces the host generator. */
int bp_gens_set_device_generation(bp_ctx* ctx, int enable);
/* Batched-affine pre-addition (csrc/msm_kernels.cuh, msm_pair_affine_kernel): before the XYZZ bucket accumulation,
 * `rounds` rounds add neighbouring same-bucket entries of the sorted pair list with the affine law, 512 additions per
 * thread sharing one field inversion (~7 instead of 10 modmul per addition). Applies to short-Weierstrass MSMs with at
 * least `min_entries` (point, window) pairs (0 keeps the current threshold, default 2^22). 0 rounds = off. */
int bp_msm_set_affine_rounds(bp_ctx* ctx, int rounds, size_t min_entries);
/* Ordering of the (bucket, point) pairs of a single large MSM (csrc/msm_sort.cuh): mode 1 (default) = the two-pass bucket
 * sort written for this pipeline (coarse bins by the top bucket bits, then one thread block per bin ranks the low bits
 * with shared-memory atomics; the order inside a bucket is free); mode 0 = cub::DeviceRadixSort. Applies to MSMs with at
 * least `min_entries` (point, window) pairs (0 keeps the current threshold, default 2^22); batched MSMs always take the
 * library sort. Same results. */
int bp_msm_set_sort(bp_ctx* ctx, int mode, size_t min_entries);
/* Two-level bucket reduction of large windows (csrc/msm_kernels.cuh, msm_reduce_windows): every 64-bucket segment
 * yields its (weighted sum, plain sum) pair and the segment offsets are applied once per slice of segments instead of by
 * a double-and-add in every segment thread (2^24 points: 4.9 -> 3.7 ms). Default on; same results. */
int bp_msm_set_two_level_reduce(bp_ctx* ctx, int enable);
/* Force the Pippenger window width (0 = automatic); for parity tests and tuning. */
int bp_msm_set_window(bp_ctx* ctx, int c);
/* MSMs with at most `max_terms` (default 768) terms each run as one kernel launch (4-bit windows, digit
 * multiples tree-summed per window) instead of the bucket pipeline, which at that size is pure latency; 0 disables
 * (a forced window also selects the bucket pipeline). Same results. */
int bp_msm_set_tiny(bp_ctx* ctx, int max_terms);

/* Sum of n affine points (host). Used to combine per-GPU partial MSM results after the
 * all-gather of SURVEY.md 8(e). */
int bp_points_sum(bp_ctx* ctx, const uint8_t* points_xy, size_t n, uint8_t out_xy[64], int* out_is_identity);
/* Same without a context (pure host arithmetic on a handful of points). */
int bp_points_sum_curve(int curve, const uint8_t* points_xy, size_t n, uint8_t out_xy[64], int* out_is_identity);

/* Fill d_out_xy (device, n*64 bytes) with the distinct points (start+i+1)*G, i < n, for
 * synthetic MSM workloads. */
int bp_synth_points_device(bp_ctx* ctx, void* d_out_xy, size_t n, uint64_t start);


/* =====================================================================================================
 * Handle-based mirror of the crate's public API. A Rust shim keeps its `Prover` / `Verifier` /
 * `BulletproofGens` types as thin wrappers around these handles.
 * ===================================================================================================== */
typedef struct bp_transcript bp_transcript; /* merlin::Transcript */
typedef struct bp_rng bp_rng;               /* rand_core::RngCore + CryptoRng */
typedef struct bp_gens bp_gens;             /* PedersenGens + BulletproofGens (party 0), resident in HBM */
typedef struct bp_cs bp_cs;                 /* trait ConstraintSystem (src/r1cs/constraint_system.rs:19-135) */
typedef struct bp_prover bp_prover;         /* r1cs::Prover   (src/r1cs/prover.rs:30-45) */
typedef struct bp_verifier bp_verifier;     /* r1cs::Verifier (src/r1cs/verifier.rs:34-58) */
typedef struct bp_proof bp_proof;           /* r1cs::R1CSProof (src/r1cs/proof.rs:25-59) */

/* r1cs::Variable (src/r1cs/linear_combination.rs:14-27) */
enum bp_var_kind { BP_VAR_COMMITTED = 0, BP_VAR_MUL_LEFT = 1, BP_VAR_MUL_RIGHT = 2, BP_VAR_MUL_OUT = 3, BP_VAR_ONE = 4 };
typedef struct { uint32_t kind; uint32_t reserved; uint64_t index; } bp_var;
/* one (Variable, coeff) term of a LinearCombination (src/r1cs/linear_combination.rs:85-88) */
typedef struct { bp_var var; uint8_t coeff[32]; } bp_term;

/* ---- merlin::Transcript (same STROBE schedule; src/transcript.rs uses append_message / challenge_bytes) */
bp_transcript* bp_transcript_new(const uint8_t* label, size_t len);
bp_transcript* bp_transcript_clone(const bp_transcript* t);
void bp_transcript_free(bp_transcript* t);
void bp_transcript_append_message(bp_transcript* t, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen);
void bp_transcript_append_u64(bp_transcript* t, const uint8_t* label, size_t llen, uint64_t v);
void bp_transcript_challenge_bytes(bp_transcript* t, const uint8_t* label, size_t llen, uint8_t* out, size_t n);
/* TranscriptProtocol::challenge_scalar (src/transcript.rs:95-101); out = Montgomery scalar */
int bp_transcript_challenge_scalar(int curve, bp_transcript* t, const uint8_t* label, size_t llen, uint8_t out[32]);

/* ---- RNGs: rand_chacha::ChaCha20Rng::from_seed, or the caller's own RngCore through callbacks */
bp_rng* bp_rng_chacha20(const uint8_t seed[32]);
bp_rng* bp_rng_from_callbacks(void* user, uint64_t (*next_u64)(void*), uint32_t (*next_u32)(void*),
                              void (*fill_bytes)(void*, uint8_t*, size_t));
void bp_rng_free(bp_rng* r);
uint64_t bp_rng_words_used(const bp_rng* r);
/* `ScalarField::rand(rng)` (ark-ff UniformRand); out = Montgomery scalar */
int bp_rng_scalar(int curve, bp_rng* r, uint8_t out[32]);
int bp_rng_scalars(int curve, bp_rng* r, size_t n, uint8_t* out); /* n successive ScalarField::rand draws */
uint64_t bp_rng_next_u64(bp_rng* r);
/* merlin `transcript.build_rng().rekey_with_witness_bytes(label, w_i)...finalize(external)` exactly as
 * Prover::prove builds its blinding RNG (src/r1cs/prover.rs:483-494); witnesses = nwit x 32 bytes. */
bp_rng* bp_transcript_build_rng(const bp_transcript* t, const uint8_t* label, size_t llen, const uint8_t* witnesses, size_t nwit,
                                bp_rng* external);
/* Host Keccak-f[1600] implementation behind the transcript and TranscriptRng: 0 = portable scalar, 1 = AVX-512
 * (default when the CPU has it), negative = query. Returns the active one, -1 if `which` is unsupported here. */
int bp_host_keccak_select(int which);

/* ---- ark-serialize canonical forms (src/transcript.rs:69-79, src/r1cs/proof.rs:74-91) */
int bp_scalar_to_bytes(int curve, const uint8_t mont[32], uint8_t out[32]);
int bp_scalar_from_bytes(int curve, const uint8_t in[32], uint8_t mont[32]);
int bp_point_compress(int curve, const uint8_t xy[64], uint8_t out[33]);
int bp_point_serialize_uncompressed(int curve, const uint8_t xy[64], uint8_t out[65]);
int bp_point_decompress(int curve, const uint8_t in[33], uint8_t xy[64]);

/* ---- generators: PedersenGens::default() + BulletproofGens::new(capacity, 1)
 * (src/generators.rs:47-66,174-221). bp_gens_generate_host needs no GPU (pure host generation);
 * bp_gens_create generates and uploads; bp_gens_from_points uploads tables the caller already has
 * (a Rust shim passes the crate's own G_vec[0] / H_vec[0]). */
int bp_gens_generate_host(int curve, size_t capacity, uint8_t* G_xy, uint8_t* H_xy, uint8_t B[64], uint8_t B_blinding[64]);
int bp_gens_create(bp_ctx* ctx, size_t capacity, bp_gens** out);
int bp_gens_from_points(bp_ctx* ctx, const uint8_t B[64], const uint8_t B_blinding[64], const uint8_t* G_xy, const uint8_t* H_xy,
                        size_t capacity, bp_gens** out);
void bp_gens_free(bp_gens* g);
size_t bp_gens_capacity(const bp_gens* g);
/* which: 0 = G, 1 = H, 2 = [B, B_blinding] */
int bp_gens_export(const bp_gens* g, int which, size_t offset, size_t count, uint8_t* out_xy);
/* PedersenGens::commit (src/generators.rs:39-44) */
int bp_pedersen_commit(const bp_gens* g, const uint8_t value[32], const uint8_t blinding[32], uint8_t out_xy[64]);

/* ---- ConstraintSystem / RandomizableConstraintSystem / RandomizedConstraintSystem
 * (src/r1cs/constraint_system.rs:19-135; prover.rs:96-268; verifier.rs:69-240).
 * Assignments are Montgomery scalars; pass NULL where the verifier passes None. */
int bp_cs_multiply(bp_cs* cs, const bp_term* left, size_t nl, const bp_term* right, size_t nr, bp_var out[3]);
int bp_cs_allocate(bp_cs* cs, const uint8_t* assignment, bp_var* out);
int bp_cs_allocate_multiplier(bp_cs* cs, const uint8_t* left, const uint8_t* right, bp_var out[3]);
int bp_cs_constrain(bp_cs* cs, const bp_term* terms, size_t n);
size_t bp_cs_multipliers_len(const bp_cs* cs);
typedef int (*bp_randomized_cb)(bp_cs* cs, void* user);
int bp_cs_specify_randomized_constraints(bp_cs* cs, bp_randomized_cb cb, void* user);
int bp_cs_challenge_scalar(bp_cs* cs, const uint8_t* label, size_t llen, uint8_t out[32]);
/* Synthetic measurement circuit: one-phase public-multiplier chain of n
 * multipliers over the committed variable v0; ks = n Montgomery scalars; x0 = the prover's witness for
 * v0 (NULL on the verifier's side). Equivalent to n allocate_multiplier + 2n constrain calls. */
int bp_cs_chain_circuit(bp_cs* cs, const bp_var* v0, size_t n, const uint8_t* ks, const uint8_t* x0);
/* The reference's k-shuffle gadget (benches/r1cs_secq256k1.rs:35-75, tests/r1cs_secq256k1.rs:16-56) over k input
 * and k output variables: registers the randomised (phase-2) constraints prod(x_i - z) == prod(y_i - z),
 * 2(k-1) multipliers. Same constraint order as the Rust gadget, so proofs are byte-identical. */
int bp_cs_shuffle_gadget(bp_cs* cs, const bp_var* x, const bp_var* y, size_t k);

/* ---- Prover::new / commit / prove (src/r1cs/prover.rs:291,327,444) */
int bp_prover_new(bp_ctx* ctx, const bp_gens* pc_gens, bp_transcript* transcript, bp_prover** out);
void bp_prover_free(bp_prover* p);
bp_cs* bp_prover_cs(bp_prover* p);
int bp_prover_commit(bp_prover* p, const uint8_t value[32], const uint8_t blinding[32], uint8_t out_commitment[64], bp_var* out_var);
/* m successive Prover::commit calls in one go: the 2m scalar multiplications run as one GPU kernel, the
 * transcript sees the same V_i in the same order (k-shuffle style circuits commit 2k inputs). Arrays are
 * m x 32 / m x 64 bytes, 16-byte aligned. */
int bp_prover_commit_batch(bp_prover* p, const uint8_t* values, const uint8_t* blindings, size_t m, uint8_t* out_commitments, bp_var* out_vars);
/* Consumes the constraint system like `Prover::prove(self, prng, bp_gens)`; bp_gens are the ones
 * given to bp_prover_new. BP_ERR_GENS when gens_capacity < padded multipliers. */
int bp_prover_prove(bp_prover* p, bp_rng* prng, bp_proof** out);

/* ---- Verifier::new / commit / verify (src/r1cs/verifier.rs:252,279,549) and batch_verify (:604) */
int bp_verifier_new(bp_ctx* ctx, bp_transcript* transcript, bp_verifier** out);
void bp_verifier_free(bp_verifier* v);
bp_cs* bp_verifier_cs(bp_verifier* v);
int bp_verifier_commit(bp_verifier* v, const uint8_t commitment[64], bp_var* out_var);
/* m successive Verifier::commit calls (src/r1cs/verifier.rs:279-291): commitments = m x 64 bytes */
int bp_verifier_commit_batch(bp_verifier* v, const uint8_t* commitments, size_t m, bp_var* out_vars);
int bp_verifier_verify(bp_verifier* v, const bp_proof* proof, const bp_gens* gens);
int bp_batch_verify(bp_ctx* ctx, bp_rng* prng, bp_verifier* const* verifiers, const bp_proof* const* proofs, size_t n, const bp_gens* gens);

/* Multi-GPU batch verification: proofs are sharded over ranks; every rank draws the
 * same alpha sequence from the caller's RNG (verifier.rs:649), passes the alphas of ITS proofs here and
 * gets its share of the final MSM (verifier.rs:685) as a point. The batch is accepted iff the sum of all
 * ranks' points (all-gather + bp_points_sum_curve) is the identity. alphas: n Montgomery scalars. */
int bp_batch_verify_partial(bp_ctx* ctx, const uint8_t* alphas, bp_verifier* const* verifiers, const bp_proof* const* proofs, size_t n,
                            const bp_gens* gens, uint8_t out_xy[64], int* out_is_identity);

/* ---- R1CSProof::{to_bytes, from_bytes} (src/r1cs/proof.rs:74-91). With out == NULL only *len is set. */
/* Device-side transcript (csrc/transcript_dev.cuh): batches of at least `min_proofs` proofs
 * (default 32; 0 = never) derive the inner-product challenges u_j (src/inner_product_proof.rs:266-277), their inverses and
 * the challenge `r` (src/r1cs/verifier.rs:516-519) of all their transcripts in one launch -- Keccak-f / STROBE / Merlin /
 * ChaCha20 / ScalarField::rand, one transcript per thread -- instead of serially on the host. Same values, same decisions. */
int bp_batch_verify_set_device_transcript(bp_ctx* ctx, int min_proofs);
/* The device transcript on one transcript, for parity tests: continues from the state of `t` (not advanced) with
 * innerproduct_domain_sep(padded_n) and the lg_n (L_j, R_j) pairs; out_u / out_u_inv = lg_n Montgomery scalars each,
 * out_r = the challenge drawn from the clone; *out_identity_seen = 1 if an L_j or R_j is the identity. */
int bp_transcript_ipa_challenges_device(bp_ctx* ctx, const bp_transcript* t, uint64_t padded_n, const uint8_t* L_xy, const uint8_t* R_xy, size_t lg_n,
                                        uint8_t* out_u, uint8_t* out_u_inv, uint8_t out_r[32], int* out_identity_seen);
void bp_proof_free(bp_proof* p);
int bp_proof_to_bytes(const bp_proof* p, uint8_t* out, size_t cap, size_t* len);
int bp_proof_from_bytes(int curve, const uint8_t* data, size_t len, bp_proof** out);
/* n x R1CSProof::from_bytes (src/r1cs/proof.rs:83-91) for batch verification: structure and scalars are parsed on the
 * host, the 11 + 2k compressed points of every proof are decompressed and validated in one GPU launch (secq256k1,
 * zorro; curve25519 uses the host path). status[i] = BP_OK or BP_ERR_FORMAT exactly as bp_proof_from_bytes decides;
 * out[i] = NULL for rejected proofs. */
int bp_proofs_from_bytes_batch(bp_ctx* ctx, const uint8_t* const* data, const size_t* lens, size_t n, bp_proof** out, int* status);
bp_proof* bp_proof_clone(const bp_proof* p);
/* Field access (tamper tests). which: 0 t_x, 1 t_x_blinding, 2 e_blinding, 3 ipp.a, 4 ipp.b (32 B, Montgomery);
 * 10..20 = A_I1,A_O1,S1,A_I2,A_O2,S2,T_1,T_3,T_4,T_5,T_6; 100+j = L_j; 200+j = R_j (64 B affine). */
int bp_proof_get_field(const bp_proof* p, int which, uint8_t* buf);
int bp_proof_set_field(bp_proof* p, int which, const uint8_t* buf);
size_t bp_proof_rounds(const bp_proof* p);

/* ---- InnerProductProof::create (src/inner_product_proof.rs:37-239) over host buffers; n a power of two.
 * out_L / out_R receive log2(n) affine points each. */
int bp_ipa_create(bp_ctx* ctx, bp_transcript* transcript, const uint8_t Q[64], const uint8_t* G_factors, const uint8_t* H_factors,
                  const uint8_t* G_xy, const uint8_t* H_xy, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out_L, uint8_t* out_R,
                  uint8_t out_a[32], uint8_t out_b[32]);
/* ---- InnerProductProof::verify (src/inner_product_proof.rs:321-382, test-only in the reference): BP_OK iff the
 * proof (L, R: log2(n) points; a, b) opens P with respect to G, H' = H*H_factors, Q. */
int bp_ipa_verify(bp_ctx* ctx, bp_transcript* transcript, size_t n, const uint8_t* L_xy, const uint8_t* R_xy, const uint8_t a[32], const uint8_t b[32],
                  const uint8_t* G_factors, const uint8_t* H_factors, const uint8_t P[64], const uint8_t Q[64], const uint8_t* G_xy,
                  const uint8_t* H_xy);

#ifdef __cplusplus
}
#endif
#endif /* BP_B200_H */

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
    BP_ERR_UNSUPPORTED = -10
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
/* Per-phase device timing of the last MSM (cudaEvents on the context's stream), for the roofline
 * numbers in bench.py: phase_ms[0..4] = digits, sort, accumulate, partial reduction, bucket
 * reduction + window sums; *c / *windows / *entries describe the Pippenger plan that ran. */
int bp_ctx_set_timing(bp_ctx* ctx, int enable);
int bp_msm_last_phases(const bp_ctx* ctx, float phase_ms[8], int* c, int* windows, uint64_t* entries);
/* Force the Pippenger window width (0 = automatic); for parity tests and tuning. */
int bp_msm_set_window(bp_ctx* ctx, int c);

/* Sum of n affine points (host). Used to combine per-GPU partial MSM results after the
 * all-gather of SURVEY.md 8(e). */
int bp_points_sum(bp_ctx* ctx, const uint8_t* points_xy, size_t n, uint8_t out_xy[64], int* out_is_identity);

/* Fill d_out_xy (device, n*64 bytes) with the distinct points (start+i+1)*G, i < n, for
 * synthetic MSM workloads. */
int bp_synth_points_device(bp_ctx* ctx, void* d_out_xy, size_t n, uint64_t start);

#ifdef __cplusplus
}
#endif
#endif /* BP_B200_H */

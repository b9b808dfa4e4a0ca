"""ctypes binding of libbp_b200.so (the C ABI in include/bp_b200.h).

There is deliberately no CPU fallback: if the CUDA library is missing or no GPU is
present, loading / context creation raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbp_b200.so")

BP_OK = 0
ERRORS = {
    -1: "BP_ERR_ARG", -2: "BP_ERR_LEN", -3: "BP_ERR_POW2", -4: "BP_ERR_GENS", -5: "BP_ERR_CUDA",
    -6: "BP_ERR_NOGPU", -7: "BP_ERR_VERIFY", -8: "BP_ERR_FORMAT", -9: "BP_ERR_MISSING", -10: "BP_ERR_UNSUPPORTED",
}

_lib = None


class BpError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d) %s" % (ERRORS.get(code, "?"), code, msg))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libbp_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "this package has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
    pi32 = ctypes.POINTER(ctypes.c_int)
    pu64 = ctypes.POINTER(u64)
    psz = ctypes.POINTER(sz)
    pvp = ctypes.POINTER(vp)
    sigs = {
        "bp_ctx_create": (i32, [i32, i32, pvp]),
        "bp_ctx_destroy": (None, [vp]),
        "bp_last_error": (ctypes.c_char_p, [vp]),
        "bp_ctx_stream": (vp, [vp]),
        "bp_ctx_sync": (i32, [vp]),
        "bp_ctx_launch_count": (u64, [vp]),
        "bp_msm": (i32, [vp, vp, vp, sz, vp, pi32]),
        "bp_msm_device": (i32, [vp, vp, vp, sz, vp, pi32]),
        "bp_bases_upload": (i32, [vp, vp, sz, pvp]),
        "bp_bases_free": (None, [vp]),
        "bp_bases_device_ptr": (vp, [vp]),
        "bp_msm_bases": (i32, [vp, vp, sz, vp, sz, vp, pi32]),
        "bp_msm_set_window": (i32, [vp, i32]),
        "bp_msm_set_tiny": (i32, [vp, i32]),
        "bp_msm_set_two_level_reduce": (i32, [vp, i32]),
        "bp_msm_set_sort": (i32, [vp, i32, sz]),
        "bp_msm_set_affine_rounds": (i32, [vp, i32, sz]),
        "bp_msm_set_chunk": (i32, [vp, sz]),
        "bp_ipa_set_nofold_threshold": (i32, [vp, sz]),
        "bp_ipa_set_geometric": (i32, [vp, i32]),
        "bp_ipa_set_glv": (i32, [vp, i32]),
        "bp_pedersen_set_table": (i32, [vp, i32]),
        "bp_gens_set_device_generation": (i32, [vp, i32]),
        "bp_ctx_set_collective": (i32, [vp, i32, i32, vp, vp]),
        "bp_batch_verify_set_device_transcript": (i32, [vp, i32]),
        "bp_transcript_ipa_challenges_device": (i32, [vp, vp, u64, vp, vp, sz, vp, vp, vp, pi32]),
        "bp_nccl_unique_id": (i32, [vp]),
        "bp_ctx_init_nccl": (i32, [vp, i32, i32, vp]),
        "bp_ctx_set_timing": (i32, [vp, i32]),
        "bp_msm_last_phases": (i32, [vp, ctypes.POINTER(ctypes.c_float), pi32, pi32, pu64]),
        "bp_points_sum": (i32, [vp, vp, sz, vp, pi32]),
        "bp_points_sum_curve": (i32, [i32, vp, sz, vp, pi32]),
        "bp_synth_points_device": (i32, [vp, vp, sz, u64]),
        "bp_transcript_new": (vp, [vp, sz]),
        "bp_transcript_clone": (vp, [vp]),
        "bp_transcript_free": (None, [vp]),
        "bp_transcript_append_message": (None, [vp, vp, sz, vp, sz]),
        "bp_transcript_append_u64": (None, [vp, vp, sz, u64]),
        "bp_transcript_challenge_bytes": (None, [vp, vp, sz, vp, sz]),
        "bp_transcript_challenge_scalar": (i32, [i32, vp, vp, sz, vp]),
        "bp_rng_chacha20": (vp, [vp]),
        "bp_rng_from_callbacks": (vp, [vp, vp, vp, vp]),
        "bp_rng_free": (None, [vp]),
        "bp_rng_words_used": (u64, [vp]),
        "bp_rng_scalar": (i32, [i32, vp, vp]),
        "bp_rng_scalars": (i32, [i32, vp, sz, vp]),
        "bp_rng_next_u64": (u64, [vp]),
        "bp_transcript_build_rng": (vp, [vp, vp, sz, vp, sz, vp]),
        "bp_host_keccak_select": (i32, [i32]),
        "bp_cs_chain_circuit": (i32, [vp, vp, sz, vp, vp]),
        "bp_cs_shuffle_gadget": (i32, [vp, vp, vp, sz]),
        "bp_ctx_last_stage_ms": (i32, [vp, ctypes.POINTER(ctypes.c_double)]),
        "bp_scalar_to_bytes": (i32, [i32, vp, vp]),
        "bp_scalar_from_bytes": (i32, [i32, vp, vp]),
        "bp_point_compress": (i32, [i32, vp, vp]),
        "bp_point_serialize_uncompressed": (i32, [i32, vp, vp]),
        "bp_point_decompress": (i32, [i32, vp, vp]),
        "bp_gens_generate_host": (i32, [i32, sz, vp, vp, vp, vp]),
        "bp_gens_create": (i32, [vp, sz, pvp]),
        "bp_gens_from_points": (i32, [vp, vp, vp, vp, vp, sz, pvp]),
        "bp_gens_free": (None, [vp]),
        "bp_gens_capacity": (sz, [vp]),
        "bp_gens_export": (i32, [vp, i32, sz, sz, vp]),
        "bp_pedersen_commit": (i32, [vp, vp, vp, vp]),
        "bp_cs_multiply": (i32, [vp, vp, sz, vp, sz, vp]),
        "bp_cs_allocate": (i32, [vp, vp, vp]),
        "bp_cs_allocate_multiplier": (i32, [vp, vp, vp, vp]),
        "bp_cs_constrain": (i32, [vp, vp, sz]),
        "bp_cs_multipliers_len": (sz, [vp]),
        "bp_cs_specify_randomized_constraints": (i32, [vp, vp, vp]),
        "bp_cs_challenge_scalar": (i32, [vp, vp, sz, vp]),
        "bp_prover_new": (i32, [vp, vp, vp, pvp]),
        "bp_prover_free": (None, [vp]),
        "bp_prover_cs": (vp, [vp]),
        "bp_prover_commit": (i32, [vp, vp, vp, vp, vp]),
        "bp_prover_commit_batch": (i32, [vp, vp, vp, sz, vp, vp]),
        "bp_prover_prove": (i32, [vp, vp, pvp]),
        "bp_verifier_new": (i32, [vp, vp, pvp]),
        "bp_verifier_free": (None, [vp]),
        "bp_verifier_cs": (vp, [vp]),
        "bp_verifier_commit": (i32, [vp, vp, vp]),
        "bp_verifier_commit_batch": (i32, [vp, vp, sz, vp]),
        "bp_verifier_verify": (i32, [vp, vp, vp]),
        "bp_batch_verify": (i32, [vp, vp, vp, vp, sz, vp]),
        "bp_batch_verify_partial": (i32, [vp, vp, vp, vp, sz, vp, vp, pi32]),
        "bp_proof_free": (None, [vp]),
        "bp_proof_to_bytes": (i32, [vp, vp, sz, psz]),
        "bp_proof_from_bytes": (i32, [i32, vp, sz, pvp]),
        "bp_proofs_from_bytes_batch": (i32, [vp, vp, vp, sz, vp, vp]),
        "bp_proof_clone": (vp, [vp]),
        "bp_proof_get_field": (i32, [vp, i32, vp]),
        "bp_proof_set_field": (i32, [vp, i32, vp]),
        "bp_proof_rounds": (sz, [vp]),
        "bp_ipa_verify": (i32, [vp, vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "bp_ipa_create": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp]),
    }
    global EXPORTED_SYMBOLS
    EXPORTED_SYMBOLS = sorted(sigs)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = []   # filled by load()

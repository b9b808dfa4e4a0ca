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
    sigs = {
        "bp_ctx_create": (i32, [i32, i32, ctypes.POINTER(vp)]),
        "bp_ctx_destroy": (None, [vp]),
        "bp_last_error": (ctypes.c_char_p, [vp]),
        "bp_ctx_stream": (vp, [vp]),
        "bp_ctx_sync": (i32, [vp]),
        "bp_ctx_launch_count": (u64, [vp]),
        "bp_msm": (i32, [vp, vp, vp, sz, vp, pi32]),
        "bp_msm_device": (i32, [vp, vp, vp, sz, vp, pi32]),
        "bp_msm_set_window": (i32, [vp, i32]),
        "bp_ctx_set_timing": (i32, [vp, i32]),
        "bp_msm_last_phases": (i32, [vp, ctypes.POINTER(ctypes.c_float), pi32, pi32, ctypes.POINTER(u64)]),
        "bp_points_sum": (i32, [vp, vp, sz, vp, pi32]),
        "bp_synth_points_device": (i32, [vp, vp, sz, u64]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "bp_ctx_create", "bp_ctx_destroy", "bp_last_error", "bp_ctx_stream", "bp_ctx_sync", "bp_ctx_launch_count",
    "bp_msm", "bp_msm_device", "bp_msm_set_window", "bp_ctx_set_timing", "bp_msm_last_phases", "bp_points_sum", "bp_synth_points_device",
]

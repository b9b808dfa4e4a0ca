"""ctypes access to oracle/c/bp_ref.c (CPU oracle / CPU baseline; test infrastructure only)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libbp_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = ctypes.CDLL(_LIB)
        _lib.ref_msm.restype = ctypes.c_int
        _lib.ref_msm.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
        _lib.ref_fold_points.restype = ctypes.c_int
        _lib.ref_fold_points.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.ref_num_threads.restype = ctypes.c_int
        _lib.ref_synth_points.restype = ctypes.c_int
        _lib.ref_synth_points.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]
    return _lib


def msm_bytes(curve_id: int, bases: bytes, scalars: bytes, n: int, threads: int = 1) -> bytes:
    out = ctypes.create_string_buffer(64)
    rc = load().ref_msm(curve_id, bases, scalars, n, out, threads)
    assert rc == 0
    return out.raw


def msm_ptr(curve_id: int, bases_ptr: int, scalars_ptr: int, n: int, threads: int = 1) -> bytes:
    out = ctypes.create_string_buffer(64)
    rc = load().ref_msm(curve_id, ctypes.c_void_p(bases_ptr), ctypes.c_void_p(scalars_ptr), n, out, threads)
    assert rc == 0
    return out.raw


def fold_points(curve_id: int, pts: bytearray, h: int, sL: bytes, sR: bytes, threads: int = 1):
    buf = (ctypes.c_char * len(pts)).from_buffer(pts)
    rc = load().ref_fold_points(curve_id, buf, h, sL, sR, threads)
    assert rc == 0


def num_threads() -> int:
    return load().ref_num_threads()


def synth_points(curve_id: int, g_xy: bytes, n: int, start: int = 0) -> bytearray:
    out = bytearray(64 * n)
    buf = (ctypes.c_char * len(out)).from_buffer(out)
    rc = load().ref_synth_points(curve_id, g_xy, buf, n, start)
    assert rc == 0
    return out

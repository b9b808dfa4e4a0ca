"""ark_bulletproofs_b200 -- B200-native hot path of ark-bulletproofs behind a C ABI.

Python is only the test/bench harness over libbp_b200.so (include/bp_b200.h); the host-side
mirror of the reference's Rust API lives in C++ inside the library."""
import ctypes

from . import _lib, codec
from ._lib import BpError


class Context:
    """bp_ctx: one curve on one GPU (include/bp_b200.h)."""

    def __init__(self, curve: str = "secq256k1", device: int = 0):
        self.lib = _lib.load()
        self.curve = curve
        h = ctypes.c_void_p()
        rc = self.lib.bp_ctx_create(codec.CURVE_IDS[curve], device, ctypes.byref(h))
        if rc != 0:
            raise BpError(rc, "bp_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.bp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BpError(rc, (self.lib.bp_last_error(self.h) or b"").decode())

    @property
    def stream_ptr(self) -> int:
        return self.lib.bp_ctx_stream(self.h)

    @property
    def launches(self) -> int:
        return self.lib.bp_ctx_launch_count(self.h)

    def sync(self):
        self._check(self.lib.bp_ctx_sync(self.h))

    def set_timing(self, enable: bool):
        self._check(self.lib.bp_ctx_set_timing(self.h, 1 if enable else 0))

    def last_phases(self):
        ph = (ctypes.c_float * 8)()
        c, w, e = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_uint64(0)
        self._check(self.lib.bp_msm_last_phases(self.h, ph, ctypes.byref(c), ctypes.byref(w), ctypes.byref(e)))
        names = ["digits", "sort", "accumulate", "partials", "reduce"]
        return {"ms": {n: ph[i] for i, n in enumerate(names)}, "c": c.value, "windows": w.value, "entries": e.value,
                "bucket_sort": ph[5] == 1.0}

    STAGES = ["rng", "commit", "flatten", "vec", "t_commit", "ipa", "ipa_msm", "ipa_fold", "ipa_host", "verify_scalars", "verify_msm", "upload", "tail"]

    def last_stage_ms(self):
        out = (ctypes.c_double * 16)()
        self._check(self.lib.bp_ctx_last_stage_ms(self.h, out))
        return {n: round(out[i], 4) for i, n in enumerate(self.STAGES)}

    def set_ipa_nofold_threshold(self, n: int):
        self._check(self.lib.bp_ipa_set_nofold_threshold(self.h, n))

    def set_collective(self, rank: int, world: int, allgather):
        """Multi-GPU mode (bp_ctx_set_collective): `allgather(send: bytes) -> bytes` must return the concatenation
        of every rank's `send` in rank order (ark_bulletproofs_b200.dist has torch.distributed and in-process
        implementations). Call before creating generators."""
        cb_t = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

        def tramp(_user, send, recv, nbytes):
            try:
                out = allgather(ctypes.string_at(send, nbytes))
                if len(out) != nbytes * world:
                    return 1
                ctypes.memmove(recv, out, len(out))
                return 0
            except Exception:       # never unwind into C
                import traceback
                traceback.print_exc()
                return 1
        self._coll_cb = cb_t(tramp)     # keep alive
        self.rank, self.world = rank, world
        self._check(self.lib.bp_ctx_set_collective(self.h, rank, world, ctypes.cast(self._coll_cb, ctypes.c_void_p), None))

    def init_nccl(self, rank: int, world: int, group=None):
        """Multi-GPU mode with the exchange owned by the library (bp_ctx_init_nccl): the 128-byte NCCL unique id is
        made on rank 0 and broadcast through torch.distributed (host-side, once); afterwards every sharded MSM ends with
        an ncclAllGather issued by the library on this context's stream. Call before creating generators."""
        import torch.distributed as dist
        box = [None]
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            self._check(self.lib.bp_nccl_unique_id(buf))
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=0, group=group)
        self._check(self.lib.bp_ctx_init_nccl(self.h, rank, world, box[0]))
        self.rank, self.world = rank, world

    def set_device_gens(self, enable: bool):
        self._check(self.lib.bp_gens_set_device_generation(self.h, 1 if enable else 0))

    def set_pedersen_table(self, enable: bool):
        self._check(self.lib.bp_pedersen_set_table(self.h, 1 if enable else 0))

    def set_ipa_glv(self, enable):
        """0 = plain 256-step fold, 1 / True = GLV with joint-sparse-form digits (default), 2 = GLV with binary digits"""
        self._check(self.lib.bp_ipa_set_glv(self.h, int(enable)))

    def set_ipa_geometric(self, enable: bool):
        self._check(self.lib.bp_ipa_set_geometric(self.h, 1 if enable else 0))

    def set_chunk(self, points: int):
        self._check(self.lib.bp_msm_set_chunk(self.h, points))

    def set_device_transcript(self, min_proofs: int):
        self._check(self.lib.bp_batch_verify_set_device_transcript(self.h, min_proofs))

    def set_affine_rounds(self, rounds: int, min_entries: int = 0):
        self._check(self.lib.bp_msm_set_affine_rounds(self.h, rounds, min_entries))

    def set_sort(self, mode: int, min_entries: int = 0):
        """1 = the pipeline's own two-pass bucket sort (default), 0 = cub::DeviceRadixSort"""
        self._check(self.lib.bp_msm_set_sort(self.h, mode, min_entries))

    def set_two_level_reduce(self, enable: bool):
        self._check(self.lib.bp_msm_set_two_level_reduce(self.h, int(enable)))

    def set_tiny(self, max_terms: int):
        self._check(self.lib.bp_msm_set_tiny(self.h, max_terms))

    def set_window(self, c: int):
        self._check(self.lib.bp_msm_set_window(self.h, c))

    def msm_bytes(self, bases: bytes, scalars: bytes, n: int):
        """bp_msm over host buffers; returns (64-byte affine, is_identity)."""
        out = ctypes.create_string_buffer(64)
        ident = ctypes.c_int(0)
        self._check(self.lib.bp_msm(self.h, bases, scalars, n, out, ctypes.byref(ident)))
        return out.raw, bool(ident.value)

    def msm(self, points, scalars):
        """Sum s_i * P_i for Python-int affine points / scalars (test helper)."""
        assert len(points) == len(scalars)
        raw, ident = self.msm_bytes(codec.enc_points(points, self.curve), codec.enc_scalars(scalars, self.curve), len(points))
        return None if ident else codec.dec_point(raw, self.curve)

    def msm_device(self, d_bases: int, d_scalars: int, n: int):
        out = ctypes.create_string_buffer(64)
        ident = ctypes.c_int(0)
        self._check(self.lib.bp_msm_device(self.h, d_bases, d_scalars, n, out, ctypes.byref(ident)))
        return out.raw, bool(ident.value)

    def bases_upload(self, bases_ptr_or_bytes, n: int):
        """bp_bases_upload: returns an opaque handle (free with bases_free)."""
        h = ctypes.c_void_p()
        self._check(self.lib.bp_bases_upload(self.h, bases_ptr_or_bytes, n, ctypes.byref(h)))
        return h

    def bases_free(self, h):
        self.lib.bp_bases_free(h)

    def msm_bases(self, h, scalars_ptr_or_bytes, n: int, offset: int = 0):
        """bp_msm_bases: bases resident on the GPU, scalars from host memory; returns (64-byte affine, is_identity)."""
        out = ctypes.create_string_buffer(64)
        ident = ctypes.c_int(0)
        self._check(self.lib.bp_msm_bases(self.h, h, offset, scalars_ptr_or_bytes, n, out, ctypes.byref(ident)))
        return out.raw, bool(ident.value)

    def points_sum(self, points):
        out = ctypes.create_string_buffer(64)
        ident = ctypes.c_int(0)
        self._check(self.lib.bp_points_sum(self.h, codec.enc_points(points, self.curve), len(points), out, ctypes.byref(ident)))
        return None if ident.value else codec.dec_point(out.raw, self.curve)

    def synth_points_device(self, d_out: int, n: int, start: int = 0):
        self._check(self.lib.bp_synth_points_device(self.h, d_out, n, start))

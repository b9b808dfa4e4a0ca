"""Accelerated primitives for the CPU oracle (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

oracle/bp_oracle.py restates the protocol in pure Python; at 2^16..2^20 multipliers its three O(n) primitives are out
of reach of an interpreter (8n Keccak-f for the TranscriptRng, the MSMs, the per-element generator fold). install()
swaps exactly those three for the C restatement (oracle/c/bp_ref.c: ark's Pippenger, the 2-point-msm fold, merlin's
TranscriptRng over STROBE/Keccak) -- the protocol schedule, transcript, padding and serialisation stay the Python
statements that cite the reference line by line. tests/test_c_oracle.py checks each swapped primitive against its
Python original; tests/golden/make_golden_large.py uses this to emit tests/golden/large.json.
"""
import ctypes
import os

import bp_oracle as O
import c_oracle

R256 = 1 << 256
_CURVE_ID = {"secq256k1": 0, "zorro": 1}
_installed = False
_orig = {}


def _threads():
    return os.cpu_count() or 1


def _enc_fe(v, m):
    return (v % m * R256 % m).to_bytes(32, "little")


def _enc_points(cv, pts):
    q = cv.q
    out = bytearray(64 * len(pts))
    for i, P in enumerate(pts):
        if P is not None:
            out[64 * i:64 * i + 32] = _enc_fe(P[0], q)
            out[64 * i + 32:64 * i + 64] = _enc_fe(P[1], q)
    return out


def _dec_points(cv, raw, n):
    q = cv.q
    rinv = pow(R256, -1, q)
    out = []
    for i in range(n):
        x = int.from_bytes(raw[64 * i:64 * i + 32], "little") * rinv % q
        y = int.from_bytes(raw[64 * i + 32:64 * i + 64], "little") * rinv % q
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


def _enc_scalars(cv, sc):
    r = cv.r
    return b"".join(_enc_fe(s, r) for s in sc)


def msm(cv, points, scalars):
    """G::Group::msm(...).into_affine() through ref_msm (ark-ec 0.4 msm_bigint_wnaf restated in C)."""
    cid = _CURVE_ID.get(cv.name)
    if cid is None or len(points) < 16:
        return _orig["msm"](cv, points, scalars)
    assert len(points) == len(scalars)
    raw = c_oracle.msm_bytes(cid, bytes(_enc_points(cv, points)), _enc_scalars(cv, scalars), len(points), _threads())
    return _dec_points(cv, raw, 1)[0]


def fold_generators(cv, L, R, sL, sR):
    cid = _CURVE_ID.get(cv.name)
    n = len(L)
    if cid is None or n < 16:
        return _orig["fold_generators"](cv, L, R, sL, sR)
    lib = c_oracle.load()
    out = ctypes.create_string_buffer(64 * n)
    rc = lib.ref_fold_points_v(cid, bytes(_enc_points(cv, L)), bytes(_enc_points(cv, R)), ctypes.c_size_t(n),
                               _enc_scalars(cv, sL), _enc_scalars(cv, sR), out, _threads())
    assert rc == 0
    return _dec_points(cv, out.raw, n)


def trng_scalars(cv_r, rng, count):
    """`count` draws of Fr::rand from a merlin TranscriptRng (bp_oracle.TranscriptRng), advancing its STROBE state."""
    st = rng.strobe
    lib = c_oracle.load()
    state = (ctypes.c_uint8 * 200).from_buffer_copy(bytes(st.state))
    io = (ctypes.c_int * 3)(st.pos, st.pos_begin, st.cur_flags)
    out = ctypes.create_string_buffer(32 * count)
    rc = lib.ref_trng_scalars(state, io, cv_r.to_bytes(32, "little"), ctypes.c_size_t(count), out)
    assert rc == 0
    st.state[:] = bytes(state)
    st.pos, st.pos_begin, st.cur_flags = io[0], io[1], io[2]
    rinv = pow(R256, -1, cv_r)
    raw = out.raw
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") * rinv % cv_r for i in range(count)]


def scalar_rand(cv, rng):
    if isinstance(rng, O.TranscriptRng):
        return trng_scalars(cv.r, rng, 1)[0]
    return _orig["scalar_rand"](cv, rng)


def _commit(self, value, blinding):            # generators.rs:39-44 as one 2-point msm in C
    cid = _CURVE_ID.get(self.cv.name)
    if cid is None:
        return _orig["commit"](self, value, blinding)
    raw = c_oracle.msm_bytes(cid, bytes(_enc_points(self.cv, [self.B, self.B_blinding])), _enc_scalars(self.cv, [value, blinding]), 2, 1)
    return _dec_points(self.cv, raw, 1)[0]


def install():
    global _installed
    if _installed:
        return
    lib = c_oracle.load()
    lib.ref_fold_points_v.restype = ctypes.c_int
    lib.ref_fold_points_v.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p,
                                      ctypes.c_void_p, ctypes.c_int]
    lib.ref_trng_scalars.restype = ctypes.c_int
    lib.ref_trng_scalars.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    _orig.update(msm=O.msm, fold_generators=O.fold_generators, scalar_rand=O.scalar_rand, commit=O.PedersenGens.commit)
    O.msm, O.fold_generators, O.scalar_rand = msm, fold_generators, scalar_rand
    O.PedersenGens.commit = _commit
    _installed = True


def uninstall():
    global _installed
    if not _installed:
        return
    O.msm, O.fold_generators, O.scalar_rand = _orig["msm"], _orig["fold_generators"], _orig["scalar_rand"]
    O.PedersenGens.commit = _orig["commit"]
    _installed = False


# ---- BulletproofGens for 2^20 generators -------------------------------------------------------------------------
# bp_oracle.BulletproofGens is a serial loop of Affine::rand; one attempt costs ~1 ms of Python modular exponentiation
# (Tonelli-Shanks, 2-adicity 6), i.e. over an hour for 2 x 2^20 generators. On secq256k1 every attempt of the
# GeneratorsChain reads exactly nine ChaCha20 words (the x draw is rejected with probability 2^-128), so attempt j can
# be evaluated from keystream position 9j by any worker; the accepted points, in stream order, are the chain.
# Each worker runs the oracle's own statements (fp_rand, sqrt_mod, the `greatest` rule) and asserts the nine-word
# invariant; tests/test_c_oracle.py checks the result against the serial BulletproofGens.
def _attempt_range(args):
    name, label, j0, j1 = args
    import hashlib
    cv = O.CURVES[name]
    seed = hashlib.sha3_512(b"GeneratorsChain" + label).digest()[:32]
    rng = O.ChaCha20Rng(seed)
    rng.counter = (9 * j0) // 16
    for _ in range((9 * j0) % 16):
        rng.next_u32()
    used0 = rng.words_used
    q = cv.q
    out = []
    for j in range(j0, j1):
        c0 = O.fp_rand(q, rng)                                   # bp_oracle.affine_rand, SW branch, one attempt
        greatest = (rng.next_u32() >> 31) & 1
        y = O.sqrt_mod((c0 * c0 * c0 + cv.a * c0 + cv.b) % q, q)
        if y is not None:
            lo, hi = sorted((y, (-y) % q))
            out.append((c0, hi if greatest else lo))
    assert rng.words_used - used0 == 9 * (j1 - j0), "an x draw was rejected: the nine-word invariant does not hold"
    return out


def parallel_gens(cv, capacity, procs=None):
    """bp_oracle.BulletproofGens(cv, capacity, 1) computed by a process pool (secq256k1 only)."""
    import multiprocessing as mp
    import struct
    assert cv.name == "secq256k1"
    procs = procs or (os.cpu_count() or 1)
    bp = O.BulletproofGens(cv, 0, 1)
    with mp.Pool(procs) as pool:
        for label, vec in ((b"G" + struct.pack("<I", 0), bp.G_vec[0]), (b"H" + struct.pack("<I", 0), bp.H_vec[0])):
            done = 0
            while len(vec) < capacity:
                need = capacity - len(vec)
                total = int(need * 2.1) + 64
                step = max(64, total // (procs * 8))
                jobs = [(cv.name, label, j, min(j + step, done + total)) for j in range(done, done + total, step)]
                for part in pool.map(_attempt_range, jobs):
                    vec.extend(part)
                done += total
            del vec[capacity:]
    bp.gens_capacity = capacity
    bp._chains = None      # not extendable (increase_capacity would restart the chains); the large goldens never extend
    return bp

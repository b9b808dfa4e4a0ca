"""ctypes mirror of the crate's public API over the C ABI (include/bp_b200.h):
Transcript, ChaChaRng, BulletproofGens/PedersenGens (Gens), Prover, Verifier, R1CSProof,
batch_verify, and the LinearCombination sugar of src/r1cs/linear_combination.rs.
All arithmetic happens inside libbp_b200.so (C++ host + CUDA); Python only marshals."""
import ctypes

from . import _lib, codec
from ._lib import BpError

COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = 0, 1, 2, 3, 4


class BpVar(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("index", ctypes.c_uint64)]


class BpTerm(ctypes.Structure):
    _fields_ = [("var", BpVar), ("coeff", ctypes.c_uint8 * 32)]


RAND_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)


def _chk(rc, what=""):
    if rc != 0:
        raise BpError(rc, what)


class Variable(tuple):
    """(kind, index)"""

    def __new__(cls, kind, index):
        return tuple.__new__(cls, (kind, index))

    def __add__(self, o): return LC.of(self) + o
    def __sub__(self, o): return LC.of(self) - o
    def __neg__(self): return -LC.of(self)
    def __mul__(self, k): return LC([(self, k)])


def one():
    return Variable(ONE, 0)


class LC:
    """LinearCombination: list of (Variable, int coeff) (linear_combination.rs:85-163)."""

    def __init__(self, terms=None):
        self.terms = list(terms or [])

    @staticmethod
    def of(x):
        if isinstance(x, LC):
            return LC(x.terms)
        if isinstance(x, Variable):
            return LC([(x, 1)])
        return LC([(one(), int(x))])

    def __add__(self, o): return LC(self.terms + LC.of(o).terms)
    def __sub__(self, o): return LC(self.terms + [(v, -c) for v, c in LC.of(o).terms])
    def __neg__(self): return LC([(v, -c) for v, c in self.terms])
    def __mul__(self, k): return LC([(v, c * k) for v, c in self.terms])

    def pack(self, curve):
        r = codec.MODULI[curve][1]
        arr = (BpTerm * max(1, len(self.terms)))()
        for i, (v, c) in enumerate(self.terms):
            arr[i].var.kind, arr[i].var.index = v[0], v[1]
            arr[i].coeff[:] = codec.enc_fe(c % r, r)
        return arr, len(self.terms)


class Transcript:
    def __init__(self, label: bytes = b"", _h=None):
        self.lib = _lib.load()
        self.h = _h if _h is not None else self.lib.bp_transcript_new(label, len(label))

    def clone(self):
        return Transcript(_h=self.lib.bp_transcript_clone(self.h))

    def append_message(self, label: bytes, msg: bytes):
        self.lib.bp_transcript_append_message(self.h, label, len(label), msg, len(msg))

    def append_u64(self, label: bytes, v: int):
        self.lib.bp_transcript_append_u64(self.h, label, len(label), v)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(n)
        self.lib.bp_transcript_challenge_bytes(self.h, label, len(label), out, n)
        return out.raw

    def challenge_scalar(self, curve: str, label: bytes) -> int:
        out = ctypes.create_string_buffer(32)
        _chk(self.lib.bp_transcript_challenge_scalar(codec.CURVE_IDS[curve], self.h, label, len(label), out))
        return codec.dec_fe(out.raw, codec.MODULI[curve][1])

    def build_rng(self, label: bytes, witnesses, external: "ChaChaRng") -> "ChaChaRng":
        """merlin TranscriptRng keyed as Prover::prove does (prover.rs:483-494); witnesses = 32-byte strings."""
        w = b"".join(witnesses)
        h = self.lib.bp_transcript_build_rng(self.h, label, len(label), w, len(witnesses), external.h)
        if not h:
            raise ValueError("bp_transcript_build_rng failed")
        return ChaChaRng(None, _h=h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bp_transcript_free(self.h)
            self.h = None


class ChaChaRng:
    def __init__(self, seed: bytes, _h=None):
        self.lib = _lib.load()
        self.h = _h if _h is not None else self.lib.bp_rng_chacha20(seed)

    def next_u64(self) -> int:
        return self.lib.bp_rng_next_u64(self.h)

    @property
    def words_used(self):
        return self.lib.bp_rng_words_used(self.h)

    def scalar(self, curve: str) -> int:
        out = ctypes.create_string_buffer(32)
        _chk(self.lib.bp_rng_scalar(codec.CURVE_IDS[curve], self.h, out))
        return codec.dec_fe(out.raw, codec.MODULI[curve][1])

    def scalars_raw(self, curve: str, n: int) -> bytes:
        """n successive ScalarField::rand draws as Montgomery bytes (32 B each)."""
        out = ctypes.create_string_buffer(32 * max(n, 1))
        _chk(self.lib.bp_rng_scalars(codec.CURVE_IDS[curve], self.h, n, out))
        return out.raw[:32 * n]

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bp_rng_free(self.h)
            self.h = None


def generate_gens_host(curve: str, capacity: int):
    """Pure-host generator generation (no GPU): returns (B, B_blinding, [G], [H]) as affine ints."""
    lib = _lib.load()
    G = ctypes.create_string_buffer(64 * max(capacity, 1))
    H = ctypes.create_string_buffer(64 * max(capacity, 1))
    B = ctypes.create_string_buffer(64)
    Bb = ctypes.create_string_buffer(64)
    _chk(lib.bp_gens_generate_host(codec.CURVE_IDS[curve], capacity, G, H, B, Bb))
    dp = lambda raw, i: codec.dec_point(raw[64 * i:64 * i + 64], curve)
    return dp(B.raw, 0), dp(Bb.raw, 0), [dp(G.raw, i) for i in range(capacity)], [dp(H.raw, i) for i in range(capacity)]


class Gens:
    """PedersenGens::default() + BulletproofGens::new(capacity, 1), resident on the GPU."""

    def __init__(self, ctx, capacity: int):
        self.ctx, self.lib = ctx, ctx.lib
        h = ctypes.c_void_p()
        ctx._check(self.lib.bp_gens_create(ctx.h, capacity, ctypes.byref(h)))
        self.h = h
        self.capacity = capacity

    def export(self, which: int, off: int, cnt: int):
        out = ctypes.create_string_buffer(64 * cnt)
        self.ctx._check(self.lib.bp_gens_export(self.h, which, off, cnt, out))
        return [codec.dec_point(out.raw[64 * i:64 * i + 64], self.ctx.curve) for i in range(cnt)]

    def commit(self, v: int, blinding: int):
        r = codec.MODULI[self.ctx.curve][1]
        out = ctypes.create_string_buffer(64)
        _chk(self.lib.bp_pedersen_commit(self.h, codec.enc_fe(v, r), codec.enc_fe(blinding, r), out))
        return codec.dec_point(out.raw, self.ctx.curve)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bp_gens_free(self.h)
            self.h = None


class _CS:
    """ConstraintSystem methods shared by Prover / Verifier / the randomised phase."""

    def __init__(self, lib, cs_handle, curve, is_prover):
        self.lib, self.cs, self.curve, self.is_prover = lib, cs_handle, curve, is_prover
        self.r = codec.MODULI[curve][1]
        self._keep = []

    def _vars3(self):
        return (BpVar * 3)()

    def multiply(self, left, right):
        la, ln = LC.of(left).pack(self.curve)
        ra, rn = LC.of(right).pack(self.curve)
        out = self._vars3()
        _chk(self.lib.bp_cs_multiply(self.cs, la, ln, ra, rn, out), "multiply")
        return tuple(Variable(o.kind, o.index) for o in out)

    def allocate(self, assignment=None):
        out = BpVar()
        a = None if assignment is None else codec.enc_fe(assignment, self.r)
        _chk(self.lib.bp_cs_allocate(self.cs, a, ctypes.byref(out)), "allocate")
        return Variable(out.kind, out.index)

    def allocate_multiplier(self, assignments=None):
        out = self._vars3()
        l = r = None
        if assignments is not None:
            l, r = codec.enc_fe(assignments[0], self.r), codec.enc_fe(assignments[1], self.r)
        _chk(self.lib.bp_cs_allocate_multiplier(self.cs, l, r, out), "allocate_multiplier")
        return tuple(Variable(o.kind, o.index) for o in out)

    def constrain(self, lc):
        arr, n = LC.of(lc).pack(self.curve)
        _chk(self.lib.bp_cs_constrain(self.cs, arr, n), "constrain")

    def multipliers_len(self):
        return self.lib.bp_cs_multipliers_len(self.cs)

    def challenge_scalar(self, label: bytes) -> int:
        out = ctypes.create_string_buffer(32)
        _chk(self.lib.bp_cs_challenge_scalar(self.cs, label, len(label), out), "challenge_scalar")
        return codec.dec_fe(out.raw, self.r)

    def chain_circuit_raw(self, v0, n: int, ks_raw: bytes, x0_raw=None):
        """Native builder of the synthetic chain circuit (bp_cs_chain_circuit)."""
        var = BpVar(v0[0], 0, v0[1])
        _chk(self.lib.bp_cs_chain_circuit(self.cs, ctypes.byref(var), n, ks_raw, x0_raw), "chain_circuit")

    def shuffle_gadget_native(self, xs, ys):
        """The reference's k-shuffle gadget built inside the library (bp_cs_shuffle_gadget); xs, ys: Variables."""
        k = len(xs)
        ax, ay = (BpVar * k)(), (BpVar * k)()
        for i in range(k):
            ax[i].kind, ax[i].index = xs[i][0], xs[i][1]
            ay[i].kind, ay[i].index = ys[i][0], ys[i][1]
        _chk(self.lib.bp_cs_shuffle_gadget(self.cs, ax, ay, k), "shuffle_gadget")

    def specify_randomized_constraints(self, fn):
        outer = self

        def tramp(cs_ptr, _user):
            try:
                fn(_CS(outer.lib, cs_ptr, outer.curve, outer.is_prover))
                return 0
            except BpError as e:
                return e.code
        cb = RAND_CB(tramp)
        self._keep.append(cb)
        _chk(self.lib.bp_cs_specify_randomized_constraints(self.cs, cb, None))


class Proof:
    def __init__(self, lib, h, curve):
        self.lib, self.h, self.curve = lib, h, curve

    def to_bytes(self) -> bytes:
        n = ctypes.c_size_t(0)
        _chk(self.lib.bp_proof_to_bytes(self.h, None, 0, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(n.value)
        _chk(self.lib.bp_proof_to_bytes(self.h, buf, n.value, ctypes.byref(n)))
        return buf.raw

    @staticmethod
    def from_bytes(curve: str, data: bytes):
        lib = _lib.load()
        h = ctypes.c_void_p()
        _chk(lib.bp_proof_from_bytes(codec.CURVE_IDS[curve], data, len(data), ctypes.byref(h)), "from_bytes")
        return Proof(lib, h, curve)

    @staticmethod
    def from_bytes_batch(ctx, blobs):
        """Many proofs at once with GPU point decompression (bp_proofs_from_bytes_batch): a list with a Proof, or
        None where from_bytes would raise a format error."""
        n = len(blobs)
        bufs = [ctypes.create_string_buffer(b, len(b)) for b in blobs]
        ptrs = (ctypes.c_void_p * max(n, 1))(*[ctypes.addressof(b) for b in bufs])
        lens = (ctypes.c_size_t * max(n, 1))(*[len(b) for b in blobs])
        out = (ctypes.c_void_p * max(n, 1))()
        status = (ctypes.c_int * max(n, 1))()
        ctx._check(ctx.lib.bp_proofs_from_bytes_batch(ctx.h, ptrs, lens, n, out, status))
        return [Proof(ctx.lib, ctypes.c_void_p(out[i]), ctx.curve) if status[i] == 0 else None for i in range(n)]

    def clone(self):
        return Proof(self.lib, ctypes.c_void_p(self.lib.bp_proof_clone(self.h)), self.curve)

    def rounds(self):
        return self.lib.bp_proof_rounds(self.h)

    def get_scalar(self, which):
        buf = ctypes.create_string_buffer(32)
        _chk(self.lib.bp_proof_get_field(self.h, which, buf))
        return codec.dec_fe(buf.raw, codec.MODULI[self.curve][1])

    def set_scalar(self, which, v):
        _chk(self.lib.bp_proof_set_field(self.h, which, codec.enc_fe(v, codec.MODULI[self.curve][1])))

    def get_point(self, which):
        buf = ctypes.create_string_buffer(64)
        _chk(self.lib.bp_proof_get_field(self.h, which, buf))
        return codec.dec_point(buf.raw, self.curve)

    def set_point(self, which, P):
        _chk(self.lib.bp_proof_set_field(self.h, which, codec.enc_point(P, self.curve)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bp_proof_free(self.h)
            self.h = None


class Prover(_CS):
    def __init__(self, ctx, gens: Gens, transcript: Transcript):
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.bp_prover_new(ctx.h, gens.h, transcript.h, ctypes.byref(h)))
        self.ctx, self.h, self.gens, self.transcript = ctx, h, gens, transcript
        super().__init__(ctx.lib, ctx.lib.bp_prover_cs(h), ctx.curve, True)

    def commit(self, v: int, v_blinding: int):
        out = ctypes.create_string_buffer(64)
        var = BpVar()
        _chk(self.lib.bp_prover_commit(self.h, codec.enc_fe(v, self.r), codec.enc_fe(v_blinding, self.r), out, ctypes.byref(var)))
        return codec.dec_point(out.raw, self.curve), Variable(var.kind, var.index)

    def commit_batch(self, vals, blindings):
        """m commits in one call (GPU scalar multiplications); returns ([V], [Variable])."""
        m = len(vals)
        import numpy as np
        vb = np.frombuffer(codec.enc_scalars(vals, self.curve), dtype=np.uint8).copy()
        bb = np.frombuffer(codec.enc_scalars(blindings, self.curve), dtype=np.uint8).copy()
        out = np.zeros(64 * max(m, 1), dtype=np.uint8)
        vars_ = (BpVar * max(m, 1))()
        _chk(self.lib.bp_prover_commit_batch(self.h, vb.ctypes.data, bb.ctypes.data, m, out.ctypes.data, vars_), "commit_batch")
        raw = out.tobytes()
        return ([codec.dec_point(raw[64 * i:64 * i + 64], self.curve) for i in range(m)], [Variable(vars_[i].kind, vars_[i].index) for i in range(m)])

    def commit_batch_raw(self, vals_raw: bytes, blindings_raw: bytes, m: int):
        """commit_batch over Montgomery byte strings; returns (m x 64 raw commitment bytes, [Variable])."""
        out = ctypes.create_string_buffer(64 * max(m, 1))
        vars_ = (BpVar * max(m, 1))()
        _chk(self.lib.bp_prover_commit_batch(self.h, vals_raw, blindings_raw, m, out, vars_), "commit_batch")
        return out.raw[:64 * m], [Variable(vars_[i].kind, vars_[i].index) for i in range(m)]

    def prove(self, rng: ChaChaRng) -> Proof:
        ph = ctypes.c_void_p()
        self.ctx._check(self.lib.bp_prover_prove(self.h, rng.h, ctypes.byref(ph)))
        return Proof(self.lib, ph, self.curve)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bp_prover_free(self.h)
            self.h = None


class Verifier(_CS):
    def __init__(self, ctx, transcript: Transcript):
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.bp_verifier_new(ctx.h, transcript.h, ctypes.byref(h)))
        self.ctx, self.h, self.transcript = ctx, h, transcript
        super().__init__(ctx.lib, ctx.lib.bp_verifier_cs(h), ctx.curve, False)

    def commit(self, V):
        var = BpVar()
        _chk(self.lib.bp_verifier_commit(self.h, codec.enc_point(V, self.curve), ctypes.byref(var)))
        return Variable(var.kind, var.index)

    def commit_batch_raw(self, commitments_raw: bytes, m: int):
        vars_ = (BpVar * max(m, 1))()
        _chk(self.lib.bp_verifier_commit_batch(self.h, commitments_raw, m, vars_), "verifier commit_batch")
        return [Variable(vars_[i].kind, vars_[i].index) for i in range(m)]

    def verify(self, proof: Proof, gens: Gens):
        self.ctx._check(self.lib.bp_verifier_verify(self.h, proof.h, gens.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bp_verifier_free(self.h)
            self.h = None


def batch_verify(ctx, rng: ChaChaRng, instances, gens: Gens):
    """instances: list of (Verifier, Proof)  (verifier.rs:604)."""
    n = len(instances)
    vs = (ctypes.c_void_p * max(n, 1))(*[v.h for v, _ in instances])
    ps = (ctypes.c_void_p * max(n, 1))(*[p.h for _, p in instances])
    ctx._check(ctx.lib.bp_batch_verify(ctx.h, rng.h, vs, ps, n, gens.h))


def batch_verify_partial(ctx, alphas, instances, gens: Gens):
    """This rank's share of a sharded batch verification: returns the partial MSM point (None = identity).
    alphas: the Python-int alpha of each instance (drawn by the caller in global proof order)."""
    n = len(instances)
    vs = (ctypes.c_void_p * max(n, 1))(*[v.h for v, _ in instances])
    ps = (ctypes.c_void_p * max(n, 1))(*[p.h for _, p in instances])
    out = ctypes.create_string_buffer(64)
    idn = ctypes.c_int(0)
    ctx._check(ctx.lib.bp_batch_verify_partial(ctx.h, codec.enc_scalars(alphas, ctx.curve), vs, ps, n, gens.h, out, ctypes.byref(idn)))
    return None if idn.value else codec.dec_point(out.raw, ctx.curve)


def ipa_create(ctx, transcript: Transcript, Q, G_factors, H_factors, G, H, a, b):
    """InnerProductProof::create over Python-int inputs; returns (L_vec, R_vec, a, b)."""
    curve = ctx.curve
    n = len(G)
    k = max(n.bit_length() - 1, 0)
    oL = ctypes.create_string_buffer(64 * max(k, 1))
    oR = ctypes.create_string_buffer(64 * max(k, 1))
    oa = ctypes.create_string_buffer(32)
    ob = ctypes.create_string_buffer(32)
    ctx._check(ctx.lib.bp_ipa_create(ctx.h, transcript.h, codec.enc_point(Q, curve), codec.enc_scalars(G_factors, curve),
                                     codec.enc_scalars(H_factors, curve), codec.enc_points(G, curve), codec.enc_points(H, curve),
                                     codec.enc_scalars(a, curve), codec.enc_scalars(b, curve), n, oL, oR, oa, ob))
    r = codec.MODULI[curve][1]
    L = [codec.dec_point(oL.raw[64 * i:64 * i + 64], curve) for i in range(k)]
    R = [codec.dec_point(oR.raw[64 * i:64 * i + 64], curve) for i in range(k)]
    return L, R, codec.dec_fe(oa.raw, r), codec.dec_fe(ob.raw, r)


def ipa_verify(ctx, transcript: Transcript, n, L, Rv, a, b, G_factors, H_factors, P, Q, G, H) -> bool:
    """InnerProductProof::verify (inner_product_proof.rs:321-382): True iff the proof opens P."""
    curve = ctx.curve
    rc = ctx.lib.bp_ipa_verify(ctx.h, transcript.h, n, codec.enc_points(L, curve), codec.enc_points(Rv, curve), codec.enc_scalars([a], curve),
                               codec.enc_scalars([b], curve), codec.enc_scalars(G_factors, curve), codec.enc_scalars(H_factors, curve),
                               codec.enc_point(P, curve), codec.enc_point(Q, curve), codec.enc_points(G, curve), codec.enc_points(H, curve))
    if rc == -7:
        return False
    ctx._check(rc)
    return True


# ---- gadgets of the reference's integration tests (tests/r1cs_secq256k1.rs), written against the
# ---- ConstraintSystem API above exactly as the Rust tests are written against the crate ----------
def example_gadget(cs, a1, a2, b1, b2, c1, c2):                    # tests/r1cs_secq256k1.rs:218-230
    _, _, c_var = cs.multiply(LC.of(a1) + a2, LC.of(b1) + b2)
    cs.constrain(LC.of(c1) + c2 - c_var)


def shuffle_gadget(cs, x, y):                                      # tests/r1cs_secq256k1.rs:16-56
    assert len(x) == len(y)
    k = len(x)
    if k == 1:
        cs.constrain(LC.of(y[0]) - x[0])
        return

    def cb(cs):
        z = cs.challenge_scalar(b"shuffle challenge")
        _, _, last_x = cs.multiply(LC.of(x[k - 1]) - z, LC.of(x[k - 2]) - z)
        first_x = last_x
        for i in reversed(range(k - 2)):
            _, _, first_x = cs.multiply(LC.of(first_x), LC.of(x[i]) - z)
        _, _, last_y = cs.multiply(LC.of(y[k - 1]) - z, LC.of(y[k - 2]) - z)
        first_y = last_y
        for i in reversed(range(k - 2)):
            _, _, first_y = cs.multiply(LC.of(first_y), LC.of(y[i]) - z)
        cs.constrain(LC.of(first_x) - first_y)

    cs.specify_randomized_constraints(cb)


def range_proof_gadget(cs, v_lc, v_assignment, n):                 # tests/r1cs_secq256k1.rs:361-393
    v = LC.of(v_lc)
    exp_2 = 1
    for i in range(n):
        assign = None
        if v_assignment is not None:
            bit = (v_assignment >> i) & 1
            assign = (1 - bit, bit)
        a, b, o = cs.allocate_multiplier(assign)
        cs.constrain(LC.of(o))
        cs.constrain(LC.of(a) + b - LC.of(1))
        v = v - LC.of(b) * exp_2
        exp_2 = exp_2 + exp_2
    cs.constrain(v)


def chain_circuit(cs, v_var, N, ks, x0, r):
    """one-phase public-multiplier chain (same as the oracle's)."""
    x = x0
    prev_o = None
    for i in range(N):
        assign = None if x is None else (x, ks[i])
        l, rr, o = cs.allocate_multiplier(assign)
        cs.constrain(LC.of(rr) - ks[i])
        if i == 0:
            cs.constrain(LC.of(l) - v_var)
        else:
            cs.constrain(LC.of(l) - prev_o)
        prev_o = o
        if x is not None:
            x = x * ks[i] % r

"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the ark-bulletproofs hot path.

This is a pure-Python big-integer restatement of the reference's algorithm
(FindoraNetwork/ark-bulletproofs v4.1.1) and of the un-vendored crates it
calls (ark-ff/ark-ec/ark-serialize 0.4, merlin 3, rand_chacha 0.3, sha3).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import it; the product (ark_bulletproofs_b200) never does.

PARITY UNPINNED: the reference's own tests draw everything from
`thread_rng()` (tests/r1cs_secq256k1.rs:140,243,251,400,422,491,508;
src/inner_product_proof.rs:414) and hold no golden bytes, and no Rust
toolchain exists in this image, so this restatement is anchored on
  * the public Merlin KAT, ChaCha20 keystream vectors, SHA3 (hashlib),
  * on-curve / group-order checks of every generator,
  * SURVEY.md Appendix B (an independent earlier write of the same spec),
  * the behavioural accept/reject matrix of the reference tests.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import hashlib
import struct
from dataclasses import dataclass, field as dc_field
from typing import Callable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# Keccak-f[1600] / STROBE-128 / Merlin   (merlin 3.0.0: strobe.rs, transcript.rs)
# --------------------------------------------------------------------------
_M64 = (1 << 64) - 1
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61],
        [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(lanes: List[int]) -> List[int]:
    """lanes[x + 5*y], 25 little-endian 64-bit words."""
    a = lanes
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= _RC[rnd]
    return a


def _permute_bytes(st: bytearray) -> None:
    lanes = list(struct.unpack("<25Q", bytes(st)))
    st[:] = struct.pack("<25Q", *keccak_f1600(lanes))


class Strobe128:
    """merlin/src/strobe.rs."""
    R = 166
    I, A, C, T, M, K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label: bytes = b"", _clone: "Strobe128" = None):
        if _clone is not None:
            self.state = bytearray(_clone.state)
            self.pos, self.pos_begin, self.cur_flags = _clone.pos, _clone.pos_begin, _clone.cur_flags
            return
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        _permute_bytes(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def clone(self):
        return Strobe128(_clone=self)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R + 1] ^= 0x80
        _permute_bytes(self.state)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.state[self.pos] ^= b
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _overwrite(self, data):
        for b in data:
            self.state[self.pos] = b
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n) -> bytes:
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & self.T == 0
        old = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old, flags]))
        if flags & (self.C | self.K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.M | self.A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.A, more)
        self._absorb(data)

    def prf(self, n, more) -> bytes:
        self._begin_op(self.I | self.A | self.C, more)
        return self._squeeze(n)

    def key(self, data, more):
        self._begin_op(self.A | self.C, more)
        self._overwrite(data)


class Transcript:
    """merlin::Transcript (merlin/src/transcript.rs)."""

    def __init__(self, label: bytes = b"", _strobe: Strobe128 = None):
        if _strobe is not None:
            self.strobe = _strobe
            return
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def clone(self):
        return Transcript(_strobe=self.strobe.clone())

    def append_message(self, label: bytes, msg: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", len(msg)), True)
        self.strobe.ad(msg, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, struct.pack("<Q", x))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", n), True)
        return self.strobe.prf(n, False)

    def build_rng(self):
        return TranscriptRngBuilder(self.strobe.clone())


class TranscriptRngBuilder:
    def __init__(self, strobe):
        self.strobe = strobe

    def rekey_with_witness_bytes(self, label: bytes, witness: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", len(witness)), True)
        self.strobe.key(witness, False)
        return self

    def finalize(self, rng) -> "TranscriptRng":
        rb = rng.fill_bytes(32)
        self.strobe.meta_ad(b"rng", False)
        self.strobe.key(rb, False)
        return TranscriptRng(self.strobe)


class TranscriptRng:
    """merlin TranscriptRng: one Keccak-f per next_u64."""

    def __init__(self, strobe):
        self.strobe = strobe

    def fill_bytes(self, n) -> bytes:
        self.strobe.meta_ad(struct.pack("<I", n), False)
        return self.strobe.prf(n, False)

    def next_u32(self) -> int:
        return struct.unpack("<I", self.fill_bytes(4))[0]

    def next_u64(self) -> int:
        return struct.unpack("<Q", self.fill_bytes(8))[0]


# --------------------------------------------------------------------------
# ChaCha20Rng (rand_chacha 0.3): 64-bit counter in words 12-13, stream id 0
# --------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def _chacha_block(key_words, counter):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + \
         [counter & _M32, (counter >> 32) & _M32, 0, 0]
    w = st[:]

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & _M32; w[d] ^= w[a]; w[d] = ((w[d] << 16) | (w[d] >> 16)) & _M32
        w[c] = (w[c] + w[d]) & _M32; w[b] ^= w[c]; w[b] = ((w[b] << 12) | (w[b] >> 20)) & _M32
        w[a] = (w[a] + w[b]) & _M32; w[d] ^= w[a]; w[d] = ((w[d] << 8) | (w[d] >> 24)) & _M32
        w[c] = (w[c] + w[d]) & _M32; w[b] ^= w[c]; w[b] = ((w[b] << 7) | (w[b] >> 25)) & _M32

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(w[i] + st[i]) & _M32 for i in range(16)]


class ChaCha20Rng:
    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf: List[int] = []
        self.words_used = 0

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = _chacha_block(self.key, self.counter)
            self.counter += 1
        self.words_used += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def fill_bytes(self, n) -> bytes:
        out = b"".join(struct.pack("<I", self.next_u32()) for _ in range((n + 3) // 4))
        return out[:n]


# --------------------------------------------------------------------------
# Fields and curves (src/curve/zorro/*.rs)
# --------------------------------------------------------------------------
R256 = 1 << 256


@dataclass
class Curve:
    name: str
    q: int            # base field modulus
    r: int            # scalar field modulus (group order / cofactor)
    kind: str         # 'sw' or 'te'
    a: int
    b: int            # SW b, or TE d
    gx: int
    gy: int
    cofactor: int = 1
    q_bits: int = 256
    r_bits: int = 256
    rinv_q: int = dc_field(init=False)
    rinv_r: int = dc_field(init=False)

    def __post_init__(self):
        self.rinv_q = pow(R256, -1, self.q)
        self.rinv_r = pow(R256, -1, self.r)
        self.q_bits = self.q.bit_length()
        self.r_bits = self.r.bit_length()

    @property
    def G(self):
        return (self.gx, self.gy)


SECP_P = 2**256 - 2**32 - 977
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141

# ark-secq256k1 0.4.0: y^2 = x^3 + 7 over F_n, order p  
SECQ256K1 = Curve(
    "secq256k1", SECP_N, SECP_P, "sw", 0, 7,
    53718550993811904772965658690407829053653678808745171666022356150019200052646,
    28941648020349172432234515805717979317553499307621291159490218670604692907903)

# src/curve/zorro/{fq.rs:4, fr.rs:1, g1.rs:25-46}
ZORRO = Curve(
    "zorro",
    57896044618658097711785492504343953927116110621106131396339151912985063395361,
    2**255 - 19, "sw", 6,
    7277470329389939148381533754641607518092114590371880995609984561067837624798,
    2, 19711758720854384559191066596451394956860102304684364148268676039962145446511)

# ark-curve25519 0.4.0 (dev-dependency, Cargo.toml:67-70): -x^2+y^2 = 1+d x^2 y^2
_ED_Q = 2**255 - 19
_ED_D = (-121665 * pow(121666, -1, _ED_Q)) % _ED_Q
CURVE25519 = Curve(
    "curve25519", _ED_Q, 2**252 + 27742317777372353535851937790883648493, "te",
    _ED_Q - 1, _ED_D,
    15112221349535400772501151409588531511454012693041857206046113283949847762202,
    46316835694926478169428394003475163141307993866256225615783033603165251855960,
    cofactor=8)

CURVES = {c.name: c for c in (SECQ256K1, ZORRO, CURVE25519)}


def sqrt_mod(a: int, p: int) -> Optional[int]:
    """Any square root or None (Tonelli-Shanks). Root choice is irrelevant:
    callers order (y, -y) canonically like ark-ec get_ys_from_x_unchecked."""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    s, t = p - 1, 0
    while s % 2 == 0:
        s //= 2
        t += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, tt, rr = t, pow(z, s, p), pow(a, s, p), pow(a, (s + 1) // 2, p)
    while tt != 1:
        i, t2 = 0, tt
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        bb = pow(c, 1 << (m - i - 1), p)
        m, c = i, bb * bb % p
        tt, rr = tt * c % p, rr * bb % p
    return rr


# ---- group law: affine tuples (x, y) or None = identity; Jacobian/extended inside
def on_curve(cv: Curve, P) -> bool:
    if P is None:
        return True
    x, y = P
    if cv.kind == "sw":
        return (y * y - (x * x * x + cv.a * x + cv.b)) % cv.q == 0
    return (cv.a * x * x + y * y - 1 - cv.b * x * x * y * y) % cv.q == 0


def _te_id(P):
    return P is None or (P[0] == 0 and P[1] == 1)


def pt_neg(cv, P):
    if P is None:
        return None
    if cv.kind == "sw":
        return (P[0], (-P[1]) % cv.q)
    return ((-P[0]) % cv.q, P[1])


def pt_add(cv: Curve, P, Q):
    q = cv.q
    if cv.kind == "te":
        if P is None:
            P = (0, 1)
        if Q is None:
            Q = (0, 1)
        x1, y1 = P
        x2, y2 = Q
        t = cv.b * x1 * x2 * y1 * y2 % q
        x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, q) % q
        y3 = (y1 * y2 - cv.a * x1 * x2) * pow(1 - t, -1, q) % q
        return None if (x3 == 0 and y3 == 1) else (x3, y3)
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % q == 0:
            return None
        lam = (3 * x1 * x1 + cv.a) * pow(2 * y1, -1, q) % q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, q) % q
    x3 = (lam * lam - x1 - x2) % q
    return (x3, (lam * (x1 - x3) - y1) % q)


# Jacobian (SW) for speed in scalar multiplications / MSM
def _jac_dbl(cv, P):
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    q = cv.q
    S = 4 * X * Y * Y % q
    M = (3 * X * X + cv.a * pow(Z, 4, q)) % q
    X3 = (M * M - 2 * S) % q
    Y3 = (M * (S - X3) - 8 * pow(Y, 4, q)) % q
    return (X3, Y3, 2 * Y * Z % q)


def _jac_add(cv, P, Qp):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Qp
    if Z1 == 0:
        return Qp
    if Z2 == 0:
        return P
    q = cv.q
    Z1Z1, Z2Z2 = Z1 * Z1 % q, Z2 * Z2 % q
    U1, U2 = X1 * Z2Z2 % q, X2 * Z1Z1 % q
    S1, S2 = Y1 * Z2 * Z2Z2 % q, Y2 * Z1 * Z1Z1 % q
    if U1 == U2:
        if S1 != S2:
            return (1, 1, 0)
        return _jac_dbl(cv, P)
    H = (U2 - U1) % q
    Rr = (S2 - S1) % q
    HH = H * H % q
    HHH = H * HH % q
    V = U1 * HH % q
    X3 = (Rr * Rr - HHH - 2 * V) % q
    Y3 = (Rr * (V - X3) - S1 * HHH) % q
    return (X3, Y3, H * Z1 * Z2 % q)


def _to_jac(P):
    return (1, 1, 0) if P is None else (P[0], P[1], 1)


def _from_jac(cv, P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, cv.q)
    zi2 = zi * zi % cv.q
    return (X * zi2 % cv.q, Y * zi2 * zi % cv.q)


def pt_mul(cv: Curve, k: int, P):
    """k*P, k taken as a non-negative integer (mul_bigint semantics)."""
    if P is None or k == 0:
        return None
    if cv.kind == "te":
        acc, base = None, P
        while k:
            if k & 1:
                acc = pt_add(cv, acc, base)
            base = pt_add(cv, base, base)
            k >>= 1
        return acc
    acc = (1, 1, 0)
    jp = _to_jac(P)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(cv, acc)
        if bit == "1":
            acc = _jac_add(cv, acc, jp)
    return _from_jac(cv, acc)


def msm(cv: Curve, points: Sequence, scalars: Sequence[int]):
    """Sum s_i*P_i (the *value* of ark-ec VariableBaseMSM::msm; algorithm free).
    Reference call sites: SURVEY.md section 8(a) row a1."""
    assert len(points) == len(scalars)
    n = len(points)
    if cv.kind == "te" or n < 8:
        acc = None
        for P, s in zip(points, scalars):
            acc = pt_add(cv, acc, pt_mul(cv, s % cv.r, P))
        return acc
    c = max(2, min(12, n.bit_length() - 2))
    nb = cv.r.bit_length()
    wins = (nb + c - 1) // c
    jp = [_to_jac(P) for P in points]
    ss = [s % cv.r for s in scalars]
    total = (1, 1, 0)
    for w in reversed(range(wins)):
        for _ in range(c):
            total = _jac_dbl(cv, total)
        buckets = [None] * (1 << c)
        for P, s in zip(jp, ss):
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = P if buckets[d] is None else _jac_add(cv, buckets[d], P)
        run, acc = (1, 1, 0), (1, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            if buckets[d] is not None:
                run = _jac_add(cv, run, buckets[d])
            acc = _jac_add(cv, acc, run)
        total = _jac_add(cv, total, acc)
    return _from_jac(cv, total)


# --------------------------------------------------------------------------
# arkworks sampling and serialization
# --------------------------------------------------------------------------
def fp_rand(modulus: int, rng) -> int:
    """ark-ff Fp::rand: 4 x next_u64 -> limbs (LS first), shave top bits,
    accept iff raw < modulus; raw IS the Montgomery representation, so the
    value is raw * R^-1 mod m."""
    bits = modulus.bit_length()
    mask = (1 << 64) - 1 >> (256 - bits)
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= mask
        raw = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if raw < modulus:
            return raw * pow(R256, -1, modulus) % modulus


def scalar_rand(cv: Curve, rng) -> int:
    return fp_rand(cv.r, rng)


def affine_rand(cv: Curve, rng):
    """ark-ec Affine::rand (SW: x then `greatest` bool; TE: y then bool), then
    mul_by_cofactor. Used by generators.rs:63,99,115."""
    q = cv.q
    while True:
        c0 = fp_rand(q, rng)
        greatest = (rng.next_u32() >> 31) & 1
        if cv.kind == "sw":
            y = sqrt_mod((c0 * c0 * c0 + cv.a * c0 + cv.b) % q, q)
            if y is None:
                continue
            lo, hi = sorted((y, (-y) % q))
            return (c0, hi if greatest else lo)
        # TE: x^2 = (1 - y^2) / (a - d y^2)
        y = c0
        num = (1 - y * y) % q
        den = (cv.a - cv.b * y * y) % q
        if den == 0:
            continue
        x = sqrt_mod(num * pow(den, -1, q) % q, q)
        if x is None:
            continue
        lo, hi = sorted((x, (-x) % q))
        P = (hi if greatest else lo, y)
        return pt_mul(cv, cv.cofactor, P)


def ser_scalar(cv: Curve, s: int) -> bytes:
    return (s % cv.r).to_bytes(32, "little")


def de_scalar(cv: Curve, b: bytes) -> Optional[int]:
    v = int.from_bytes(b, "little")
    return v if v < cv.r else None


def point_size(cv: Curve, compressed: bool) -> int:
    if cv.kind == "sw":
        return 33 if compressed else 65
    return 32 if compressed else 64


def ser_point(cv: Curve, P, compressed: bool) -> bytes:
    q = cv.q
    if cv.kind == "sw":
        if P is None:
            return bytes(32 if compressed else 64) + b"\x40"
        x, y = P
        flag = 0x80 if y > (-y) % q else 0
        if compressed:
            return x.to_bytes(32, "little") + bytes([flag])
        return x.to_bytes(32, "little") + y.to_bytes(32, "little") + bytes([flag])
    if P is None:
        P = (0, 1)
    x, y = P
    if compressed:
        out = bytearray(y.to_bytes(32, "little"))
        if x > (-x) % q:
            out[31] |= 0x80
        return bytes(out)
    return x.to_bytes(32, "little") + y.to_bytes(32, "little")


def de_point_compressed(cv: Curve, b: bytes):
    """deserialize_compressed with validation; raises ValueError on failure
    (-> R1CSError::FormatError, src/r1cs/proof.rs:83-91)."""
    q = cv.q
    if cv.kind == "sw":
        if len(b) != 33:
            raise ValueError("len")
        # ark-ff 0.4 deserialize_with_flags (restated from the published crate): the 33rd byte carries SWFlags in its top
        # two bits (from_u8_remove_flags clears only those; both set -> UnexpectedFlags) and the integer is rebuilt from
        # the first 32 bytes alone, so the six low bits of the flag byte are never looked at; x >= q -> InvalidData.
        # ark-ec Affine::deserialize_with_mode returns the identity on the infinity flag without examining x.
        flags = b[32]
        if (flags & 0xC0) == 0xC0:
            raise ValueError("both flags")
        x = int.from_bytes(b[:32], "little")
        if x >= q:
            raise ValueError("x >= q")
        if flags & 0x40:
            return None
        y = sqrt_mod((x * x * x + cv.a * x + cv.b) % q, q)
        if y is None:
            raise ValueError("not on curve")
        lo, hi = sorted((y, (-y) % q))
        return (x, hi if flags & 0x80 else lo)
    if len(b) != 32:
        raise ValueError("len")
    raw = int.from_bytes(b, "little")
    sign = raw >> 255
    y = raw & ((1 << 255) - 1)
    if y >= q:
        raise ValueError("y >= q")
    den = (cv.a - cv.b * y * y) % q
    x = sqrt_mod((1 - y * y) * pow(den, -1, q) % q, q)
    if x is None:
        raise ValueError("not on curve")
    lo, hi = sorted((x, (-x) % q))
    P = (hi if sign else lo, y)
    if pt_mul(cv, cv.r, P) is not None and not _te_id(pt_mul(cv, cv.r, P)):
        raise ValueError("subgroup")
    return None if _te_id(P) else P


# --------------------------------------------------------------------------
# src/transcript.rs
# --------------------------------------------------------------------------
class ProofError(Exception):
    pass


class R1CSError(Exception):
    pass


def append_scalar(cv, t: Transcript, label: bytes, s: int):      # transcript.rs:69-73
    t.append_message(label, ser_scalar(cv, s))


def append_point(cv, t: Transcript, label: bytes, P):            # transcript.rs:75-79
    t.append_message(label, ser_point(cv, P, False))


def validate_and_append_point(cv, t, label, P):                   # transcript.rs:81-93
    if P is None or (cv.kind == "te" and _te_id(P)):
        raise ProofError("VerificationError")
    t.append_message(label, ser_point(cv, P, False))


def challenge_scalar(cv, t: Transcript, label: bytes) -> int:     # transcript.rs:95-101
    return scalar_rand(cv, ChaCha20Rng(t.challenge_bytes(label, 32)))


# --------------------------------------------------------------------------
# src/generators.rs
# --------------------------------------------------------------------------
class PedersenGens:
    def __init__(self, cv: Curve):                                 # generators.rs:47-66
        self.cv = cv
        self.B = cv.G
        seed = hashlib.sha3_512(ser_point(cv, cv.G, False)).digest()[:32]
        self.B_blinding = affine_rand(cv, ChaCha20Rng(seed))

    def commit(self, value: int, blinding: int):                   # generators.rs:39-44
        cv = self.cv
        return pt_add(cv, pt_mul(cv, value % cv.r, self.B), pt_mul(cv, blinding % cv.r, self.B_blinding))


def generators_chain(cv: Curve, label: bytes):                     # generators.rs:78-121
    seed = hashlib.sha3_512(b"GeneratorsChain" + label).digest()[:32]
    rng = ChaCha20Rng(seed)
    while True:
        yield affine_rand(cv, rng)


class BulletproofGens:
    def __init__(self, cv: Curve, gens_capacity: int, party_capacity: int = 1):   # generators.rs:174-183
        self.cv = cv
        self.gens_capacity = 0
        self.party_capacity = party_capacity
        self.G_vec = [[] for _ in range(party_capacity)]
        self.H_vec = [[] for _ in range(party_capacity)]
        self._chains = None
        self.increase_capacity(gens_capacity)

    def increase_capacity(self, new_capacity: int):                # generators.rs:196-221
        if self.gens_capacity >= new_capacity:
            return
        if self._chains is None:
            self._chains = [(generators_chain(self.cv, b"G" + struct.pack("<I", i)),
                             generators_chain(self.cv, b"H" + struct.pack("<I", i)))
                            for i in range(self.party_capacity)]
        for i in range(self.party_capacity):
            gc, hc = self._chains[i]     # a live chain == new chain fast-forwarded
            for _ in range(new_capacity - self.gens_capacity):
                self.G_vec[i].append(next(gc))
            for _ in range(new_capacity - self.gens_capacity):
                self.H_vec[i].append(next(hc))
        self.gens_capacity = new_capacity

    def G(self, n):                                                # generators.rs:296-298 (share 0)
        return self.G_vec[0][:n]

    def H(self, n):                                                # generators.rs:301-303
        return self.H_vec[0][:n]


# --------------------------------------------------------------------------
# src/inner_product_proof.rs
# --------------------------------------------------------------------------
def inner_product(cv, a, b) -> int:                                # inner_product_proof.rs:390-399
    assert len(a) == len(b)
    return sum(x * y for x, y in zip(a, b)) % cv.r


@dataclass
class InnerProductProof:
    L_vec: list
    R_vec: list
    a: int
    b: int

    def to_bytes(self, cv) -> bytes:
        out = struct.pack("<Q", len(self.L_vec)) + b"".join(ser_point(cv, P, True) for P in self.L_vec)
        out += struct.pack("<Q", len(self.R_vec)) + b"".join(ser_point(cv, P, True) for P in self.R_vec)
        return out + ser_scalar(cv, self.a) + ser_scalar(cv, self.b)


def fold_generators(cv: Curve, L, R, sL, sR):
    """inner_product_proof.rs:139-156 (first round, per-element factors) and :216-225: one 2-point msm + into_affine
    per output, out[i] = sL[i]*L[i] + sR[i]*R[i]. (oracle/fast.py swaps in the C restatement for the large goldens.)"""
    return [pt_add(cv, pt_mul(cv, a_, P), pt_mul(cv, b_, Q_)) for P, Q_, a_, b_ in zip(L, R, sL, sR)]


def ipa_create(cv: Curve, t: Transcript, Q, G_factors, H_factors, G, H, a, b) -> InnerProductProof:
    """inner_product_proof.rs:37-239 — the reference's loop structure verbatim
    (per-element 2-point msm for the generator fold)."""
    r = cv.r
    G, H, a, b = list(G), list(H), list(a), list(b)
    n = len(G)
    assert len(H) == n and len(a) == n and len(b) == n and len(G_factors) == n and len(H_factors) == n
    assert n & (n - 1) == 0 and n > 0
    t.append_message(b"dom-sep", b"ipp v1")                        # transcript.rs:52-55
    t.append_u64(b"n", n)
    L_vec, R_vec = [], []
    first = True
    while n != 1:
        n //= 2
        aL, aR, bL, bR = a[:n], a[n:2 * n], b[:n], b[n:2 * n]
        GL, GR, HL, HR = G[:n], G[n:2 * n], H[:n], H[n:2 * n]
        cL, cR = inner_product(cv, aL, bR), inner_product(cv, aR, bL)
        if first:
            gf, hf = G_factors, H_factors
            Ls = [aL[i] * gf[n + i] % r for i in range(n)] + [bR[i] * hf[i] % r for i in range(n)] + [cL]
            Rs = [aR[i] * gf[i] % r for i in range(n)] + [bL[i] * hf[n + i] % r for i in range(n)] + [cR]
        else:
            Ls = aL + bR + [cL]
            Rs = aR + bL + [cR]
        Lp = msm(cv, GR + HL + [Q], Ls)
        Rp = msm(cv, GL + HR + [Q], Rs)
        L_vec.append(Lp)
        R_vec.append(Rp)
        append_point(cv, t, b"L", Lp)
        append_point(cv, t, b"R", Rp)
        u = challenge_scalar(cv, t, b"u")
        ui = pow(u, -1, r)
        for i in range(n):
            a[i] = (aL[i] * u + ui * aR[i]) % r
            b[i] = (bL[i] * ui + u * bR[i]) % r
        if first:
            g0 = [ui * gf[i] % r for i in range(n)]
            g1 = [u * gf[n + i] % r for i in range(n)]
            h0 = [u * hf[i] % r for i in range(n)]
            h1 = [ui * hf[n + i] % r for i in range(n)]
        else:
            g0, g1, h0, h1 = [ui] * n, [u] * n, [u] * n, [ui] * n
        G[:n] = fold_generators(cv, GL, GR, g0, g1)
        H[:n] = fold_generators(cv, HL, HR, h0, h1)
        a, b, G, H = a[:n], b[:n], G[:n], H[:n]
        first = False
    return InnerProductProof(L_vec, R_vec, a[0], b[0])


def ipa_verification_scalars(cv, proof: InnerProductProof, n: int, t: Transcript):
    """inner_product_proof.rs:244-314."""
    r = cv.r
    lg_n = len(proof.L_vec)
    if lg_n >= 32 or n != (1 << lg_n) or len(proof.R_vec) != lg_n:
        raise ProofError("VerificationError")
    t.append_message(b"dom-sep", b"ipp v1")
    t.append_u64(b"n", n)
    ch = []
    for L, Rp in zip(proof.L_vec, proof.R_vec):
        validate_and_append_point(cv, t, b"L", L)
        validate_and_append_point(cv, t, b"R", Rp)
        ch.append(challenge_scalar(cv, t, b"u"))
    ch_inv = [pow(c, -1, r) if c else 0 for c in ch]
    allinv = 1
    for f in ch_inv:
        if f:
            allinv = allinv * f % r
    ch_sq = [c * c % r for c in ch]
    ch_inv_sq = [c * c % r for c in ch_inv]
    s = [allinv]
    for i in range(1, n):
        lg_i = i.bit_length() - 1
        s.append(s[i - (1 << lg_i)] * ch_sq[(lg_n - 1) - lg_i] % r)
    return ch_sq, ch_inv_sq, s


def ipa_verify(cv, proof, n, t, G_factors, H_factors, P, Q, G, H):
    """inner_product_proof.rs:321-382 (test-only in the reference)."""
    r = cv.r
    u_sq, u_inv_sq, s = ipa_verification_scalars(cv, proof, n, t)
    gs = [proof.a * s[i] % r * G_factors[i] % r for i in range(n)]
    hs = [proof.b * s[n - 1 - i] % r * H_factors[i] % r for i in range(n)]
    bases = [Q] + list(G) + list(H) + proof.L_vec + proof.R_vec
    scal = [proof.a * proof.b % r] + gs + hs + [(-x) % r for x in u_sq] + [(-x) % r for x in u_inv_sq]
    if msm(cv, bases, scal) != P:
        raise ProofError("VerificationError")


# --------------------------------------------------------------------------
# src/r1cs/linear_combination.rs, constraint_system.rs
# --------------------------------------------------------------------------
COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = "V", "L", "R", "O", "1"


class LC:
    """LinearCombination: list of (Variable, coeff); Variable = (kind, index)."""

    def __init__(self, terms=None):
        self.terms = list(terms or [])

    @staticmethod
    def of(x) -> "LC":
        if isinstance(x, LC):
            return LC(x.terms)
        if isinstance(x, tuple):
            return LC([(x, 1)])
        return LC([((ONE, 0), int(x))])

    def __add__(self, o):
        return LC(self.terms + LC.of(o).terms)

    def __sub__(self, o):
        return LC(self.terms + [(v, -c) for v, c in LC.of(o).terms])

    def __neg__(self):
        return LC([(v, -c) for v, c in self.terms])

    def scale(self, k):
        return LC([(v, c * k) for v, c in self.terms])


def var_one():
    return (ONE, 0)


# --------------------------------------------------------------------------
# src/r1cs/proof.rs
# --------------------------------------------------------------------------
@dataclass
class R1CSProof:
    A_I1: object
    A_O1: object
    S1: object
    A_I2: object
    A_O2: object
    S2: object
    T_1: object
    T_3: object
    T_4: object
    T_5: object
    T_6: object
    t_x: int
    t_x_blinding: int
    e_blinding: int
    ipp_proof: InnerProductProof

    def to_bytes(self, cv) -> bytes:                               # proof.rs:74-78
        pts = [self.A_I1, self.A_O1, self.S1, self.A_I2, self.A_O2, self.S2,
               self.T_1, self.T_3, self.T_4, self.T_5, self.T_6]
        out = b"".join(ser_point(cv, P, True) for P in pts)
        out += ser_scalar(cv, self.t_x) + ser_scalar(cv, self.t_x_blinding) + ser_scalar(cv, self.e_blinding)
        return out + self.ipp_proof.to_bytes(cv)

    @staticmethod
    def from_bytes(cv, data: bytes) -> "R1CSProof":                # proof.rs:83-91
        try:
            ps = point_size(cv, True)
            off = 0
            pts = []
            for _ in range(11):
                pts.append(de_point_compressed(cv, data[off:off + ps]))
                off += ps
            sc = []
            for _ in range(3):
                v = de_scalar(cv, data[off:off + 32])
                if v is None or len(data[off:off + 32]) != 32:
                    raise ValueError("scalar")
                sc.append(v)
                off += 32
            vecs = []
            for _ in range(2):
                if off + 8 > len(data):
                    raise ValueError("short")
                (ln,) = struct.unpack("<Q", data[off:off + 8])
                off += 8
                if ln > (len(data) - off) // ps:
                    raise ValueError("short")
                v = []
                for _ in range(ln):
                    v.append(de_point_compressed(cv, data[off:off + ps]))
                    off += ps
                vecs.append(v)
            ab = []
            for _ in range(2):
                if len(data[off:off + 32]) != 32:
                    raise ValueError("short")
                v = de_scalar(cv, data[off:off + 32])
                if v is None:
                    raise ValueError("scalar")
                ab.append(v)
                off += 32
            return R1CSProof(*pts, *sc, InnerProductProof(vecs[0], vecs[1], ab[0], ab[1]))
        except ValueError as e:
            raise R1CSError("FormatError") from e


# --------------------------------------------------------------------------
# src/r1cs/prover.rs
# --------------------------------------------------------------------------
class Prover:
    def __init__(self, cv: Curve, pc_gens: PedersenGens, transcript: Transcript):   # prover.rs:291-308
        self.cv, self.pc_gens, self.transcript = cv, pc_gens, transcript
        transcript.append_message(b"dom-sep", b"r1cs v1")
        self.v, self.v_blinding = [], []
        self.a_L, self.a_R, self.a_O = [], [], []
        self.constraints: List[LC] = []
        self.deferred: List[Callable] = []
        self.pending_multiplier = None
        self.randomizing = False

    # ConstraintSystem (prover.rs:96-194)
    def _eval(self, lc: LC) -> int:                                # prover.rs:399-414
        tot = 0
        for (k, i), c in lc.terms:
            val = {COMMITTED: lambda: self.v[i], MUL_LEFT: lambda: self.a_L[i], MUL_RIGHT: lambda: self.a_R[i],
                   MUL_OUT: lambda: self.a_O[i], ONE: lambda: 1}[k]()
            tot += c * val
        return tot % self.cv.r

    def multiply(self, left, right):                               # prover.rs:103-133
        left, right = LC.of(left), LC.of(right)
        l, rr = self._eval(left), self._eval(right)
        o = l * rr % self.cv.r
        i = len(self.a_L)
        lv, rv, ov = (MUL_LEFT, i), (MUL_RIGHT, i), (MUL_OUT, i)
        self.a_L.append(l); self.a_R.append(rr); self.a_O.append(o)
        left.terms.append((lv, -1))
        right.terms.append((rv, -1))
        self.constrain(left)
        self.constrain(right)
        return lv, rv, ov

    def allocate(self, assignment):                                # prover.rs:135-157
        if assignment is None:
            raise R1CSError("MissingAssignment")
        if self.pending_multiplier is None:
            i = len(self.a_L)
            self.pending_multiplier = i
            self.a_L.append(assignment % self.cv.r); self.a_R.append(0); self.a_O.append(0)
            return (MUL_LEFT, i)
        i = self.pending_multiplier
        self.pending_multiplier = None
        self.a_R[i] = assignment % self.cv.r
        self.a_O[i] = self.a_L[i] * self.a_R[i] % self.cv.r
        return (MUL_RIGHT, i)

    def allocate_multiplier(self, assignments):                    # prover.rs:159-183
        if assignments is None:
            raise R1CSError("MissingAssignment")
        l, rr = assignments
        i = len(self.a_L)
        self.a_L.append(l % self.cv.r); self.a_R.append(rr % self.cv.r); self.a_O.append(l * rr % self.cv.r)
        return (MUL_LEFT, i), (MUL_RIGHT, i), (MUL_OUT, i)

    def multipliers_len(self):
        return len(self.a_L)

    def constrain(self, lc):                                       # prover.rs:189-193
        self.constraints.append(LC.of(lc))

    def specify_randomized_constraints(self, cb):                  # prover.rs:201-207
        self.deferred.append(cb)

    def challenge_scalar(self, label: bytes) -> int:               # prover.rs:262-267
        assert self.randomizing
        return challenge_scalar(self.cv, self.transcript, label)

    def commit(self, v: int, v_blinding: int):                     # prover.rs:327-341
        i = len(self.v)
        self.v.append(v % self.cv.r)
        self.v_blinding.append(v_blinding % self.cv.r)
        V = self.pc_gens.commit(v, v_blinding)
        append_point(self.cv, self.transcript, b"V", V)
        return V, (COMMITTED, i)

    def _flattened_constraints(self, z):                           # prover.rs:354-397
        r = self.cv.r
        n, m = len(self.a_L), len(self.v)
        wL, wR, wO, wV = [0] * n, [0] * n, [0] * n, [0] * m
        exp_z = z
        for lc in self.constraints:
            for (k, i), c in lc.terms:
                if k == MUL_LEFT:
                    wL[i] = (wL[i] + exp_z * c) % r
                elif k == MUL_RIGHT:
                    wR[i] = (wR[i] + exp_z * c) % r
                elif k == MUL_OUT:
                    wO[i] = (wO[i] + exp_z * c) % r
                elif k == COMMITTED:
                    wV[i] = (wV[i] - exp_z * c) % r
            exp_z = exp_z * z % r
        return wL, wR, wO, wV

    def _create_randomized_constraints(self):                      # prover.rs:418-441
        self.pending_multiplier = None
        if not self.deferred:
            self.transcript.append_message(b"dom-sep", b"r1cs-1phase")
        else:
            self.transcript.append_message(b"dom-sep", b"r1cs-2phase")
            cbs, self.deferred = self.deferred, []
            self.randomizing = True
            for cb in cbs:
                cb(self)
            self.randomizing = False

    def prove(self, prng, bp_gens: BulletproofGens, trace: dict = None) -> R1CSProof:   # prover.rs:454-831
        cv, r, t = self.cv, self.cv.r, self.transcript
        t.append_u64(b"m", len(self.v))
        builder = t.build_rng()
        for vb in self.v_blinding:
            builder = builder.rekey_with_witness_bytes(b"v_blinding", ser_scalar(cv, vb))
        rng = builder.finalize(prng)
        n1 = len(self.a_L)
        if bp_gens.gens_capacity < n1:
            raise R1CSError("InvalidGeneratorsLength")
        Bb = self.pc_gens.B_blinding
        i_bl1, o_bl1, s_bl1 = scalar_rand(cv, rng), scalar_rand(cv, rng), scalar_rand(cv, rng)
        s_L1 = [scalar_rand(cv, rng) for _ in range(n1)]
        s_R1 = [scalar_rand(cv, rng) for _ in range(n1)]
        G1, H1 = bp_gens.G(n1), bp_gens.H(n1)
        A_I1 = msm(cv, [Bb] + G1 + H1, [i_bl1] + self.a_L + self.a_R)
        A_O1 = msm(cv, [Bb] + G1, [o_bl1] + self.a_O)
        S1 = msm(cv, [Bb] + G1 + H1, [s_bl1] + s_L1 + s_R1)
        append_point(cv, t, b"A_I1", A_I1)
        append_point(cv, t, b"A_O1", A_O1)
        append_point(cv, t, b"S1", S1)
        self._create_randomized_constraints()
        n = len(self.a_L)
        n2 = n - n1
        padded_n = 1 if n == 0 else 1 << (n - 1).bit_length()
        pad = padded_n - n
        if bp_gens.gens_capacity < padded_n:
            raise R1CSError("InvalidGeneratorsLength")
        if n2 > 0:
            i_bl2, o_bl2, s_bl2 = scalar_rand(cv, rng), scalar_rand(cv, rng), scalar_rand(cv, rng)
        else:
            i_bl2 = o_bl2 = s_bl2 = 0
        s_L2 = [scalar_rand(cv, rng) for _ in range(n2)]
        s_R2 = [scalar_rand(cv, rng) for _ in range(n2)]
        if n2 > 0:
            G2, H2 = bp_gens.G(n)[n1:], bp_gens.H(n)[n1:]
            A_I2 = msm(cv, [Bb] + G2 + H2, [i_bl2] + self.a_L[n1:] + self.a_R[n1:])
            A_O2 = msm(cv, [Bb] + G2, [o_bl2] + self.a_O[n1:])
            S2 = msm(cv, [Bb] + G2 + H2, [s_bl2] + s_L2 + s_R2)
        else:
            A_I2 = A_O2 = S2 = None
        append_point(cv, t, b"A_I2", A_I2)
        append_point(cv, t, b"A_O2", A_O2)
        append_point(cv, t, b"S2", S2)
        y = challenge_scalar(cv, t, b"y")
        z = challenge_scalar(cv, t, b"z")
        wL, wR, wO, wV = self._flattened_constraints(z)
        y_inv = pow(y, -1, r)
        exp_y_inv = [1] * padded_n
        for i in range(1, padded_n):
            exp_y_inv[i] = exp_y_inv[i - 1] * y_inv % r
        sL, sR = s_L1 + s_L2, s_R1 + s_R2
        l1, l2, l3 = [0] * n, [0] * n, [0] * n
        r0, r1, r3 = [0] * n, [0] * n, [0] * n
        exp_y = 1
        for i in range(n):                                         # prover.rs:684-701
            l1[i] = (self.a_L[i] + exp_y_inv[i] * wR[i]) % r
            l2[i] = self.a_O[i]
            l3[i] = sL[i]
            r0[i] = (wO[i] - exp_y) % r
            r1[i] = (exp_y * self.a_R[i] + wL[i]) % r
            r3[i] = exp_y * sR[i] % r
            exp_y = exp_y * y % r
        ip = lambda a, b: inner_product(cv, a, b)
        zero = [0] * n                                             # util.rs:75-93
        t1 = ip(l1, r0)
        t2 = (ip(l1, r1) + ip(l2, r0)) % r
        t3 = (ip(l2, r1) + ip(l3, r0)) % r
        t4 = (ip(l1, r3) + ip(l3, r1)) % r
        t5 = ip(l2, r3)
        t6 = ip(l3, r3)
        tb1, tb3, tb4, tb5, tb6 = (scalar_rand(cv, rng) for _ in range(5))
        T_1 = self.pc_gens.commit(t1, tb1)
        T_3 = self.pc_gens.commit(t3, tb3)
        T_4 = self.pc_gens.commit(t4, tb4)
        T_5 = self.pc_gens.commit(t5, tb5)
        T_6 = self.pc_gens.commit(t6, tb6)
        for lab, P in ((b"T_1", T_1), (b"T_3", T_3), (b"T_4", T_4), (b"T_5", T_5), (b"T_6", T_6)):
            append_point(cv, t, lab, P)
        u = challenge_scalar(cv, t, b"u")
        x = challenge_scalar(cv, t, b"x")
        tb2 = sum(c * vb for c, vb in zip(wV, self.v_blinding)) % r
        poly6 = lambda c: x * (c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * (c[4] + x * c[5]))))) % r
        t_x = poly6([t1, t2, t3, t4, t5, t6])
        t_x_blinding = poly6([tb1, tb2, tb3, tb4, tb5, tb6])
        l_vec = [(x * (l1[i] + x * (l2[i] + x * l3[i]))) % r for i in range(n)] + [0] * pad
        r_vec = [(r0[i] + x * (r1[i] + x * (x * r3[i]))) % r for i in range(n)] + [0] * pad
        for i in range(n, padded_n):                               # prover.rs:753-756
            r_vec[i] = (-exp_y) % r
            exp_y = exp_y * y % r
        i_bl = (i_bl1 + u * i_bl2) % r
        o_bl = (o_bl1 + u * o_bl2) % r
        s_bl = (s_bl1 + u * s_bl2) % r
        e_blinding = x * (i_bl + x * (o_bl + x * s_bl)) % r
        append_scalar(cv, t, b"t_x", t_x)
        append_scalar(cv, t, b"t_x_blinding", t_x_blinding)
        append_scalar(cv, t, b"e_blinding", e_blinding)
        w = challenge_scalar(cv, t, b"w")
        Q = pt_mul(cv, w, self.pc_gens.B)
        G_factors = [1] * n1 + [u] * (n2 + pad)
        H_factors = [exp_y_inv[i] * G_factors[i] % r for i in range(padded_n)]
        if trace is not None:
            trace.update(dict(y=y, z=z, u=u, x=x, w=w, Q=Q, l_vec=list(l_vec), r_vec=list(r_vec),
                              G_factors=G_factors, H_factors=H_factors, n1=n1, n2=n2, padded_n=padded_n,
                              wL=wL, wR=wR, wO=wO, wV=wV, s_L=sL, s_R=sR, t=[t1, t2, t3, t4, t5, t6],
                              blindings=dict(i1=i_bl1, o1=o_bl1, s1=s_bl1, i2=i_bl2, o2=o_bl2, s2=s_bl2,
                                             t=[tb1, tb2, tb3, tb4, tb5, tb6])))
        ipp = ipa_create(cv, t, Q, G_factors, H_factors, bp_gens.G(padded_n), bp_gens.H(padded_n), l_vec, r_vec)
        return R1CSProof(A_I1, A_O1, S1, A_I2, A_O2, S2, T_1, T_3, T_4, T_5, T_6, t_x, t_x_blinding, e_blinding, ipp)


# --------------------------------------------------------------------------
# src/r1cs/verifier.rs
# --------------------------------------------------------------------------
class Verifier:
    def __init__(self, cv: Curve, transcript: Transcript):         # verifier.rs:252-263
        self.cv, self.transcript = cv, transcript
        transcript.append_message(b"dom-sep", b"r1cs v1")
        self.num_vars = 0
        self.V = []
        self.constraints: List[LC] = []
        self.deferred = []
        self.pending_multiplier = None
        self.randomizing = False

    def multiply(self, left, right):                               # verifier.rs:74-98
        left, right = LC.of(left), LC.of(right)
        i = self.num_vars
        self.num_vars += 1
        lv, rv, ov = (MUL_LEFT, i), (MUL_RIGHT, i), (MUL_OUT, i)
        left.terms.append((lv, -1))
        right.terms.append((rv, -1))
        self.constrain(left)
        self.constrain(right)
        return lv, rv, ov

    def allocate(self, _assignment=None):                          # verifier.rs:100-116
        if self.pending_multiplier is None:
            i = self.num_vars
            self.num_vars += 1
            self.pending_multiplier = i
            return (MUL_LEFT, i)
        i = self.pending_multiplier
        self.pending_multiplier = None
        return (MUL_RIGHT, i)

    def allocate_multiplier(self, _assignments=None):              # verifier.rs:118-137
        i = self.num_vars
        self.num_vars += 1
        return (MUL_LEFT, i), (MUL_RIGHT, i), (MUL_OUT, i)

    def multipliers_len(self):
        return self.num_vars

    def constrain(self, lc):
        self.constraints.append(LC.of(lc))

    def specify_randomized_constraints(self, cb):
        self.deferred.append(cb)

    def challenge_scalar(self, label):
        assert self.randomizing
        return challenge_scalar(self.cv, self.transcript, label)

    def commit(self, V):                                           # verifier.rs:279-287
        i = len(self.V)
        self.V.append(V)
        append_point(self.cv, self.transcript, b"V", V)
        return (COMMITTED, i)

    def _flattened_constraints(self, z):                           # verifier.rs:304-349
        r = self.cv.r
        n, m = self.num_vars, len(self.V)
        wL, wR, wO, wV, wc = [0] * n, [0] * n, [0] * n, [0] * m, 0
        exp_z = z
        for lc in self.constraints:
            for (k, i), c in lc.terms:
                if k == MUL_LEFT:
                    wL[i] = (wL[i] + exp_z * c) % r
                elif k == MUL_RIGHT:
                    wR[i] = (wR[i] + exp_z * c) % r
                elif k == MUL_OUT:
                    wO[i] = (wO[i] + exp_z * c) % r
                elif k == COMMITTED:
                    wV[i] = (wV[i] - exp_z * c) % r
                elif k == ONE:
                    wc = (wc - exp_z * c) % r
            exp_z = exp_z * z % r
        return wL, wR, wO, wV, wc

    def _create_randomized_constraints(self):                      # verifier.rs:353-376
        self.pending_multiplier = None
        if not self.deferred:
            self.transcript.append_message(b"dom-sep", b"r1cs-1phase")
        else:
            self.transcript.append_message(b"dom-sep", b"r1cs-2phase")
            cbs, self.deferred = self.deferred, []
            self.randomizing = True
            for cb in cbs:
                cb(self)
            self.randomizing = False

    def verification_scalars(self, proof: R1CSProof, bp_gens: BulletproofGens):   # verifier.rs:394-541
        cv, r, t = self.cv, self.cv.r, self.transcript
        try:
            t.append_u64(b"m", len(self.V))
            n1 = self.num_vars
            validate_and_append_point(cv, t, b"A_I1", proof.A_I1)
            validate_and_append_point(cv, t, b"A_O1", proof.A_O1)
            validate_and_append_point(cv, t, b"S1", proof.S1)
            self._create_randomized_constraints()
            n = self.num_vars
            n2 = n - n1
            padded_n = 1 if n == 0 else 1 << (n - 1).bit_length()
            pad = padded_n - n
            if bp_gens.gens_capacity < padded_n:
                raise R1CSError("InvalidGeneratorsLength")
            append_point(cv, t, b"A_I2", proof.A_I2)
            append_point(cv, t, b"A_O2", proof.A_O2)
            append_point(cv, t, b"S2", proof.S2)
            y = challenge_scalar(cv, t, b"y")
            z = challenge_scalar(cv, t, b"z")
            for lab, P in ((b"T_1", proof.T_1), (b"T_3", proof.T_3), (b"T_4", proof.T_4),
                           (b"T_5", proof.T_5), (b"T_6", proof.T_6)):
                validate_and_append_point(cv, t, lab, P)
            u = challenge_scalar(cv, t, b"u")
            x = challenge_scalar(cv, t, b"x")
            append_scalar(cv, t, b"t_x", proof.t_x)
            append_scalar(cv, t, b"t_x_blinding", proof.t_x_blinding)
            append_scalar(cv, t, b"e_blinding", proof.e_blinding)
            w = challenge_scalar(cv, t, b"w")
            wL, wR, wO, wV, wc = self._flattened_constraints(z)
            u_sq, u_inv_sq, s = ipa_verification_scalars(cv, proof.ipp_proof, padded_n, t)
        except ProofError as e:
            raise R1CSError("VerificationError") from e
        a, b = proof.ipp_proof.a, proof.ipp_proof.b
        y_inv = pow(y, -1, r)
        y_inv_vec = [1] * padded_n
        for i in range(1, padded_n):
            y_inv_vec[i] = y_inv_vec[i - 1] * y_inv % r
        yneg_wR = [wR[i] * y_inv_vec[i] % r for i in range(n)] + [0] * pad
        delta = inner_product(cv, yneg_wR[:n], wL)
        uf = [1] * n1 + [u] * (n2 + pad)
        wLp, wOp = wL + [0] * pad, wO + [0] * pad
        g_scalars = [uf[i] * (x * yneg_wR[i] - a * s[i]) % r for i in range(padded_n)]
        h_scalars = [uf[i] * (y_inv_vec[i] * (x * wLp[i] + wOp[i] - b * s[padded_n - 1 - i]) - 1) % r
                     for i in range(padded_n)]
        rch = challenge_scalar(cv, t.clone(), b"r")                # verifier.rs:516-519
        xx = x * x % r
        rxx = rch * xx % r
        xxx = x * xx % r
        T_scalars = [rch * x % r, rxx * x % r, rxx * xx % r, rxx * xxx % r, rxx * xx % r * xx % r]
        scalars = [(w * (proof.t_x - a * b) + rch * (xx * (wc + delta) - proof.t_x)) % r,
                   (-proof.e_blinding - rch * proof.t_x_blinding) % r]
        scalars += g_scalars + h_scalars
        scalars += [x, xx, xxx, u * x % r, u * xx % r, u * xxx % r]
        scalars += [wVi * rxx % r for wVi in wV]
        scalars += T_scalars + u_sq + u_inv_sq
        return scalars

    def mega_points(self, proof, pc_gens, bp_gens):
        padded_n = 1 if self.num_vars == 0 else 1 << (self.num_vars - 1).bit_length()
        return ([pc_gens.B, pc_gens.B_blinding] + bp_gens.G(padded_n) + bp_gens.H(padded_n) +
                [proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2] + self.V +
                [proof.T_1, proof.T_3, proof.T_4, proof.T_5, proof.T_6] +
                proof.ipp_proof.L_vec + proof.ipp_proof.R_vec)

    def verify(self, proof: R1CSProof, pc_gens: PedersenGens, bp_gens: BulletproofGens):   # verifier.rs:559-600
        scalars = self.verification_scalars(proof, bp_gens)
        pts = self.mega_points(proof, pc_gens, bp_gens)
        if msm(self.cv, pts, scalars) is not None:
            raise R1CSError("VerificationError")


def batch_verify(cv, prng, instances, pc_gens, bp_gens):           # verifier.rs:604-691
    r = cv.r
    max_n = 0
    items = []
    for verifier, proof in instances:
        scalars = verifier.verification_scalars(proof, bp_gens)
        n = 1 if verifier.num_vars == 0 else 1 << (verifier.num_vars - 1).bit_length()
        max_n = max(max_n, n)
        items.append((verifier, proof, scalars, n))
    all_scalars = [0] * (2 * max_n + 2)
    all_elems = [pc_gens.B, pc_gens.B_blinding] + bp_gens.G(max_n) + bp_gens.H(max_n)
    for verifier, proof, scalars, pn in items:
        alpha = scalar_rand(cv, prng)
        sc = [alpha * s % r for s in scalars]
        all_scalars[0] = (all_scalars[0] + sc[0]) % r
        all_scalars[1] = (all_scalars[1] + sc[1]) % r
        for i in range(pn):
            all_scalars[2 + i] = (all_scalars[2 + i] + sc[2 + i]) % r
            all_scalars[2 + max_n + i] = (all_scalars[2 + max_n + i] + sc[2 + pn + i]) % r
        all_scalars += sc[2 + 2 * pn:]
        all_elems += [proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2] + verifier.V + \
                     [proof.T_1, proof.T_3, proof.T_4, proof.T_5, proof.T_6] + \
                     proof.ipp_proof.L_vec + proof.ipp_proof.R_vec
    if msm(cv, all_elems, all_scalars) is not None:
        raise R1CSError("VerificationError")


# --------------------------------------------------------------------------
# Gadgets restated from tests/r1cs_secq256k1.rs (identical for the other curves)
# --------------------------------------------------------------------------
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


def range_proof_gadget(cs, v_lc, v_assignment: Optional[int], n: int):   # tests/r1cs_secq256k1.rs:361-393
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
        v = v - LC.of(b).scale(exp_2)
        exp_2 = exp_2 + exp_2
    cs.constrain(v)


# --------------------------------------------------------------------------
# Synthetic measurement circuit:
# one-phase "public-multiplier chain", N multipliers, m = 1.
# --------------------------------------------------------------------------
def chain_circuit_witness(cv: Curve, N: int, seed: bytes = bytes([3] * 32)):
    rng = ChaCha20Rng(seed)
    x0 = scalar_rand(cv, rng)
    ks = [scalar_rand(cv, rng) for _ in range(N)]
    return x0, ks


def chain_circuit(cs, v_var, N: int, ks: Sequence[int], x0: Optional[int], r: int):
    """(L_i,R_i,O_i) = allocate_multiplier((x_i,k_i)); constrain(R_i - k_i);
    constrain(L_{i+1} - O_i); constrain(L_0 - V_0)."""
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

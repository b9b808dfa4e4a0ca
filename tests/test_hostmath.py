"""CPU unit tests of the device math templates (csrc/fp.cuh, csrc/ec.cuh) compiled for the
host (carry primitives emulated) against the Python oracle's big-integer arithmetic."""
import ctypes
import os
import random
import subprocess

import pytest

import bp_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
R = 1 << 256
FIELDS = [O.SECP_N, O.SECP_P, O.ZORRO.q, 2**255 - 19, O.CURVE25519.r]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hm") / "libhostmath.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-x", "c++",
                           os.path.join(HERE, "native", "hostmath.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


def _b(v):
    return (ctypes.c_uint32 * 8).from_buffer_copy(v.to_bytes(32, "little"))


def _fp(lib, field, op, a, b=0):
    out = (ctypes.c_uint32 * 8)()
    assert lib.hm_fp_op(field, op, _b(a), _b(b), out) == 0
    return int.from_bytes(bytes(out), "little")


@pytest.mark.parametrize("field", list(range(5)) + list(range(10, 15)))
def test_field_ops(lib, field):
    m = FIELDS[field % 10]
    rnd = random.Random(1234 + field)
    edge = [0, 1, 2, m - 1, m - 2, (1 << 255) % m, R % m, (m - 1) // 2, 0xFFFFFFFF, (1 << 224) - 1]
    vals = edge + [rnd.randrange(m) for _ in range(40)]
    rinv = pow(R, -1, m)
    for a in vals:
        for b in rnd.sample(vals, 8) + [m - 1, a]:
            assert _fp(lib, field, 0, a, b) == a * b * rinv % m
            assert _fp(lib, field, 8, a, b) == a * b * rinv % m      # sparse-modulus reduction variant
            assert _fp(lib, field, 1, a, b) == (a + b) % m
            assert _fp(lib, field, 2, a, b) == (a - b) % m
        assert _fp(lib, field, 4, a) == a * rinv % m
        assert _fp(lib, field, 5, a) == a * R % m
        assert _fp(lib, field, 6, a) == (-a) % m
        assert _fp(lib, field, 7, a) == a * a * rinv % m
    for a in vals[:12]:
        am = a * R % m
        inv = _fp(lib, field, 3, am)
        if a == 0:
            assert inv == 0
        else:
            assert inv == pow(a, -1, m) * R % m


@pytest.mark.parametrize("field", list(range(5)) + list(range(10, 15)))
def test_field_fused_products_and_square(lib, field):
    """Fp::mul2 (a*b -/+ c*d with one Montgomery reduction) and the dedicated squaring against big-integer
    arithmetic: edge values that maximise every column (all-ones limbs, m - 1, top-bit patterns) and random ones."""
    m = FIELDS[field % 10]
    rnd = random.Random(99 + field)
    rinv = pow(R, -1, m)
    edge = [0, 1, 2, m - 1, m - 2, (m - 1) // 2, (1 << 255) % m, ((1 << 256) - 1) % m, 0xFFFFFFFF, (1 << 224) - 1,
            int("ffffffff00000000" * 4, 16) % m, int("00000000ffffffff" * 4, 16) % m, int("80000000" * 8, 16) % m,
            int("7fffffff" * 8, 16) % m, m - 0xFFFFFFFF, m - (1 << 128)]
    vals = edge + [rnd.randrange(m) for _ in range(60)]

    def op4(op, a, b, c, d):
        out = (ctypes.c_uint32 * 8)()
        assert lib.hm_fp_op4(field, op, _b(a), _b(b), _b(c), _b(d), out) == 0
        return int.from_bytes(bytes(out), "little")

    for a in vals:
        assert _fp(lib, field, 7, a) == a * a * rinv % m
    for a in edge:
        for b in edge:
            for c, d in [(m - 1, m - 1), (m - 1, 0), (0, m - 1), (a, b), (b, a), (m - 1, 1)]:
                assert op4(0, a, b, c, d) == (a * b - c * d) * rinv % m
                assert op4(1, a, b, c, d) == (a * b + c * d) * rinv % m
    for _ in range(400):
        a, b, c, d = (rnd.choice(vals) for _ in range(4))
        assert op4(0, a, b, c, d) == (a * b - c * d) * rinv % m
        assert op4(1, a, b, c, d) == (a * b + c * d) * rinv % m


def _pt(cv, P):
    if P is None:
        return (ctypes.c_uint32 * 16)()
    x, y = P
    return (ctypes.c_uint32 * 16).from_buffer_copy((x * R % cv.q).to_bytes(32, "little") + (y * R % cv.q).to_bytes(32, "little"))


def _unpt(cv, buf):
    raw = bytes(buf)
    x = int.from_bytes(raw[:32], "little") * pow(R, -1, cv.q) % cv.q
    y = int.from_bytes(raw[32:], "little") * pow(R, -1, cv.q) % cv.q
    return None if (x == 0 and y == 0) else (x, y)


def _ec(lib, cid, cv, op, P, Q, s=0):
    out = (ctypes.c_uint32 * 16)()
    rc = lib.hm_ec_op(cid, op, _pt(cv, P), _pt(cv, Q), _b(s), out)
    assert rc == 0
    return _unpt(cv, out)


@pytest.mark.parametrize("cid,cv", [(0, O.SECQ256K1), (1, O.ZORRO), (2, O.CURVE25519), (10, O.SECQ256K1), (11, O.ZORRO), (12, O.CURVE25519)])
def test_curve_ops(lib, cid, cv):
    rnd = random.Random(99 + cid)
    G = cv.G
    pts = [O.pt_mul(cv, rnd.randrange(1, cv.r), G) for _ in range(4)]
    add, mul, neg = (lambda a, b: O.pt_add(cv, a, b)), (lambda k, a: O.pt_mul(cv, k, a)), (lambda a: O.pt_neg(cv, a))
    cases = [(P, Q) for P in pts[:2] for Q in pts[2:]] + [(pts[0], pts[0]), (pts[0], neg(pts[0])), (pts[0], None), (None, pts[1]), (None, None)]
    for P, Q in cases:
        assert _ec(lib, cid, cv, 0, P, Q) == add(P, Q)
        assert _ec(lib, cid, cv, 1, P, Q) == add(mul(2, P), mul(2, Q))
        assert _ec(lib, cid, cv, 5, P, Q) == add(mul(2, P), mul(2, Q))
        assert _ec(lib, cid, cv, 2, P, Q) == mul(4, P)
        assert _ec(lib, cid, cv, 4, P, Q) == add(mul(2, P), Q)
    # 2P + Q with Q = 2P (mixed add hitting the doubling branch) and Q = -2P
    P = pts[0]
    assert _ec(lib, cid, cv, 4, P, mul(2, P)) == mul(4, P)
    assert _ec(lib, cid, cv, 4, P, neg(mul(2, P))) is None
    for s in [0, 1, 2, cv.r - 1, rnd.randrange(cv.r), rnd.randrange(cv.r)]:
        assert _ec(lib, cid, cv, 3, P, None, s) == mul(s, P)
    for k in [0, 1, 5, 65535, 0xFFFFFFFF]:
        assert _ec(lib, cid, cv, 6, P, None, k) == mul(2 * k, P)


R261 = 1 << 261


@pytest.mark.parametrize("field", [0, 1, 2, 3, 4])
def test_fp29_ops(lib, field):
    """csrc/fp29.cuh: balanced 9 x 29-bit limbs, Montgomery domain 2^261, carry-free columns. Every multiplication in
    the host build also checks the operand contract and the column bound (hm_fp29_violations)."""
    m = FIELDS[field]
    rnd = random.Random(4321 + field)
    rinv = pow(R261, -1, m)
    lib.hm_fp29_violations.restype = ctypes.c_long

    def op(o, a, b=0):
        out = (ctypes.c_uint32 * 8)()
        rc = lib.hm_fp29_op(field, o, _b(a), _b(b), out)
        assert rc >= 0
        return int.from_bytes(bytes(out), "little"), rc
    half = sum(1 << (29 * k + 28) for k in range(8))          # every balanced digit at its extreme
    edge = [0, 1, 2, m - 1, m - 2, (1 << 255) % m, (1 << 232) - 1, (1 << 232), (m - 1) // 2, (1 << 145) - 1, m >> 1, (1 << 29) - 1,
            half % m, (half - 1) % m, (half + 1) % m, (m - half) % m, ((1 << 256) - 1) % m]
    vals = edge + [rnd.randrange(m) for _ in range(60)]
    for a in vals:
        for b in rnd.sample(vals, 6) + [m - 1, a, 0]:
            assert op(0, a, b)[0] == a * b * rinv % m
            assert op(1, a, b)[0] == (a + b) % m
            assert op(2, a, b)[0] == (a - b) % m
            assert op(6, a, b)[0] == ((a - b - b + a) * (a + b + b)) * rinv % m
            assert op(7, a, b)[0] == (((a - b) ** 2 * rinv - 3 * a) * (b - a)) * rinv % m
            d, e = op(9, a, b)
            assert d == (a - b) % m and e == (1 if a == b else 0)
            assert op(14, a, b)[1] == 1
            assert op(13, a, b)[0] == a * (b & 7) % m
            v, z = op(15, a, b)
            assert v == a and z == 1
        assert op(3, a)[0] == a * a * rinv % m
        assert op(4, a)[0] == (-a) % m
        assert op(5, a)[0] == 3 * a % m
        st, z = op(8, a)
        assert st == a * (1 << 256) * rinv % m and z == (1 if a == 0 else 0)
        assert op(10, a)[0] == a * 32 % m
        assert op(11, a)[0] == a
    for a in vals[:8] + vals[-4:]:
        inv = op(12, a)[0]
        assert inv == (0 if a == 0 else pow(a * rinv, -1, m) * R261 % m)
    assert lib.hm_fp29_violations() == 0


@pytest.mark.parametrize("cid,cv", [(0, O.SECQ256K1), (1, O.ZORRO), (2, O.CURVE25519)])
def test_curve_ops_fp29(lib, cid, cv):
    rnd = random.Random(199 + cid)
    q = cv.q
    rinv = pow(R261, -1, q)
    lib.hm_fp29_violations.restype = ctypes.c_long

    def enc(P):
        if P is None:
            return (ctypes.c_uint32 * 16)()
        return (ctypes.c_uint32 * 16).from_buffer_copy((P[0] * R261 % q).to_bytes(32, "little") + (P[1] * R261 % q).to_bytes(32, "little"))

    def run(o, P, Q, s=0):
        out = (ctypes.c_uint32 * 32)()
        rc = lib.hm_ec29_op(cid, o, enc(P), enc(Q), _b(s), out)
        assert rc >= 0
        raw = bytes(out)
        X, Y, ZZ, ZZZ = [int.from_bytes(raw[32 * i:32 * i + 32], "little") * rinv % q for i in range(4)]
        if rc == 1:
            return None
        if cv.kind == "sw":
            return (X * pow(ZZ, -1, q) % q, Y * pow(ZZZ, -1, q) % q)
        x, y = X * pow(ZZ, -1, q) % q, Y * pow(ZZ, -1, q) % q      # extended: (X:Y:Z:T) in (x,y,zz,zzz)
        return None if (x == 0 and y == 1) else (x, y)
    G = cv.G
    pts = [O.pt_mul(cv, rnd.randrange(1, cv.r), G) for _ in range(4)]
    add, mul, neg = (lambda a, b: O.pt_add(cv, a, b)), (lambda k, a: O.pt_mul(cv, k, a)), (lambda a: O.pt_neg(cv, a))
    cases = [(P, Q) for P in pts[:2] for Q in pts[2:]] + [(pts[0], pts[0]), (pts[0], neg(pts[0])), (pts[0], None), (None, pts[1]), (None, None)]
    for P, Q in cases:
        assert run(0, P, Q) == add(P, Q)
        assert run(1, P, Q) == add(mul(2, P), mul(2, Q))
        assert run(5, P, Q) == add(mul(2, P), mul(2, Q))
        assert run(2, P, Q) == mul(4, P)
        assert run(4, P, Q) == add(mul(2, P), Q)
    P = pts[0]
    assert run(4, P, mul(2, P)) == mul(4, P)
    assert run(4, P, neg(mul(2, P))) is None
    for s in [0, 1, 2, cv.r - 1, rnd.randrange(cv.r)]:
        assert run(3, P, None, s) == mul(s, P)
        assert run(7, P, None, s) == mul(s, P)       # through to_affine (one inversion) and on_curve
    # long chains keep the value bounds: P + 32(P + Q) doubled 8 times, times 9
    Q = pts[1]
    want = mul(9 * 256, add(P, mul(32, add(P, Q))))
    assert run(6, P, Q) == want
    assert lib.hm_fp29_violations() == 0


def test_glv_split(lib):
    """secq256k1 GLV decomposition used by the IPA generator fold: k = k1 + k2*lambda (mod r) with both halves below
    2^130 -- for random scalars and the edge values; lambda is the cube root of unity matching beta (x -> beta*x)."""
    r = FIELDS[1]
    lam = 0x7ae96a2b657c07106e64479eac3434e99cf0497512f58995c1396c28719501ee
    beta = 0x5363ad4cc05c30e0a5261c028812645a122e22ea20816678df02967c1b23bd72
    q = FIELDS[0]
    assert pow(lam, 3, r) == 1 and lam != 1 and pow(beta, 3, q) == 1 and beta != 1
    # phi(G) = lambda*G on the curve (checked with the oracle's arithmetic)
    import bp_oracle as O
    cv = O.SECQ256K1
    assert O.pt_mul(cv, lam, cv.G) == (beta * cv.G[0] % q, cv.G[1])
    rnd = random.Random(77)
    vals = [0, 1, 2, r - 1, r - 2, lam, r - lam, (1 << 128), (1 << 255) % r] + [rnd.randrange(r) for _ in range(300)]
    for k in vals:
        out = (ctypes.c_uint32 * 13)()
        assert lib.hm_glv_split(_b(k * R % r), out) == 1
        k1 = sum(out[i] << (32 * i) for i in range(5)) * (-1 if out[10] else 1)
        k2 = sum(out[5 + i] << (32 * i) for i in range(5)) * (-1 if out[11] else 1)
        assert (k1 + k2 * lam - k) % r == 0
        assert abs(k1) < 1 << 130 and abs(k2) < 1 << 130
        top = max(abs(k1).bit_length(), abs(k2).bit_length()) - 1
        assert ctypes.c_int32(out[12]).value == top


def test_joint_sparse_form(lib):
    """jsf_digits (host/glv_host.hpp): the signed digits reproduce both magnitudes, no two consecutive columns are both
    non-zero, and the joint weight is about half of the length (Solinas)."""
    rnd = random.Random(7)
    cases = [(0, 0), (1, 0), (0, 1), (1, 1), (3, 5), ((1 << 130) - 1, (1 << 130) - 1), ((1 << 159) + 1, 1), (11, (1 << 160) - 1)]
    cases += [(rnd.getrandbits(130), rnd.getrandbits(130)) for _ in range(300)]
    cases += [(rnd.getrandbits(rnd.randrange(1, 160)), rnd.getrandbits(rnd.randrange(1, 160))) for _ in range(300)]
    weight = steps = 0
    for k1, k2 in cases:
        a = (ctypes.c_uint32 * 5).from_buffer_copy(k1.to_bytes(20, "little"))
        b = (ctypes.c_uint32 * 5).from_buffer_copy(k2.to_bytes(20, "little"))
        code = (ctypes.c_uint32 * 21)()
        top = lib.hm_jsf(a, b, code)
        assert top >= -1
        u = [((code[j >> 3] >> (4 * (j & 7))) & 3) - 1 for j in range(top + 1)]
        v = [((code[j >> 3] >> (4 * (j & 7) + 2)) & 3) - 1 for j in range(top + 1)]
        assert all(x in (-1, 0, 1) for x in u + v)
        assert sum(x << j for j, x in enumerate(u)) == k1
        assert sum(x << j for j, x in enumerate(v)) == k2
        nzc = [bool(x or y) for x, y in zip(u, v)]
        if k1.bit_length() > 100 and k2.bit_length() > 100:
            weight += sum(nzc)
            steps += top + 1
    assert 0.45 < weight / steps < 0.55


def test_bucket_sort_plan(lib):
    """make_sort_plan (csrc/msm_sort.cuh) for every window width the MSM planner can choose and sizes from 2^12 to 2^26:
    at most 1024 coarse bins and 1024 low-bit counters, bins that fit the shared-memory stage with a margin, the last
    window's bins scaled to the range its digits really span, and a refusal (library sort) everywhere else."""
    out = (ctypes.c_uint32 * 8)()
    accepted = 0
    for c in range(4, 21):
        W, cb = 256 // c + 1, c - 1
        for n in [1 << k for k in range(12, 27)] + [(1 << 20) + 12345, 3 * (1 << 20), (1 << 24) - 1, (1 << 24) + 1]:
            lib.hm_sort_plan(ctypes.c_uint64(n), W, cb, out)
            ok, low, low_top, nb1, tiles, cap, ts = out[0], out[1], out[2], out[3], out[4], out[5], out[6]
            if not ok:
                continue
            accepted += 1
            assert nb1 & (nb1 - 1) == 0 and 1 <= nb1 <= 1024
            assert 0 <= low <= 10 and nb1 << low == 1 << cb
            assert tiles == -(-n // ts)
            assert n // nb1 <= cap * 4 // 5                    # average bin fits the stage with 25 % to spare
            top_bits = max(0, min(cb, 256 - (W - 1) * c))      # the last window's buckets are < 2^top_bits
            assert low_top <= low and (nb1 << low_top) >= (1 << top_bits)   # its coarse bin index stays below nb1
            if top_bits > low_top:                             # ... and it uses as many of them as its range allows
                assert (1 << (top_bits - low_top)) == min(nb1, 1 << top_bits)
    lib.hm_sort_plan(ctypes.c_uint64(1 << 24), 13, 19, out)
    assert list(out)[:5] == [1, 9, 6, 1024, (1 << 24) // out[6]]   # the headline plan: 1024 bins of 16384 pairs, 512 low-bit counters
    lib.hm_sort_plan(ctypes.c_uint64(1 << 25), 13, 19, out)
    assert out[0] == 0                                         # bins would overflow the stage: library sort
    assert accepted > 100


def test_host_workers(lib):
    """host/workers.hpp: batches of 1..8 jobs on the persistent workers; every job runs exactly once and run() returns only
    after the last one (the Horner chains of a batch of MSMs write into the caller's stack)."""
    assert lib.hm_workers_stress(5000) == 0

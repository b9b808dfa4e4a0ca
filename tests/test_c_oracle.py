"""The C restatement (oracle/c/bp_ref.c: ark-style wNAF Pippenger, Jacobian formulas, the
reference's per-element fold) must agree with the Python restatement."""
import random

import pytest

import bp_oracle as O
import c_oracle
from ark_bulletproofs_b200 import codec


def _pts(cv, n, rnd):
    base = O.pt_mul(cv, rnd.randrange(1, cv.r), cv.G)
    pts, P = [], None
    for _ in range(n):
        P = O.pt_add(cv, P, base)
        pts.append(P)
    return pts


@pytest.mark.parametrize("name,cid", [("secq256k1", 0), ("zorro", 1)])
@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 300])
def test_c_msm_matches_python(name, cid, n):
    cv = O.CURVES[name]
    rnd = random.Random(n * 7 + cid)
    pts = _pts(cv, n, rnd)
    if n > 4:
        pts[3] = None
        pts[1] = pts[2]
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    sc[0] = cv.r - 1
    if n > 2:
        sc[2] = 0
    for threads in (1, 4):
        out = c_oracle.msm_bytes(cid, codec.enc_points(pts, name), codec.enc_scalars(sc, name), n, threads)
        assert codec.dec_point(out, name) == O.msm(cv, pts, sc)


def test_c_fold_matches_reference_loop():
    cv = O.SECQ256K1
    rnd = random.Random(3)
    h = 4
    pts = _pts(cv, 2 * h, rnd)
    u = rnd.randrange(1, cv.r)
    ui = pow(u, -1, cv.r)
    buf = bytearray(codec.enc_points(pts, "secq256k1"))
    c_oracle.fold_points(0, buf, h, codec.enc_scalars([ui], "secq256k1"), codec.enc_scalars([u], "secq256k1"), 2)
    for i in range(h):
        want = O.pt_add(cv, O.pt_mul(cv, ui, pts[i]), O.pt_mul(cv, u, pts[h + i]))   # inner_product_proof.rs:219-221
        assert codec.dec_point(bytes(buf[64 * i:64 * i + 64]), "secq256k1") == want


def test_c_synth_points():
    cv = O.SECQ256K1
    out = c_oracle.synth_points(0, codec.enc_point(cv.G, "secq256k1"), 1500, 5)
    for i in (0, 1, 1023, 1024, 1499):
        assert codec.dec_point(bytes(out[64 * i:64 * i + 64]), "secq256k1") == O.pt_mul(cv, 5 + i + 1, cv.G)


# ---- oracle/fast.py: the three primitives swapped in for the large goldens, each against its Python original ----
@pytest.mark.parametrize("name", ["secq256k1", "zorro"])
def test_fast_trng_scalars_match_python(name):
    """merlin TranscriptRng + Fr::rand in C (bp_ref.c ref_trng_scalars) == bp_oracle.TranscriptRng/fp_rand, including the
    STROBE state after the draws (zorro's Fr = 2^255 - 19 shaves one bit; rejections are astronomically rare there,
    the curve25519 scalar field that rejects half the draws has no C curve -- its rejection loop is the same code)."""
    import fast
    cv = O.CURVES[name]

    def make():
        t = O.Transcript(b"fast-test")
        t.append_message(b"x", b"hello")
        b = t.build_rng().rekey_with_witness_bytes(b"v_blinding", bytes(range(32)))
        return b.finalize(O.ChaCha20Rng(bytes(range(32))))
    a, b = make(), make()
    want = [O.fp_rand(cv.r, a) for _ in range(45)]
    got = fast.trng_scalars(cv.r, b, 20) + fast.trng_scalars(cv.r, b, 25)
    assert got == want
    assert bytes(a.strobe.state) == bytes(b.strobe.state) and (a.strobe.pos, a.strobe.pos_begin) == (b.strobe.pos, b.strobe.pos_begin)
    assert O.fp_rand(cv.r, a) == fast.trng_scalars(cv.r, b, 1)[0]


def test_fast_trng_rejection_path():
    """A modulus that rejects ~15/16 of the draws drives the rejection loop of ref_trng_scalars."""
    import fast
    m = (1 << 252) + 27742317777372353535851937790883648493       # curve25519's scalar field: 253 bits, rejects ~1/2
    t = O.Transcript(b"rej")
    a = t.build_rng().finalize(O.ChaCha20Rng(bytes([7] * 32)))
    b = t.build_rng().finalize(O.ChaCha20Rng(bytes([7] * 32)))
    assert [O.fp_rand(m, a) for _ in range(30)] == fast.trng_scalars(m, b, 30)


def test_fast_fold_and_msm_match_python():
    import fast
    cv = O.SECQ256K1
    rnd = random.Random(11)
    n = 24
    pts = _pts(cv, 2 * n, rnd)
    sL = [rnd.randrange(cv.r) for _ in range(n)]
    sR = [rnd.randrange(cv.r) for _ in range(n)]
    sL[3], sR[4] = 0, 1
    assert fast.fold_generators(cv, pts[:n], pts[n:], sL, sR) == O.fold_generators(cv, pts[:n], pts[n:], sL, sR)
    sc = [rnd.randrange(cv.r) for _ in range(2 * n)]
    fast._orig.setdefault("msm", O.msm)
    assert fast.msm(cv, pts, sc) == O.msm(cv, pts, sc)


def test_fast_oracle_reproduces_golden_proof():
    """With the C primitives installed the oracle emits the same bytes as pure Python (golden chain proof, 64 multipliers)."""
    import hashlib
    import json
    import os
    import fast
    import oracle_cases as C
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs.json")))
    name = "chain100"
    g = gold[name]
    cv = O.CURVES[g["curve"]]
    fast.install()
    try:
        pc, bp = O.PedersenGens(cv), O.BulletproofGens(cv, g["gens_capacity"], 1)
        proof, _ = C.oracle_prove_case(g["kind"], g["params"], cv, pc, bp)
        assert hashlib.sha256(proof.to_bytes(cv)).hexdigest() == g["sha256"]
    finally:
        fast.uninstall()


def test_fast_parallel_gens_match_serial():
    import fast
    cv = O.SECQ256K1
    a = fast.parallel_gens(cv, 300, procs=3)
    b = O.BulletproofGens(cv, 300, 1)
    assert a.G(300) == b.G(300) and a.H(300) == b.H(300)

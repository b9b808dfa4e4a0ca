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

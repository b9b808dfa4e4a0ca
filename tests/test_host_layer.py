"""CPU tests of the C++ host layer behind the C ABI (no GPU needed): Merlin/STROBE/Keccak,
ChaCha20Rng, Fr::rand, challenge_scalar, generator generation, (de)serialisation -- all against
the oracle and the public KATs."""
import ctypes

import pytest

import bp_oracle as O
from ark_bulletproofs_b200 import _lib, codec
from ark_bulletproofs_b200 import r1cs as R

cv = O.SECQ256K1


def test_merlin_kat():
    t = R.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_transcript_matches_oracle_long():
    # crosses the STROBE rate boundary (166) several times, clones, u64s
    a, b = R.Transcript(b"x" * 7), O.Transcript(b"x" * 7)
    for i in range(40):
        msg = bytes((i * 7 + j) % 256 for j in range(3 + 11 * i))
        a.append_message(b"lbl%d" % i, msg)
        b.append_message(b"lbl%d" % i, msg)
        a.append_u64(b"n", i * 1000003)
        b.append_u64(b"n", i * 1000003)
        if i % 5 == 0:
            assert a.challenge_bytes(b"c", 64 + i) == b.challenge_bytes(b"c", 64 + i)
    c1, c2 = a.clone(), b.clone()
    assert c1.challenge_bytes(b"z", 200) == c2.challenge_bytes(b"z", 200)
    assert a.challenge_bytes(b"z", 200) == b.challenge_bytes(b"z", 200)


@pytest.mark.parametrize("curve", ["secq256k1", "zorro", "curve25519"])
def test_challenge_scalar_and_rng(curve):
    c = O.CURVES[curve]
    t1, t2 = R.Transcript(b"kat"), O.Transcript(b"kat")
    assert t1.challenge_scalar(curve, b"c") == O.challenge_scalar(c, t2, b"c")
    r1, r2 = R.ChaChaRng(bytes(range(32))), O.ChaCha20Rng(bytes(range(32)))
    for _ in range(20):
        assert r1.scalar(curve) == O.scalar_rand(c, r2)
    assert r1.words_used == r2.words_used


def test_kat_challenge_appendix_b():
    t = R.Transcript(b"kat")
    v = t.challenge_scalar("secq256k1", b"c")
    assert v.to_bytes(32, "little").hex() == "4d4c04d1ab63e641ff5fe0f5b50d61c58107288fb2a92ed3558d7cd9ec97692f"


@pytest.mark.parametrize("curve,cap", [("secq256k1", 300), ("zorro", 20), ("curve25519", 12)])
def test_generators_match_oracle(curve, cap):
    # cap = 300 exercises the threaded seek path on secq256k1 (count >= 256)
    c = O.CURVES[curve]
    B, Bb, G, H = R.generate_gens_host(curve, cap)
    pc = O.PedersenGens(c)
    assert B == pc.B and Bb == pc.B_blinding
    bp = O.BulletproofGens(c, cap, 1)
    assert G == bp.G(cap) and H == bp.H(cap)


def test_serialisation_roundtrip():
    lib = _lib.load()
    P = O.pt_mul(cv, 123456789, cv.G)
    for pt in (P, O.pt_neg(cv, P), None):
        out = ctypes.create_string_buffer(33)
        assert lib.bp_point_compress(0, codec.enc_point(pt, "secq256k1"), out) == 0
        assert out.raw == O.ser_point(cv, pt, True)
        unc = ctypes.create_string_buffer(65)
        assert lib.bp_point_serialize_uncompressed(0, codec.enc_point(pt, "secq256k1"), unc) == 0
        assert unc.raw == O.ser_point(cv, pt, False)
        back = ctypes.create_string_buffer(64)
        assert lib.bp_point_decompress(0, out.raw, back) == 0
        assert codec.dec_point(back.raw, "secq256k1") == pt
    # not on curve / bad flags / x >= q -> FormatError
    bad = bytearray(O.ser_point(cv, P, True))
    for x in range(1, 50):
        cand = (x).to_bytes(32, "little") + b"\x00"
        if O.sqrt_mod((x ** 3 + 7) % cv.q, cv.q) is None:
            assert lib.bp_point_decompress(0, cand, ctypes.create_string_buffer(64)) == -8
            break
    # ark-ff reads the integer from the first 32 bytes only: the six padding bits of the flag byte are ignored, both
    # flag bits together are UnexpectedFlags, and the infinity flag gives the identity whatever x is
    good = O.ser_point(cv, P, True)
    for pad in (0x01, 0x3F):
        back = ctypes.create_string_buffer(64)
        enc = good[:32] + bytes([good[32] | pad])
        assert lib.bp_point_decompress(0, enc, back) == 0
        assert codec.dec_point(back.raw, "secq256k1") == P == O.de_point_compressed(cv, enc)
    assert lib.bp_point_decompress(0, good[:32] + b"\xC0", ctypes.create_string_buffer(64)) == -8
    back = ctypes.create_string_buffer(64)
    assert lib.bp_point_decompress(0, good[:32] + b"\x40", back) == 0 and codec.dec_point(back.raw, "secq256k1") is None
    assert O.de_point_compressed(cv, good[:32] + b"\x40") is None
    assert lib.bp_point_decompress(0, b"\xff" * 32 + b"\x00", ctypes.create_string_buffer(64)) == -8
    s = ctypes.create_string_buffer(32)
    assert lib.bp_scalar_from_bytes(0, (cv.r).to_bytes(32, "little"), s) == -8
    assert lib.bp_scalar_from_bytes(0, (cv.r - 1).to_bytes(32, "little"), s) == 0
    assert codec.dec_fe(s.raw, cv.r) == cv.r - 1


def test_proof_from_bytes_matches_oracle_bytes():
    import oracle_cases as C
    pc, bp = O.PedersenGens(cv), O.BulletproofGens(cv, 8, 1)
    proof, _, _ = C.prove_shuffle(cv, pc, bp, [5, 9, 2], [2, 5, 9])
    b = proof.to_bytes(cv)
    p = R.Proof.from_bytes("secq256k1", b)
    assert p.to_bytes() == b and p.rounds() == 2
    assert p.get_scalar(0) == proof.t_x and p.get_point(100) == proof.ipp_proof.L_vec[0]
    for cut in (0, 10, 33 * 11 + 5, len(b) - 1):
        with pytest.raises(R.BpError):
            R.Proof.from_bytes("secq256k1", b[:cut])
    corrupt = bytearray(b)
    corrupt[33 * 11 + 96] = 0xFF        # L length prefix -> huge
    with pytest.raises(R.BpError):
        R.Proof.from_bytes("secq256k1", bytes(corrupt))


def test_curve25519_serialisation():
    """TE points: 32-byte compressed (y + sign of x), 64-byte uncompressed, identity = (0,1); decompression
    checks the curve equation and the prime-order subgroup."""
    c = O.CURVE25519
    lib = _lib.load()
    P = O.pt_mul(c, 987654321, c.G)
    for pt in (P, O.pt_neg(c, P), None):
        out = ctypes.create_string_buffer(33)
        assert lib.bp_point_compress(2, codec.enc_point(pt, "curve25519"), out) == 0
        assert out.raw[:32] == O.ser_point(c, pt, True)
        unc = ctypes.create_string_buffer(65)
        assert lib.bp_point_serialize_uncompressed(2, codec.enc_point(pt, "curve25519"), unc) == 0
        assert unc.raw[:64] == O.ser_point(c, pt, False)
        back = ctypes.create_string_buffer(64)
        assert lib.bp_point_decompress(2, out.raw[:32], back) == 0
        assert codec.dec_point(back.raw, "curve25519") == pt
    # a point of small order (order 8 component) must be rejected: (x, y) with y = 0 -> x^2 = -1... use a torsion point
    # 8-torsion point of ed25519: y = 0x7a03ac9277fdc74ec6cc392cfa53202a0f67100d760b3cba4fd84d3d706a17c7 (order 8)
    y8 = 0x7a03ac9277fdc74ec6cc392cfa53202a0f67100d760b3cba4fd84d3d706a17c7
    raw = y8.to_bytes(32, "little")
    assert lib.bp_point_decompress(2, raw, ctypes.create_string_buffer(64)) == -8
    with pytest.raises(ValueError):
        O.de_point_compressed(c, raw)


def _oracle_trng(curve, nwit):
    t = O.Transcript(b"rngtest")
    b = t.build_rng()
    for i in range(nwit):
        b.rekey_with_witness_bytes(b"v_blinding", bytes([i + 1]) * 32)
    return b.finalize(O.ChaCha20Rng(bytes(range(32))))


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("curve", ["secq256k1", "curve25519"])
def test_transcript_rng_matches_oracle(impl, curve):
    """merlin TranscriptRng behind Prover::prove (prover.rs:483-513): single draws, the register-resident bulk
    path and its rejection handling (curve25519's Fr rejects about half of the draws), scalar and AVX-512 Keccak."""
    lib = _lib.load()
    if lib.bp_host_keccak_select(impl) != impl:
        pytest.skip("no AVX-512 on this host")
    try:
        c = O.CURVES[curve]
        want = _oracle_trng(curve, 2)
        got = R.Transcript(b"rngtest").build_rng(b"v_blinding", [bytes([1]) * 32, bytes([2]) * 32], R.ChaChaRng(bytes(range(32))))
        for _ in range(3):
            assert got.next_u64() == want.next_u64()
        for _ in range(3):
            assert got.scalar(curve) == O.scalar_rand(c, want)
        raw = got.scalars_raw(curve, 37)
        r = codec.MODULI[curve][1]
        assert [codec.dec_fe(raw[32 * i:32 * i + 32], r) for i in range(37)] == [O.scalar_rand(c, want) for _ in range(37)]
        assert got.next_u64() == want.next_u64()
        # transcript challenges go through the same permutation
        t1, t2 = R.Transcript(b"kat"), O.Transcript(b"kat")
        assert t1.challenge_bytes(b"c", 300) == t2.challenge_bytes(b"c", 300)
    finally:
        lib.bp_host_keccak_select(1)

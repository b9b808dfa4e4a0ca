"""Pin the CPU oracle (oracle/bp_oracle.py) against every anchor we have:
public KATs and SURVEY.md Appendix B (the reference itself pins no bytes:
all its tests use thread_rng, tests/r1cs_secq256k1.rs:140,243,...)."""
import hashlib

import pytest

import bp_oracle as O
import oracle_cases as C

cv = O.SECQ256K1


@pytest.fixture(scope="module")
def gens():
    return O.PedersenGens(cv), O.BulletproofGens(cv, 128, 1)


def test_merlin_kat():
    t = O.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == \
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_chacha20_block0():
    assert O._chacha_block([0] * 8, 0)[:4] == [0xADE0B876, 0x903DF1A0, 0xE56A5D40, 0x28BD8653]


def test_sha3_keccak_consistency():
    # our Keccak-f against hashlib's SHA3-256 of the empty string
    st = bytearray(200)
    st[0] ^= 0x06
    st[135] ^= 0x80
    O._permute_bytes(st)
    assert bytes(st[:32]) == hashlib.sha3_256(b"").digest()


@pytest.mark.parametrize("c", [O.SECQ256K1, O.ZORRO, O.CURVE25519])
def test_curve_sanity(c):
    assert O.on_curve(c, c.G)
    assert O.pt_mul(c, c.r, c.G) is None
    P = O.pt_mul(c, 12345, c.G)
    assert O.pt_add(c, P, O.pt_neg(c, P)) is None
    assert O.msm(c, [c.G, P], [5, 7]) == O.pt_mul(c, 5 + 7 * 12345, c.G)


def test_appendix_b_generators(gens):
    pc, bp = gens
    h = lambda P: O.ser_point(cv, P, True).hex()
    assert h(pc.B) == "a6ed0277e38842a2a68177095ae43431e232cea2876cb0b60e16cb85559fc37600"
    assert h(pc.B_blinding) == "e263e9c38a1ecf41f1065f7060c1ddbefa63b034540ba8ecbd308e32ac964ff700"
    assert [h(P) for P in bp.G(4)] == [
        "d8dccd81a021e31a8ef4e6d56191ced2fbcb57a87170d916ddf2aac86ba4ece280",
        "6de790bca5b3f8c199806006e8ed4b8c3b09b15bd0ede3e5df20d5ffb8573ad500",
        "095270a928843bef8e029d9cada3f6792161692449ab753217ca2f7712cf744980",
        "b695337666c81aa5fa877ed4529464637aa5f781dc7b28909f95b2885225c4f800"]
    assert [h(P) for P in bp.H(4)] == [
        "d1d7941e554d89e991cbabec869eb785ce0c8517921e2c2112c18fce170a729d00",
        "71b80fd7646b5bfd8c2b18c2e805a959d73470d1477f8b876cb131fdc0032fda00",
        "5a6fd7995db7d04f790a39eabcebca3e7eb34db85918dd4d6286745c8bc3e17980",
        "f1c1eec966890ab0b910c2e372805c4b327e847d6d92d6f47760e321360b767f80"]
    t = O.Transcript(b"kat")
    assert O.ser_scalar(cv, O.challenge_scalar(cv, t, b"c")).hex() == \
        "4d4c04d1ab63e641ff5fe0f5b50d61c58107288fb2a92ed3558d7cd9ec97692f"


def test_generators_extension(gens):
    # src/generators.rs:354-376: increase_capacity(32 -> 64) == new(64)
    g = O.BulletproofGens(cv, 8, 1)
    g.increase_capacity(16)
    assert g.G(16) == gens[1].G(16) and g.H(16) == gens[1].H(16)


def test_v1_example_gadget(gens):
    pc, bp = gens
    rng = C.seed_a_rng()
    proof, coms = C.prove_example(cv, pc, bp, 9, rng)
    b = proof.to_bytes(cv)
    assert len(b) == 539 and rng.words_used == 48
    assert hashlib.sha256(b).hexdigest() == "765eaf6d94e0cd313db2691a02e39583483aa765bea4250d1b8032be830a83d5"
    assert O.ser_point(cv, coms[0], True).hex() == "a0049bd7772de81349ab4cd1c4546ad96a09ab93ab5a543fd5fe72c73b1538bb00"
    C.verify_example(cv, pc, bp, O.R1CSProof.from_bytes(cv, b), coms, 9)
    with pytest.raises(O.R1CSError):          # tests/r1cs_secq256k1.rs:347
        C.verify_example(cv, pc, bp, proof, coms, 10)
    # c2=10 yields the same proof bytes (prover ignores constants, prover.rs:387-389)
    assert C.prove_example(cv, pc, bp, 10)[0].to_bytes(cv) == b


def test_v2_shuffle3(gens):
    pc, bp = gens
    proof, ic, oc = C.prove_shuffle(cv, pc, bp, [5, 9, 2], [2, 5, 9])
    b = proof.to_bytes(cv)
    assert len(b) == 671
    assert hashlib.sha256(b).hexdigest() == "a51ae53f5e0fac5dc5bb9d5f520615763e3e845b66b6c4179645b8c01ee32151"
    C.shuffle_verifier(cv, ic, oc).verify(proof, pc, bp)


def test_v3_range8(gens):
    pc, bp = gens
    proof, com = C.prove_range(cv, pc, bp, 0xA5, 8)
    b = proof.to_bytes(cv)
    assert len(b) == 737
    assert hashlib.sha256(b).hexdigest() == "d54cb874f445db631163dd4ce11056658f7bfdea5f9f88783af266c73bee658e"
    C.range_verifier(cv, com, 8).verify(proof, pc, bp)


def test_behaviour_matrix(gens):
    pc, bp = gens
    # 4-bit range of 16 rejected (tests/r1cs_secq256k1.rs:409)
    proof, com = C.prove_range(cv, pc, bp, 16, 4)
    with pytest.raises(O.R1CSError):
        C.range_verifier(cv, com, 4).verify(proof, pc, bp)
    # 4-shuffle: 6 multipliers -> padded 8
    proof, ic, oc = C.prove_shuffle(cv, pc, bp, [1, 2, 3, 4], [4, 2, 1, 3])
    C.shuffle_verifier(cv, ic, oc).verify(proof, pc, bp)
    assert len(proof.ipp_proof.L_vec) == 3
    # tampering
    import copy
    for fld in ("t_x", "e_blinding"):
        bad = copy.deepcopy(proof)
        setattr(bad, fld, (getattr(bad, fld) + 1) % cv.r)
        with pytest.raises(O.R1CSError):
            C.shuffle_verifier(cv, ic, oc).verify(bad, pc, bp)
    bad = copy.deepcopy(proof)
    bad.ipp_proof.a = (bad.ipp_proof.a + 1) % cv.r
    with pytest.raises(O.R1CSError):
        C.shuffle_verifier(cv, ic, oc).verify(bad, pc, bp)
    bad = copy.deepcopy(proof)
    bad.A_I1 = None                         # identity -> validate_and_append_point error
    with pytest.raises(O.R1CSError):
        C.shuffle_verifier(cv, ic, oc).verify(bad, pc, bp)


def test_batch_verify(gens):
    # tests/r1cs_secq256k1.rs:447-475 (mixed sizes; one invalid member rejects)
    pc, bp = gens
    good = [(0, 4), (3, 4), (15, 4), (16, 8)]
    items = [C.prove_range(cv, pc, bp, v, n, O.ChaCha20Rng(bytes([i] * 32))) + (n,) for i, (v, n) in enumerate(good)]
    inst = [(C.range_verifier(cv, com, n), proof) for proof, com, n in items]
    O.batch_verify(cv, O.ChaCha20Rng(bytes([5] * 32)), inst, pc, bp)
    bad = [(0, 4), (16, 4), (16, 8)]
    items = [C.prove_range(cv, pc, bp, v, n, O.ChaCha20Rng(bytes([i] * 32))) + (n,) for i, (v, n) in enumerate(bad)]
    inst = [(C.range_verifier(cv, com, n), proof) for proof, com, n in items]
    with pytest.raises(O.R1CSError):
        O.batch_verify(cv, O.ChaCha20Rng(bytes([5] * 32)), inst, pc, bp)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_ipa_roundtrip(gens, n):
    # src/inner_product_proof.rs:407-553 (make_ipp_*)
    pc, bp = gens
    rng = O.ChaCha20Rng(bytes([7] * 32))
    seed = hashlib.sha3_512(b"test point").digest()[:32]
    Q = O.affine_rand(cv, O.ChaCha20Rng(seed))
    a = [O.scalar_rand(cv, rng) for _ in range(n)]
    b = [O.scalar_rand(cv, rng) for _ in range(n)]
    y_inv = O.scalar_rand(cv, rng)
    Gf = [1] * n
    Hf = [pow(y_inv, i, cv.r) for i in range(n)]
    c = O.inner_product(cv, a, b)
    P = O.msm(cv, bp.G(n) + bp.H(n) + [Q], a + [x * h % cv.r for x, h in zip(b, Hf)] + [c])
    proof = O.ipa_create(cv, O.Transcript(b"innerproducttest"), Q, Gf, Hf, bp.G(n), bp.H(n), a, b)
    O.ipa_verify(cv, proof, n, O.Transcript(b"innerproducttest"), Gf, Hf, P, Q, bp.G(n), bp.H(n))


def test_small_kats():
    # util.rs:147-166, inner_product_proof.rs:556-562
    assert O.inner_product(cv, [1, 2, 3, 4], [2, 3, 4, 5]) == 40


@pytest.mark.parametrize("c", [O.ZORRO, O.CURVE25519])
def test_other_curves_roundtrip(c):
    pc, bp = O.PedersenGens(c), O.BulletproofGens(c, 8, 1)
    proof, coms = C.prove_example(c, pc, bp, 9)
    b = proof.to_bytes(c)
    C.verify_example(c, pc, bp, O.R1CSProof.from_bytes(c, b), coms, 9)
    with pytest.raises(O.R1CSError):
        C.verify_example(c, pc, bp, proof, coms, 10)

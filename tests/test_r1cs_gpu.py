"""GPU parity of the IPA / R1CS prover / verifier / batch verifier (C ABI -> C++ host -> CUDA)
against the CPU oracle: byte-identical proofs under the SEED-A RNG convention, proofs verify under
the oracle's verifier, and accept/reject decisions match on valid and tampered proofs.
Mirrors tests/r1cs_secq256k1.rs:131-475 and src/inner_product_proof.rs:407-553."""
import hashlib
import json
import os
import random

import pytest

import bp_oracle as O
import oracle_cases as C

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "proofs.json")))


@pytest.fixture(scope="module")
def env():
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    ctxs, gens = {}, {}

    def get(curve, cap):
        if curve not in ctxs:
            ctxs[curve] = Context(curve, 0)
        key = (curve, cap)
        if key not in gens:
            gens[key] = R.Gens(ctxs[curve], cap)
        return ctxs[curve], gens[key]
    return get


def seed_a():
    from ark_bulletproofs_b200 import r1cs as R
    return R.ChaChaRng(bytes(range(32)))


def gpu_prove_case(R, ctx, gens, kind, params, curve):
    cv = O.CURVES[curve]
    rng = seed_a()
    if kind == "example":
        p = R.Prover(ctx, gens, R.Transcript(b"R1CSExampleGadget"))
        cvs = [p.commit(v, rng.scalar(curve)) for v in (3, 4, 6, 1, 40)]
        R.example_gadget(p, *[v for _, v in cvs], 9)
        return p.prove(rng), [V for V, _ in cvs]
    if kind in ("shuffle", "shuffle_fixed"):
        inp, out = (params["inp"], params["out"]) if kind == "shuffle_fixed" else C.shuffle_values(params["k"], params["seed"])
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", len(inp))
        p = R.Prover(ctx, gens, t)
        ic = [p.commit(v, rng.scalar(curve)) for v in inp]
        oc = [p.commit(v, rng.scalar(curve)) for v in out]
        R.shuffle_gadget(p, [v for _, v in ic], [v for _, v in oc])
        return p.prove(rng), [V for V, _ in ic] + [V for V, _ in oc]
    if kind == "range":
        p = R.Prover(ctx, gens, R.Transcript(b"RangeProofTest"))
        com, var = p.commit(params["value"], rng.scalar(curve))
        R.range_proof_gadget(p, var, params["value"], params["bits"])
        return p.prove(rng), [com]
    if kind == "chain":
        x0, ks = O.chain_circuit_witness(cv, params["N"])
        p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
        com, var = p.commit(x0, rng.scalar(curve))
        R.chain_circuit(p, var, params["N"], ks, x0, cv.r)
        return p.prove(rng), [com]
    raise ValueError(kind)


def gpu_verifier(R, ctx, kind, params, curve, coms):
    cv = O.CURVES[curve]
    if kind == "example":
        v = R.Verifier(ctx, R.Transcript(b"R1CSExampleGadget"))
        vs = [v.commit(V) for V in coms]
        R.example_gadget(v, *vs, params.get("c2", 9))
        return v
    if kind in ("shuffle", "shuffle_fixed"):
        k = len(coms) // 2
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", k)
        v = R.Verifier(ctx, t)
        iv = [v.commit(V) for V in coms[:k]]
        ov = [v.commit(V) for V in coms[k:]]
        R.shuffle_gadget(v, iv, ov)
        return v
    if kind == "range":
        v = R.Verifier(ctx, R.Transcript(b"RangeProofTest"))
        var = v.commit(coms[0])
        R.range_proof_gadget(v, var, None, params["bits"])
        return v
    if kind == "chain":
        _, ks = O.chain_circuit_witness(cv, params["N"])
        v = R.Verifier(ctx, R.Transcript(b"ChainCircuit"))
        var = v.commit(coms[0])
        R.chain_circuit(v, var, params["N"], ks, None, cv.r)
        return v
    raise ValueError(kind)


def oracle_verifier(kind, params, curve, coms):
    cv = O.CURVES[curve]
    if kind == "example":
        v = O.Verifier(cv, O.Transcript(b"R1CSExampleGadget"))
        vs = [v.commit(V) for V in coms]
        O.example_gadget(v, *vs, params.get("c2", 9))
        return v
    if kind in ("shuffle", "shuffle_fixed"):
        k = len(coms) // 2
        return C.shuffle_verifier(cv, coms[:k], coms[k:])
    if kind == "range":
        return C.range_verifier(cv, coms[0], params["bits"])
    _, ks = O.chain_circuit_witness(cv, params["N"])
    return C.chain_verifier(cv, coms[0], params["N"], ks)


def test_generators_on_device(env):
    from ark_bulletproofs_b200 import r1cs as R
    ctx, gens = env("secq256k1", 128)
    cv = O.SECQ256K1
    bp = O.BulletproofGens(cv, 128, 1)
    pc = O.PedersenGens(cv)
    assert gens.export(0, 0, 128) == bp.G(128) and gens.export(1, 0, 128) == bp.H(128)
    assert gens.export(2, 0, 2) == [pc.B, pc.B_blinding]
    assert gens.commit(5, 7) == pc.commit(5, 7)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 32, 64])
@pytest.mark.parametrize("factors", ["ones_geometric", "random"])
@pytest.mark.parametrize("nofold", [0, 8, 1 << 14])
def test_ipa_create_matches_oracle(env, n, factors, nofold):
    """nofold = 0: every round folds the generators (joint first round, uniform-scalar later rounds);
    8: folds down to 8 then switches to MSMs over the stage generators; 2^14: never folds."""
    from ark_bulletproofs_b200 import r1cs as R
    cv = O.SECQ256K1
    ctx, _ = env("secq256k1", 128)
    ctx.set_ipa_nofold_threshold(nofold)
    bp = O.BulletproofGens(cv, 64, 1)
    rnd = random.Random(n * 3 + len(factors))
    Q = O.affine_rand(cv, O.ChaCha20Rng(hashlib.sha3_512(b"test point").digest()[:32]))   # inner_product_proof.rs:422-433
    a = [rnd.randrange(cv.r) for _ in range(n)]
    b = [rnd.randrange(cv.r) for _ in range(n)]
    y_inv = rnd.randrange(1, cv.r)
    if factors == "random":
        Gf = [rnd.randrange(1, cv.r) for _ in range(n)]
        Hf = [rnd.randrange(1, cv.r) for _ in range(n)]
    else:
        Gf = [1] * n
        Hf = [pow(y_inv, i, cv.r) for i in range(n)]
    want = O.ipa_create(cv, O.Transcript(b"innerproducttest"), Q, Gf, Hf, bp.G(n), bp.H(n), a, b)
    try:
        L, Rv, ao, bo = R.ipa_create(ctx, R.Transcript(b"innerproducttest"), Q, Gf, Hf, bp.G(n), bp.H(n), a, b)
    finally:
        ctx.set_ipa_nofold_threshold(1 << 13)      # library default
    assert (L, Rv, ao, bo) == (want.L_vec, want.R_vec, want.a, want.b)
    # and the proof verifies (make_ipp_*): P = <a,G> + <b*Hf,H> + <a,b>Q with Gf applied
    c = O.inner_product(cv, a, b)
    P = O.msm(cv, bp.G(n) + bp.H(n) + [Q], [x * g % cv.r for x, g in zip(a, Gf)] + [x * h % cv.r for x, h in zip(b, Hf)] + [c])
    O.ipa_verify(cv, O.InnerProductProof(L, Rv, ao, bo), n, O.Transcript(b"innerproducttest"), Gf, Hf, P, Q, bp.G(n), bp.H(n))
    # InnerProductProof::verify on the GPU (row a5): accepts the proof, rejects a wrong P and a tampered a
    assert R.ipa_verify(ctx, R.Transcript(b"innerproducttest"), n, L, Rv, ao, bo, Gf, Hf, P, Q, bp.G(n), bp.H(n))
    assert not R.ipa_verify(ctx, R.Transcript(b"innerproducttest"), n, L, Rv, ao, bo, Gf, Hf, O.pt_add(cv, P, Q), Q, bp.G(n), bp.H(n))
    assert not R.ipa_verify(ctx, R.Transcript(b"innerproducttest"), n, L, Rv, (ao + 1) % cv.r, bo, Gf, Hf, P, Q, bp.G(n), bp.H(n))


@pytest.mark.parametrize("name", [c[0] for c in C.GOLDEN_CASES])
def test_golden_proofs_byte_identical(env, name):
    from ark_bulletproofs_b200 import r1cs as R
    g = GOLDEN[name]
    curve, kind, params = g["curve"], g["kind"], g["params"]
    cv = O.CURVES[curve]
    ctx, gens = env(curve, max(g["gens_capacity"], 1))
    proof, coms = gpu_prove_case(R, ctx, gens, kind, params, curve)
    b = proof.to_bytes()
    assert [O.ser_point(cv, V, True).hex() for V in coms] == g["commitments_hex"]
    assert hashlib.sha256(b).hexdigest() == g["sha256"], "GPU proof bytes differ from the oracle's"
    assert b.hex() == g["proof_hex"]
    # the GPU verifier accepts it, also after a to_bytes/from_bytes round trip (tests/r1cs_secq256k1.rs:335-356)
    gpu_verifier(R, ctx, kind, params, curve, coms).verify(proof, gens)
    gpu_verifier(R, ctx, kind, params, curve, coms).verify(R.Proof.from_bytes(curve, b), gens)
    # and the oracle's verifier accepts the GPU proof (small cases; pure Python)
    if g["gens_capacity"] <= 128:
        pc, bp = O.PedersenGens(cv), O.BulletproofGens(cv, max(g["gens_capacity"], 1), 1)
        oracle_verifier(kind, params, curve, coms).verify(O.R1CSProof.from_bytes(cv, b), pc, bp)


def test_appendix_b_hashes(env):
    assert GOLDEN["v1_example"]["sha256"] == "765eaf6d94e0cd313db2691a02e39583483aa765bea4250d1b8032be830a83d5"
    assert GOLDEN["v2_shuffle3"]["sha256"] == "a51ae53f5e0fac5dc5bb9d5f520615763e3e845b66b6c4179645b8c01ee32151"
    assert GOLDEN["v3_range8"]["sha256"] == "d54cb874f445db631163dd4ce11056658f7bfdea5f9f88783af266c73bee658e"


def test_reject_matrix(env):
    """Wrong statements and tampered proofs are rejected exactly where the oracle rejects."""
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    cv = O.SECQ256K1
    ctx, gens = env(curve, 128)
    # example gadget with the wrong constant (tests/r1cs_secq256k1.rs:347)
    proof, coms = gpu_prove_case(R, ctx, gens, "example", {}, curve)
    with pytest.raises(R.BpError) as e:
        gpu_verifier(R, ctx, "example", {"c2": 10}, curve, coms).verify(proof, gens)
    assert e.value.code == -7
    # 4-bit range proof of 16 (tests/r1cs_secq256k1.rs:409)
    proof, coms = gpu_prove_case(R, ctx, gens, "range", {"value": 16, "bits": 4}, curve)
    with pytest.raises(R.BpError):
        gpu_verifier(R, ctx, "range", {"bits": 4}, curve, coms).verify(proof, gens)
    # tampering with every kind of field of a 2-phase proof
    kind, params = "shuffle", {"k": 4, "seed": 4}
    proof, coms = gpu_prove_case(R, ctx, gens, kind, params, curve)
    gpu_verifier(R, ctx, kind, params, curve, coms).verify(proof, gens)
    pc, bp = O.PedersenGens(cv), O.BulletproofGens(cv, 8, 1)
    for which in (0, 1, 2, 3, 4):
        bad = proof.clone()
        bad.set_scalar(which, (bad.get_scalar(which) + 1) % cv.r)
        with pytest.raises(R.BpError) as e:
            gpu_verifier(R, ctx, kind, params, curve, coms).verify(bad, gens)
        assert e.value.code == -7
        with pytest.raises(O.R1CSError):
            oracle_verifier(kind, params, curve, coms).verify(O.R1CSProof.from_bytes(cv, bad.to_bytes()), pc, bp)
    other = O.pt_mul(cv, 999, cv.G)
    for which in (10, 13, 16, 100, 201):
        bad = proof.clone()
        bad.set_point(which, other)
        with pytest.raises(R.BpError):
            gpu_verifier(R, ctx, kind, params, curve, coms).verify(bad, gens)
    for which in (10, 16, 100):           # identity -> validate_and_append_point error (transcript.rs:81-93)
        bad = proof.clone()
        bad.set_point(which, None)
        with pytest.raises(R.BpError):
            gpu_verifier(R, ctx, kind, params, curve, coms).verify(bad, gens)
    # insufficient generators -> InvalidGeneratorsLength (prover.rs:577-579, verifier.rs:425-427)
    _, small = env(curve, 4)
    with pytest.raises(R.BpError) as e:
        gpu_verifier(R, ctx, kind, params, curve, coms).verify(proof, small)
    assert e.value.code == -4
    with pytest.raises(R.BpError) as e:
        gpu_prove_case(R, ctx, small, kind, params, curve)
    assert e.value.code == -4


def test_batch_verify(env):
    """tests/r1cs_secq256k1.rs:447-475: mixed sizes; any invalid member rejects the batch."""
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    cv = O.SECQ256K1
    ctx, gens = env(curve, 128)

    def make(vals):
        inst = []
        for v, n in vals:
            proof, coms = gpu_prove_case(R, ctx, gens, "range", {"value": v, "bits": n}, curve)
            inst.append((gpu_verifier(R, ctx, "range", {"bits": n}, curve, coms), proof, coms, n))
        return inst
    good = [(0, 16), (3, 16), ((1 << 16) - 1, 16), (1 << 16, 32), (1 << 63, 64)]
    inst = make(good)
    R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), [(v, p) for v, p, _, _ in inst], gens)
    # same decision as the oracle's batch_verify on the GPU-made proofs
    pc, bp = O.PedersenGens(cv), O.BulletproofGens(cv, 64, 1)
    oinst = [(C.range_verifier(cv, coms[0], n), O.R1CSProof.from_bytes(cv, p.to_bytes())) for _, p, coms, n in inst]
    O.batch_verify(cv, O.ChaCha20Rng(bytes([5] * 32)), oinst, pc, bp)
    bad = [(0, 16), (3, 16), (1 << 16, 16), (1 << 16, 32)]
    inst = make(bad)
    with pytest.raises(R.BpError) as e:
        R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), [(v, p) for v, p, _, _ in inst], gens)
    assert e.value.code == -7
    R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), [], gens)     # empty batch: MSM of zero scalars -> accept


@pytest.mark.parametrize("k", [1, 2, 3, 5, 6])
def test_kshuffle_sizes(env, k):
    """tests/r1cs_secq256k1.rs:172-215 (k = 4,7,24,42 are golden cases above)."""
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    cv = O.SECQ256K1
    cap = 1 << (2 * k - 1).bit_length()
    ctx, gens = env(curve, cap)
    params = {"k": k, "seed": 100 + k}
    proof, coms = gpu_prove_case(R, ctx, gens, "shuffle", params, curve)
    gpu_verifier(R, ctx, "shuffle", params, curve, coms).verify(proof, gens)
    pc, bp = O.PedersenGens(cv), O.BulletproofGens(cv, cap, 1)
    inp, out = C.shuffle_values(k, 100 + k)
    want, _, _ = C.prove_shuffle(cv, pc, bp, inp, out)
    assert proof.to_bytes() == want.to_bytes(cv)


def test_batch_verify_sharded(env):
    """batch verification sharded over ranks -- each rank's partial MSM point; the sum over
    ranks is the identity iff the batch verifies. Emulated here with two shards on one GPU."""
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    cv = O.SECQ256K1
    ctx, gens = env(curve, 128)
    vals = [(0, 16), (3, 16), ((1 << 16) - 1, 16), (1 << 16, 32), (1 << 63, 64), (77, 8)]

    def run(vals):
        proofs = [gpu_prove_case(R, ctx, gens, "range", {"value": v, "bits": n}, curve) for v, n in vals]
        rng = R.ChaChaRng(bytes([5] * 32))
        alphas = [rng.scalar(curve) for _ in vals]             # every rank draws the same sequence
        parts = []
        for rank in range(2):
            idx = [i for i in range(len(vals)) if i % 2 == rank]
            inst = [(gpu_verifier(R, ctx, "range", {"bits": vals[i][1]}, curve, proofs[i][1]), proofs[i][0]) for i in idx]
            parts.append(R.batch_verify_partial(ctx, [alphas[i] for i in idx], inst, gens))
        return parts
    parts = run(vals)
    assert O.pt_add(cv, parts[0], parts[1]) is None            # (every valid proof's share is itself the identity)
    bad = list(vals)
    bad[2] = (1 << 16, 16)                                     # proof 2 lives on rank 0
    parts = run(bad)
    assert parts[0] is not None and parts[1] is None
    assert O.pt_add(cv, parts[0], parts[1]) is not None


@pytest.mark.parametrize("geo", [1, 0])
@pytest.mark.parametrize("nofold", [0, 16, 256])
@pytest.mark.parametrize("name", ["chain1000", "shuffle42", "shuffle24", "v3_range8", "zorro_chain20", "zorro_shuffle3", "c25519_chain20",
                                  "c25519_range8"])
def test_golden_proofs_all_ipa_paths(env, name, nofold, geo):
    """The same golden proofs with the generator-folding IPA (threshold 0) and mixed fold / no-fold, with and
    without the uniform-scalar fold for geometric factor vectors (shuffles and the 8-bit range proof qualify)."""
    from ark_bulletproofs_b200 import r1cs as R
    g = GOLDEN[name]
    curve, kind, params = g["curve"], g["kind"], g["params"]
    ctx, gens = env(curve, max(g["gens_capacity"], 1))
    ctx.set_ipa_nofold_threshold(nofold)
    ctx.set_ipa_geometric(bool(geo))
    try:
        proof, _ = gpu_prove_case(R, ctx, gens, kind, params, curve)
    finally:
        ctx.set_ipa_nofold_threshold(1 << 13)      # library default
        ctx.set_ipa_geometric(True)
    assert proof.to_bytes().hex() == g["proof_hex"]


@pytest.mark.parametrize("N,nofold", [(64, 0), (64, 8), (1 << 15, 1 << 14), (1 << 12, 1 << 10)])
def test_pow2_chain_geometric_equals_general(env, N, nofold):
    """One-phase circuit with a power-of-two multiplier count (no padding): the IPA factor vectors are
    (1, y^-i), so every fold round takes the uniform-scalar path. The proof must be byte-identical to the one
    from the general first round (itself pinned to the oracle by the golden cases), and verify."""
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    ctx, gens = env(curve, N)
    out = []
    ctx.set_ipa_nofold_threshold(nofold)
    try:
        for geo in (True, False):
            ctx.set_ipa_geometric(geo)
            proof, coms = gpu_prove_case(R, ctx, gens, "chain", {"N": N}, curve)
            out.append(proof.to_bytes())
    finally:
        ctx.set_ipa_nofold_threshold(1 << 13)      # library default
        ctx.set_ipa_geometric(True)
    assert out[0] == out[1]
    gpu_verifier(R, ctx, "chain", {"N": N}, curve, coms).verify(R.Proof.from_bytes(curve, out[0]), gens)


@pytest.mark.parametrize("curve", ["secq256k1", "curve25519"])
def test_commit_batch_equals_commit(env, curve):
    """bp_prover_commit_batch (GPU kernel) == m Prover::commit calls: same V_i, same transcript, same proof."""
    from ark_bulletproofs_b200 import r1cs as R
    cv = O.CURVES[curve]
    ctx, gens = env(curve, 16)
    inp, out = C.shuffle_values(5, 55)

    def transcript():
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", 5)
        return t
    rng = R.ChaChaRng(bytes(range(32)))
    p1 = R.Prover(ctx, gens, transcript())
    blinds = [rng.scalar(curve) for _ in range(10)]
    Vs, vars_ = p1.commit_batch(inp + out, blinds)
    pc = O.PedersenGens(cv)
    assert Vs == [pc.commit(v, b) for v, b in zip(inp + out, blinds)]
    R.shuffle_gadget(p1, vars_[:5], vars_[5:])
    proof1 = p1.prove(rng)
    rng2 = R.ChaChaRng(bytes(range(32)))
    p2 = R.Prover(ctx, gens, transcript())
    cs = [p2.commit(v, rng2.scalar(curve)) for v in inp + out]
    R.shuffle_gadget(p2, [v for _, v in cs[:5]], [v for _, v in cs[5:]])
    assert proof1.to_bytes() == p2.prove(rng2).to_bytes()
    # edge scalars
    p3 = R.Prover(ctx, gens, transcript())
    ev = [0, 1, cv.r - 1, 0]
    eb = [0, cv.r - 1, 1, 5]
    Vs, _ = p3.commit_batch(ev, eb)
    assert Vs == [pc.commit(v, b) for v, b in zip(ev, eb)]
    # the fixed-base table (default) and the double-and-add kernel agree, also on random and byte-boundary scalars
    rnd = random.Random(9)
    vals = [rnd.randrange(cv.r) for _ in range(20)] + [255, 256, (1 << 64) - 1, 1 << 248, cv.r - 2]
    bls = [rnd.randrange(cv.r) for _ in range(24)] + [0]
    want = [pc.commit(v, b) for v, b in zip(vals[:6], bls[:6])]
    for table in (True, False):
        ctx.set_pedersen_table(table)
        try:
            got, _ = R.Prover(ctx, gens, transcript()).commit_batch(vals, bls)
        finally:
            ctx.set_pedersen_table(True)
        assert got[:6] == want
        if table:
            ref = got
        else:
            assert got == ref


# ---- multi-GPU mode on one GPU: `world` contexts in threads, cyclic generator shards ---------------
def _sharded(world, curve, fn):
    """fn(R, ctx, rank) on `world` threads, each with its own sharded bp_ctx on cuda:0."""
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    from ark_bulletproofs_b200.dist import ThreadGroup

    def work(rank, allgather):
        ctx = Context(curve, 0)
        ctx.set_collective(rank, world, allgather)
        return fn(R, ctx, rank)
    return ThreadGroup(world).run(work)


@pytest.mark.parametrize("world,n,nofold", [(2, 64, 0), (2, 64, 8), (2, 8, 1 << 14), (2, 4, 0), (4, 64, 0), (4, 64, 8), (4, 8, 1 << 14), (4, 4, 0),
                                            (8, 64, 0), (8, 16, 4), (8, 8, 0)])
def test_sharded_ipa_matches_oracle(world, n, nofold):
    """InnerProductProof::create with G, H sharded cyclically over `world` contexts: the same L, R, a, b as the
    oracle on every rank, through local folds (n >= 2*world), the no-fold tail and arbitrary factor vectors."""
    curve = "secq256k1"
    cv = O.SECQ256K1
    rnd = random.Random(n * 31 + world)
    bp = O.BulletproofGens(cv, n, 1)
    Q = O.pt_mul(cv, 777, cv.G)
    a = [rnd.randrange(cv.r) for _ in range(n)]
    b = [rnd.randrange(cv.r) for _ in range(n)]
    Gf = [rnd.randrange(1, cv.r) for _ in range(n)]
    Hf = [rnd.randrange(1, cv.r) for _ in range(n)]
    want = O.ipa_create(cv, O.Transcript(b"innerproducttest"), Q, Gf, Hf, bp.G(n), bp.H(n), a, b)

    def run(R, ctx, rank):
        ctx.set_ipa_nofold_threshold(nofold)
        return R.ipa_create(ctx, R.Transcript(b"innerproducttest"), Q, Gf, Hf, bp.G(n), bp.H(n), a, b)
    for got in _sharded(world, curve, run):
        assert got == (want.L_vec, want.R_vec, want.a, want.b)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name,nofold", [("chain100", 1 << 14), ("chain100", 0), ("shuffle42", 16), ("chain1000", 64), ("range63", 0),
                                         ("zorro_chain20", 0), ("c25519_chain20", 8)])
def test_sharded_prove_verify_golden(world, name, nofold):
    """Prover and verifier with the generators sharded over `world` contexts: every rank emits the golden proof
    bytes (commitment MSMs, IPA L/R and folds run on the shards, partial points all-gathered), every rank's
    verifier accepts them and rejects a tampered copy."""
    g = GOLDEN[name]
    curve, kind, params = g["curve"], g["kind"], g["params"]

    def run(R, ctx, rank):
        gens = R.Gens(ctx, max(g["gens_capacity"], 1))
        ctx.set_ipa_nofold_threshold(nofold)
        proof, coms = gpu_prove_case(R, ctx, gens, kind, params, curve)
        raw = proof.to_bytes()
        gpu_verifier(R, ctx, kind, params, curve, coms).verify(R.Proof.from_bytes(curve, raw), gens)
        bad = bytearray(raw)
        bad[-32] ^= 1                                              # IPA scalar b (low byte: still canonical)
        try:
            gpu_verifier(R, ctx, kind, params, curve, coms).verify(R.Proof.from_bytes(curve, bytes(bad)), gens)
            rejected = False
        except Exception:
            rejected = True
        return raw.hex(), rejected, ctx.launches
    for hexs, rejected, launches in _sharded(world, curve, run):
        assert hexs == g["proof_hex"]
        assert rejected
        assert launches > 0


@pytest.mark.parametrize("name", ["shuffle7", "shuffle42", "zorro_shuffle3", "c25519_shuffle3"])
def test_native_shuffle_gadget_golden(env, name):
    """bp_cs_shuffle_gadget (the reference's bench/test gadget built inside the library) with batched commitments
    gives the same golden proof bytes as the gadget driven call by call, and its verifier side accepts them."""
    from ark_bulletproofs_b200 import codec
    from ark_bulletproofs_b200 import r1cs as R
    g = GOLDEN[name]
    curve, kind, params = g["curve"], g["kind"], g["params"]
    ctx, gens = env(curve, max(g["gens_capacity"], 1))
    inp, out = (params["inp"], params["out"]) if kind == "shuffle_fixed" else C.shuffle_values(params["k"], params["seed"])
    k = len(inp)

    def transcript():
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", k)
        return t
    rng = seed_a()
    blinds_raw = rng.scalars_raw(curve, 2 * k)
    p = R.Prover(ctx, gens, transcript())
    coms_raw, vars_ = p.commit_batch_raw(codec.enc_scalars(inp + out, curve), blinds_raw, 2 * k)
    p.shuffle_gadget_native(vars_[:k], vars_[k:])
    proof = p.prove(rng)
    assert proof.to_bytes().hex() == g["proof_hex"]
    v = R.Verifier(ctx, transcript())
    vv = v.commit_batch_raw(coms_raw, 2 * k)
    v.shuffle_gadget_native(vv[:k], vv[k:])
    v.verify(proof, gens)


@pytest.mark.parametrize("cap", [256, 1000, 5000])
def test_device_generators_match_host(cap):
    """BulletproofGens chains generated on the GPU (gens_kernels.cuh: seekable ChaCha, Tonelli-Shanks, stream-order
    compaction) are the points of the host generator, which tests/test_host_layer.py pins to the oracle."""
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    ctx = Context(curve, 0)
    g = R.Gens(ctx, cap)                       # device path (default)
    _, _, G, H = R.generate_gens_host(curve, cap)
    assert g.export(0, 0, cap) == G
    assert g.export(1, 0, cap) == H
    ctx.set_device_gens(False)
    g2 = R.Gens(ctx, cap)                      # host path, uploaded
    assert g2.export(0, 0, cap) == G and g2.export(1, 0, cap) == H


def test_device_generators_sharded():
    """Sharded contexts keep generator i on rank i mod world, whichever side generated the chain."""
    from ark_bulletproofs_b200 import r1cs as R
    curve, cap, world = "secq256k1", 600, 4
    _, _, G, H = R.generate_gens_host(curve, cap)

    def run(R_, ctx, rank):
        g = R_.Gens(ctx, cap)
        n_loc = len(range(rank, cap, world))
        return g.export(0, 0, n_loc), g.export(1, 0, n_loc)
    for rank, (gl, hl) in enumerate(_sharded(world, curve, run)):
        assert gl == G[rank::world] and hl == H[rank::world]


@pytest.mark.parametrize("name", ["chain1000", "shuffle42"])
def test_golden_proofs_without_glv(env, name):
    """The generator fold defaults to the GLV split on secq256k1 (every other test); the plain 256-step fold must
    give the same golden bytes."""
    from ark_bulletproofs_b200 import r1cs as R
    g = GOLDEN[name]
    curve, kind, params = g["curve"], g["kind"], g["params"]
    ctx, gens = env(curve, max(g["gens_capacity"], 1))
    ctx.set_ipa_nofold_threshold(0)
    ctx.set_ipa_glv(False)
    try:
        proof, _ = gpu_prove_case(R, ctx, gens, kind, params, curve)
    finally:
        ctx.set_ipa_nofold_threshold(1 << 13)      # library default
        ctx.set_ipa_glv(True)
    assert proof.to_bytes().hex() == g["proof_hex"]


@pytest.mark.parametrize("name", ["v1_example", "shuffle7", "chain100", "zorro_shuffle3", "c25519_range8"])
def test_golden_proofs_without_tiny_msm(env, name):
    """Small proofs run their MSMs through the single-launch kernel by default; the bucket pipeline must give the same
    golden bytes (it is the only path for anything larger)."""
    from ark_bulletproofs_b200 import r1cs as R
    g = GOLDEN[name]
    curve, kind, params = g["curve"], g["kind"], g["params"]
    ctx, gens = env(curve, max(g["gens_capacity"], 1))
    ctx.set_tiny(0)
    try:
        proof, coms = gpu_prove_case(R, ctx, gens, kind, params, curve)
        gpu_verifier(R, ctx, kind, params, curve, coms).verify(proof, gens)
    finally:
        ctx.set_tiny(768)
    assert proof.to_bytes().hex() == g["proof_hex"]


@pytest.mark.parametrize("curve", ["secq256k1", "zorro", "curve25519"])
def test_proofs_from_bytes_batch(curve):
    """Batch ingest with GPU point decompression: the same proofs (re-serialised byte for byte) and the same
    accept / reject decisions as bp_proof_from_bytes -- golden proofs plus copies with a non-residue x, a flipped sign
    bit (still a valid encoding: accepted, different point), bad flag bits, a non-canonical x, a truncated blob, a
    non-canonical scalar."""
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    ctx = Context(curve, 0)
    names = [n for n, g in GOLDEN.items() if g["curve"] == curve]
    blobs = [bytes.fromhex(GOLDEN[n]["proof_hex"]) for n in names]
    pc = 32 if curve == "curve25519" else 33
    base = blobs[0]
    bad = []
    for delta in range(1, 40):                         # some x + delta is not on the curve
        b = bytearray(base)
        b[0] = (b[0] + delta) & 0xFF
        bad.append(bytes(b))
    sign = bytearray(base); sign[pc - 1] ^= 0x80; bad.append(bytes(sign))
    if pc == 33:
        fl = bytearray(base); fl[32] |= 0x01; bad.append(bytes(fl))            # padding bit: ark-ff never reads it -> accepted
        inf = bytearray(base); inf[32] = 0xC0; bad.append(bytes(inf))           # infinity + sign: UnexpectedFlags
        inx = bytearray(base); inx[pc * 3 + 32] = 0x40; bad.append(bytes(inx))  # A_I2 := infinity flag with x != 0: the identity for ark-ec
        big = bytearray(base); big[0:32] = b"\xff" * 32; bad.append(bytes(big))  # x >= q
    bad.append(base[:-5])
    sc = bytearray(base); sc[-32:] = b"\xff" * 32; bad.append(bytes(sc))        # scalar >= r
    allb = blobs + bad
    got = R.Proof.from_bytes_batch(ctx, allb)
    n_rej = 0
    for blob, pr in zip(allb, got):
        try:
            want = R.Proof.from_bytes(curve, blob)
        except Exception:
            want = None
        assert (pr is None) == (want is None)
        if pr is not None:
            assert pr.to_bytes() == want.to_bytes()
            malleated = pc == 33 and any(blob[33 * k + 32] & 0x3F or (blob[33 * k + 32] & 0x40 and any(blob[33 * k:33 * k + 32])) for k in range(11))
            assert (pr.to_bytes() == blob) != malleated          # re-serialisation is canonical
            cvo = O.CURVES[curve]
            assert O.R1CSProof.from_bytes(cvo, blob).to_bytes(cvo) == pr.to_bytes()
        else:
            n_rej += 1
    assert n_rej >= 3 and all(p is not None for p in got[:len(blobs)])
    assert R.Proof.from_bytes_batch(ctx, []) == []


# ---- ABI hardening: foreign variables, unchecked points, mixed contexts ----------------------------------------
def test_foreign_variables_rejected(env):
    """A Variable that does not belong to the constraint system (stale, forged kind, out-of-range index) is BP_ERR_ARG at
    multiply / constrain instead of an out-of-bounds index in flatten (the Rust reference panics on the index)."""
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    ctx, gens = env(curve, 8)
    for make in (lambda: R.Prover(ctx, gens, R.Transcript(b"t")), lambda: R.Verifier(ctx, R.Transcript(b"t"))):
        cs = make()
        if isinstance(cs, R.Prover):
            _, v0 = cs.commit(5, 7)
            l, r_, o = cs.allocate_multiplier((3, 4))
        else:
            v0 = cs.commit(O.pt_mul(O.SECQ256K1, 9, O.SECQ256K1.G))
            l, r_, o = cs.allocate_multiplier(None)
        cs.constrain(R.LC.of(l) + v0 - o)                                        # fine
        for bad in (R.Variable(1, 1), R.Variable(2, 1 << 20), R.Variable(3, (1 << 29) + 0), R.Variable(0, 1), R.Variable(5, 0), R.Variable(9, 0)):
            with pytest.raises(R.BpError) as e:
                cs.constrain(R.LC.of(bad) - l)
            assert e.value.code == -1
            with pytest.raises(R.BpError) as e:
                cs.multiply(R.LC.of(bad), R.LC.of(l))
            assert e.value.code == -1


@pytest.mark.parametrize("curve", ["secq256k1", "zorro", "curve25519"])
def test_invalid_points_rejected_at_the_boundary(env, curve):
    """Statement commitments, caller-supplied generators and bp_proof_set_field take raw 64-byte points: off-curve,
    non-canonical and (curve25519, cofactor 8) small-order points are BP_ERR_FORMAT, as ark-serialize's validation would
    have rejected them before the reference ever held them as `G`."""
    import ctypes as ct
    from ark_bulletproofs_b200 import codec
    from ark_bulletproofs_b200 import r1cs as R
    cv = O.CURVES[curve]
    ctx, gens = env(curve, 8)
    q = cv.q
    good = O.pt_mul(cv, 12345, cv.G)
    off_curve = codec.enc_fe(good[0], q) + codec.enc_fe((good[1] + 1) % q, q)
    # a Montgomery residue >= q (not canonical): x + q still fits 256 bits for all three base fields
    raw_x = int.from_bytes(codec.enc_fe(good[0], q), "little")
    bads = [off_curve]
    if raw_x + q < 1 << 256:
        bads.append((raw_x + q).to_bytes(32, "little") + codec.enc_fe(good[1], q))
    if curve == "curve25519":
        bads.append(codec.enc_fe(0, q) + codec.enc_fe(q - 1, q))                  # (0, -1): order 2
        # an order-8 point times nothing: good + (0,-1) is on the curve but outside the prime-order subgroup
        bads.append(codec.enc_point(O.pt_add(cv, good, (0, q - 1)), curve))
    v = R.Verifier(ctx, R.Transcript(b"t"))
    for b in bads:
        var = R.BpVar()
        assert v.lib.bp_verifier_commit(v.h, b, ct.byref(var)) == -8
    v.commit(good)
    v.commit(None)                                                               # the identity is a valid group element
    # proof fields
    proof, _ = gpu_prove_case(R, ctx, gens, "shuffle_fixed", {"inp": [5, 9, 2], "out": [2, 5, 9]}, curve)
    for b in bads:
        assert proof.lib.bp_proof_set_field(proof.h, 10, b) == -8
    # generators
    G = [O.pt_mul(cv, 100 + i, cv.G) for i in range(4)]
    H = [O.pt_mul(cv, 200 + i, cv.G) for i in range(4)]
    pc = O.PedersenGens(cv)
    lib = ctx.lib
    h = ct.c_void_p()
    ok_args = [codec.enc_point(pc.B, curve), codec.enc_point(pc.B_blinding, curve), codec.enc_points(G, curve), codec.enc_points(H, curve)]
    assert lib.bp_gens_from_points(ctx.h, *ok_args, 4, ct.byref(h)) == 0
    lib.bp_gens_free(h)
    for slot in range(4):
        for b in bads:
            args = list(ok_args)
            args[slot] = b + args[slot][64:] if slot >= 2 else b
            assert lib.bp_gens_from_points(ctx.h, *args, 4, ct.byref(h)) == -8


def test_batch_verify_rejects_foreign_contexts(env):
    """bp_batch_verify / bp_verifier_verify with a verifier or generators of another context: BP_ERR_ARG (their buffers live
    on that context's stream)."""
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    curve = "secq256k1"
    ctx, gens = env(curve, 8)
    other = Context(curve, 0)
    gens2 = R.Gens(other, 8)
    kind, params = "shuffle_fixed", {"inp": [5, 9, 2], "out": [2, 5, 9]}
    proof, coms = gpu_prove_case(R, ctx, gens, kind, params, curve)
    v_other = gpu_verifier(R, other, kind, params, curve, coms)
    with pytest.raises(R.BpError) as e:
        R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), [(v_other, proof)], gens)
    assert e.value.code == -1
    with pytest.raises(R.BpError) as e:
        R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), [(gpu_verifier(R, ctx, kind, params, curve, coms), proof)], gens2)
    assert e.value.code == -1
    with pytest.raises(R.BpError) as e:
        gpu_verifier(R, ctx, kind, params, curve, coms).verify(proof, gens2)
    assert e.value.code == -1
    R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), [(gpu_verifier(R, ctx, kind, params, curve, coms), proof)], gens)
    gpu_verifier(R, other, kind, params, curve, coms).verify(proof, gens2)


# ---- device-side transcript (csrc/transcript_dev.cuh) ------------------------------------------------
@pytest.mark.parametrize("curve", ["secq256k1", "zorro", "curve25519"])
def test_device_transcript_challenges_match_oracle(curve):
    """Keccak-f / STROBE / Merlin / ChaCha20 / ScalarField::rand on the GPU: every inner-product challenge u_j, its
    inverse, and the cloned-transcript challenge r equal the oracle's, for transcripts in different STROBE positions,
    1..20 rounds (each round crosses the 166-byte rate at a different offset) and all three curves (curve25519's scalar
    field rejects half of the ChaCha draws; its points serialise without a flag byte)."""
    import ctypes as ct
    from ark_bulletproofs_b200 import Context, codec
    from ark_bulletproofs_b200 import r1cs as R
    cv = O.CURVES[curve]
    ctx = Context(curve, 0)
    rnd = random.Random(4242)
    r = cv.r
    for lg_n, prefix in [(1, 0), (2, 3), (5, 17), (8, 40), (16, 101), (20, 7)]:
        pts = [O.pt_mul(cv, rnd.randrange(1, r), cv.G) for _ in range(2 * lg_n)]
        L, Rp = pts[:lg_n], pts[lg_n:]
        n = 1 << lg_n
        # same prefix on both sides, so the IPA part starts at different byte positions of the sponge
        ot, gt = O.Transcript(b"devtr"), R.Transcript(b"devtr")
        msg = bytes(range(prefix))
        ot.append_message(b"pad", msg)
        gt.append_message(b"pad", msg)
        assert O.challenge_scalar(cv, ot, b"w") == gt.challenge_scalar(curve, b"w")
        # oracle: inner_product_proof.rs:266-277, then verifier.rs:516-519
        ot.append_message(b"dom-sep", b"ipp v1")
        ot.append_u64(b"n", n)
        want_u = []
        for j in range(lg_n):
            O.validate_and_append_point(cv, ot, b"L", L[j])
            O.validate_and_append_point(cv, ot, b"R", Rp[j])
            want_u.append(O.challenge_scalar(cv, ot, b"u"))
        want_r = O.challenge_scalar(cv, ot.clone(), b"r")
        out_u, out_ui, out_r = ct.create_string_buffer(32 * lg_n), ct.create_string_buffer(32 * lg_n), ct.create_string_buffer(32)
        bad = ct.c_int(0)
        rc = ctx.lib.bp_transcript_ipa_challenges_device(ctx.h, gt.h, n, codec.enc_points(L, curve), codec.enc_points(Rp, curve), lg_n,
                                                         out_u, out_ui, out_r, ct.byref(bad))
        assert rc == 0 and bad.value == 0
        got_u = [codec.dec_fe(out_u.raw[32 * j:32 * j + 32], r) for j in range(lg_n)]
        got_ui = [codec.dec_fe(out_ui.raw[32 * j:32 * j + 32], r) for j in range(lg_n)]
        assert got_u == want_u
        assert got_ui == [pow(u, -1, r) for u in want_u]
        assert codec.dec_fe(out_r.raw, r) == want_r
        # the host transcript was not advanced: the next host challenge is what the oracle gets before the IPA part
        # an identity point is reported (validate_and_append_point -> VerificationError)
        Lbad = list(L)
        Lbad[lg_n // 2] = None
        rc = ctx.lib.bp_transcript_ipa_challenges_device(ctx.h, gt.h, n, codec.enc_points(Lbad, curve), codec.enc_points(Rp, curve), lg_n,
                                                         out_u, out_ui, out_r, ct.byref(bad))
        assert rc == 0 and bad.value == 1


@pytest.mark.parametrize("curve", ["secq256k1", "curve25519"])
def test_batch_verify_device_transcript(env, curve):
    """batch_verify with the IPA challenges of all proofs derived on the device (forced: threshold 1) makes the same
    decisions as with the host transcript: accepts the golden proofs, rejects a batch with one tampered proof, and
    reports an identity L_j as a verification error."""
    from ark_bulletproofs_b200 import r1cs as R
    ctx, gens = env(curve, 128)
    cases = [("shuffle_fixed", {"inp": [5, 9, 2], "out": [2, 5, 9]}), ("chain", {"N": 20}), ("range", {"value": 0xA5, "bits": 8}), ("example", {})]
    made = [(k, p) + gpu_prove_case(R, ctx, gens, k, p, curve) for k, p in cases]

    def instances(tamper=None):
        out = []
        for i, (k, p, proof, coms) in enumerate(made):
            pr = proof
            if tamper is not None and i == tamper[0]:
                pr = proof.clone()
                tamper[1](pr)
            out.append((gpu_verifier(R, ctx, k, p, curve, coms), pr))
        return out
    for thresh in (1, 0):
        ctx.set_device_transcript(thresh)
        try:
            R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), instances(), gens)
            with pytest.raises(R.BpError) as e:
                R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), instances((1, lambda pr: pr.set_scalar(3, (pr.get_scalar(3) + 1) % O.CURVES[curve].r))), gens)
            assert e.value.code == -7
            with pytest.raises(R.BpError) as e:
                R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), instances((2, lambda pr: pr.set_point(100, None))), gens)
            assert e.value.code == -7
            other = O.pt_mul(O.CURVES[curve], 999, O.CURVES[curve].G)
            with pytest.raises(R.BpError) as e:
                R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), instances((0, lambda pr: pr.set_point(200, other))), gens)
            assert e.value.code == -7
        finally:
            ctx.set_device_transcript(32)

"""Shared SEED-A proving scenarios used by the oracle
tests, the golden-fixture generator and the GPU parity tests."""
import bp_oracle as O


def seed_a_rng():
    return O.ChaCha20Rng(bytes(range(32)))


def prove_example(cv, pc, bp, c2=9, rng=None):
    rng = rng or seed_a_rng()
    p = O.Prover(cv, pc, O.Transcript(b"R1CSExampleGadget"))
    cvs = [p.commit(v, O.scalar_rand(cv, rng)) for v in (3, 4, 6, 1, 40)]
    O.example_gadget(p, *[v for _, v in cvs], c2)
    return p.prove(rng, bp), [V for V, _ in cvs]


def verify_example(cv, pc, bp, proof, commitments, c2=9):
    v = O.Verifier(cv, O.Transcript(b"R1CSExampleGadget"))
    vs = [v.commit(V) for V in commitments]
    O.example_gadget(v, *vs, c2)
    v.verify(proof, pc, bp)


def prove_shuffle(cv, pc, bp, inp, out, rng=None, trace=None):
    rng = rng or seed_a_rng()
    t = O.Transcript(b"ShuffleProofTest")
    t.append_message(b"dom-sep", b"ShuffleProof")
    t.append_u64(b"k", len(inp))
    p = O.Prover(cv, pc, t)
    ic = [p.commit(v, O.scalar_rand(cv, rng)) for v in inp]
    oc = [p.commit(v, O.scalar_rand(cv, rng)) for v in out]
    O.shuffle_gadget(p, [v for _, v in ic], [v for _, v in oc])
    return p.prove(rng, bp, trace), [V for V, _ in ic], [V for V, _ in oc]


def shuffle_verifier(cv, ic, oc):
    t = O.Transcript(b"ShuffleProofTest")
    t.append_message(b"dom-sep", b"ShuffleProof")
    t.append_u64(b"k", len(ic))
    v = O.Verifier(cv, t)
    iv = [v.commit(V) for V in ic]
    ov = [v.commit(V) for V in oc]
    O.shuffle_gadget(v, iv, ov)
    return v


def prove_range(cv, pc, bp, value, nbits, rng=None):
    rng = rng or seed_a_rng()
    p = O.Prover(cv, pc, O.Transcript(b"RangeProofTest"))
    com, var = p.commit(value, O.scalar_rand(cv, rng))
    O.range_proof_gadget(p, var, value, nbits)
    return p.prove(rng, bp), com


def range_verifier(cv, com, nbits):
    v = O.Verifier(cv, O.Transcript(b"RangeProofTest"))
    var = v.commit(com)
    O.range_proof_gadget(v, var, None, nbits)
    return v


# ---- deterministic case table shared by tests/golden/make_golden.py and the GPU parity tests ----
def shuffle_values(k, seed):
    import random
    rnd = random.Random(seed)
    inp = [rnd.randrange(1 << 64) for _ in range(k)]
    out = list(inp)
    rnd.shuffle(out)
    return inp, out


def prove_chain(cv, pc, bp, N, rng=None):
    rng = rng or seed_a_rng()
    x0, ks = O.chain_circuit_witness(cv, N)
    p = O.Prover(cv, pc, O.Transcript(b"ChainCircuit"))
    com, var = p.commit(x0, O.scalar_rand(cv, rng))
    O.chain_circuit(p, var, N, ks, x0, cv.r)
    return p.prove(rng, bp), com, ks


def chain_verifier(cv, com, N, ks):
    v = O.Verifier(cv, O.Transcript(b"ChainCircuit"))
    var = v.commit(com)
    O.chain_circuit(v, var, N, ks, None, cv.r)
    return v


GOLDEN_CASES = [
    # (name, curve, kind, params)
    ("v1_example", "secq256k1", "example", {}),
    ("v2_shuffle3", "secq256k1", "shuffle_fixed", {"inp": [5, 9, 2], "out": [2, 5, 9]}),
    ("v3_range8", "secq256k1", "range", {"value": 0xA5, "bits": 8}),
    ("shuffle4", "secq256k1", "shuffle", {"k": 4, "seed": 4}),
    ("shuffle7", "secq256k1", "shuffle", {"k": 7, "seed": 7}),
    ("shuffle24", "secq256k1", "shuffle", {"k": 24, "seed": 24}),
    ("shuffle42", "secq256k1", "shuffle", {"k": 42, "seed": 42}),
    ("range63", "secq256k1", "range", {"value": (1 << 63) - 12345, "bits": 63}),
    ("chain5", "secq256k1", "chain", {"N": 5}),
    ("chain100", "secq256k1", "chain", {"N": 100}),
    ("chain1000", "secq256k1", "chain", {"N": 1000}),
    ("zorro_example", "zorro", "example", {}),
    ("zorro_shuffle3", "zorro", "shuffle_fixed", {"inp": [5, 9, 2], "out": [2, 5, 9]}),
    ("zorro_chain20", "zorro", "chain", {"N": 20}),
    ("c25519_example", "curve25519", "example", {}),
    ("c25519_shuffle3", "curve25519", "shuffle_fixed", {"inp": [5, 9, 2], "out": [2, 5, 9]}),
    ("c25519_range8", "curve25519", "range", {"value": 0xA5, "bits": 8}),
    ("c25519_chain20", "curve25519", "chain", {"N": 20}),
]


def oracle_prove_case(kind, params, cv, pc, bp):
    """Returns (proof, [commitments])."""
    if kind == "example":
        proof, coms = prove_example(cv, pc, bp, 9)
        return proof, coms
    if kind in ("shuffle", "shuffle_fixed"):
        inp, out = (params["inp"], params["out"]) if kind == "shuffle_fixed" else shuffle_values(params["k"], params["seed"])
        proof, ic, oc = prove_shuffle(cv, pc, bp, inp, out)
        return proof, ic + oc
    if kind == "range":
        proof, com = prove_range(cv, pc, bp, params["value"], params["bits"])
        return proof, [com]
    if kind == "chain":
        proof, com, _ = prove_chain(cv, pc, bp, params["N"])
        return proof, [com]
    raise ValueError(kind)


def gens_capacity(kind, params):
    if kind == "example":
        return 1
    if kind == "shuffle_fixed":
        return 8
    if kind == "shuffle":
        return 1 << (2 * params["k"] - 1).bit_length()
    if kind == "range":
        return 1 << (params["bits"] - 1).bit_length()
    return 1 << (params["N"] - 1).bit_length()

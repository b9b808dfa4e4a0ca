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

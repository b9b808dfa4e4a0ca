"""The reference's own bench circuit (benches/r1cs_secq256k1.rs:35-75, k-shuffle) scaled to 2(k-1) = 2^lg
multipliers: two-phase, n1 = 0, m = 2k commitments. Prove + verify timing."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context  # noqa: E402
from ark_bulletproofs_b200 import r1cs as R  # noqa: E402

curve = "secq256k1"
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 12
k = (1 << (lg - 1)) + 1
ctx = Context(curve, 0)
ctx.set_timing(True)
t0 = time.perf_counter()
gens = R.Gens(ctx, 1 << lg)
t_gens = time.perf_counter() - t0
rnd = random.Random(k)
inp = [rnd.randrange(1 << 64) for _ in range(k)]
out = list(inp)
rnd.shuffle(out)


def transcript():
    t = R.Transcript(b"ShuffleProofTest")
    t.append_message(b"dom-sep", b"ShuffleProof")
    t.append_u64(b"k", k)
    return t


rng = R.ChaChaRng(bytes(range(32)))
p = R.Prover(ctx, gens, transcript())
blinds = [rng.scalar(curve) for _ in range(2 * k)]
t0 = time.perf_counter()
Vs, vars_ = p.commit_batch(inp + out, blinds)
t_commit = time.perf_counter() - t0
R.shuffle_gadget(p, vars_[:k], vars_[k:])
t0 = time.perf_counter()
proof = p.prove(rng)
t_prove = time.perf_counter() - t0
st_p = ctx.last_stage_ms()
v = R.Verifier(ctx, transcript())
t0 = time.perf_counter()
vv = [v.commit(V) for V in Vs]
t_vcommit = time.perf_counter() - t0
R.shuffle_gadget(v, vv[:k], vv[k:])
t0 = time.perf_counter()
v.verify(proof, gens)
t_verify = time.perf_counter() - t0
st_v = ctx.last_stage_ms()
row = {"circuit": "k-shuffle k=%d (2^%d multipliers, two-phase, m=%d)" % (k, lg, 2 * k), "gens_s": round(t_gens, 2),
       "commit_batch_ms": round(t_commit * 1e3, 1), "prove_ms": round(t_prove * 1e3, 1), "verify_ms": round(t_verify * 1e3, 1),
       "verifier_commit_ms": round(t_vcommit * 1e3, 1), "prove_stages": {a: b for a, b in st_p.items() if b},
       "verify_stages": {a: b for a, b in st_v.items() if b}, "proof_bytes": len(proof.to_bytes())}
print(json.dumps(row))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(row, open(os.path.join(ROOT, "gpurun_out", "shuffle_quick_%d.json" % lg), "w"), indent=1)

"""The reference's own criterion benches re-run through the C ABI (benches/r1cs_secq256k1.rs:156-250, r1cs_zorro.rs):
k-shuffle proof creation and verification for k = 8 ... 1024 (2(k-1) multipliers), plus scaled-up sizes, built with
the native gadget (bp_cs_shuffle_gadget). Timed like the reference does it: proving includes the 2k Pedersen
commitments of ShuffleProof::prove and the gadget; verification includes the 2k verifier.commit calls.
    python tools/shuffle_bench.py [curve] [k,k,...]"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context, codec  # noqa: E402
from ark_bulletproofs_b200 import r1cs as R  # noqa: E402

curve = sys.argv[1] if len(sys.argv) > 1 else "secq256k1"
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32, 64, 128, 256, 512, 1024, 32769]
ctx = Context(curve, 0)
ctx.set_timing(True)
r = codec.MODULI[curve][1]
rows = []
for k in ks:
    cap = 1
    while cap < 2 * (k - 1):
        cap <<= 1
    gens = R.Gens(ctx, cap)
    rnd = random.Random(k)
    inp = [rnd.randrange(1 << 64) for _ in range(k)]
    out = list(inp)
    rnd.shuffle(out)
    vals_raw = codec.enc_scalars(inp + out, curve)

    def transcript():
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", k)
        return t
    best_p = best_v = None
    for rep in range(3):
        rng = R.ChaChaRng(bytes(range(32)))
        blinds_raw = rng.scalars_raw(curve, 2 * k)
        t0 = time.perf_counter()
        p = R.Prover(ctx, gens, transcript())
        coms_raw, vars_ = p.commit_batch_raw(vals_raw, blinds_raw, 2 * k)
        p.shuffle_gadget_native(vars_[:k], vars_[k:])
        proof = p.prove(rng)
        t_p = (time.perf_counter() - t0) * 1e3
        st_p = ctx.last_stage_ms()
        t0 = time.perf_counter()
        v = R.Verifier(ctx, transcript())
        vv = v.commit_batch_raw(coms_raw, 2 * k)
        v.shuffle_gadget_native(vv[:k], vv[k:])
        v.verify(proof, gens)
        t_v = (time.perf_counter() - t0) * 1e3
        st_v = ctx.last_stage_ms()
        if best_p is None or t_p < best_p[0]:
            best_p = (t_p, st_p)
        if best_v is None or t_v < best_v[0]:
            best_v = (t_v, st_v)
    row = {"curve": curve, "k": k, "multipliers": 2 * (k - 1), "prove_ms": round(best_p[0], 2), "verify_ms": round(best_v[0], 2),
           "proof_bytes": len(proof.to_bytes()), "prove_stages": {a: b for a, b in best_p[1].items() if b},
           "verify_stages": {a: b for a, b in best_v[1].items() if b}}
    print(json.dumps(row), flush=True)
    rows.append(row)
    del gens
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "shuffle_bench_%s.json" % curve), "w"), indent=1)

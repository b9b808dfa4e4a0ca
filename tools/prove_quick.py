"""Quick prove/verify timing of the synthetic chain circuit (not the bench contract)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context, codec  # noqa: E402
from ark_bulletproofs_b200 import r1cs as R  # noqa: E402

# sizes: log2 of the multiplier count, or a literal count prefixed with n (e.g. n50000: a padded circuit)
lgs = [x for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else ["10", "12", "14", "16"]
timing = len(sys.argv) > 2 and sys.argv[2] == "timing"
curve = sys.argv[3] if len(sys.argv) > 3 else "secq256k1"
ctx = Context(curve, 0)
ctx.set_timing(timing)
if os.environ.get("BP_GLV"):
    ctx.set_ipa_glv(int(os.environ["BP_GLV"]))
if os.environ.get("BP_NOFOLD"):
    ctx.set_ipa_nofold_threshold(int(os.environ["BP_NOFOLD"]))
r = codec.MODULI[curve][1]
res = []
for lg in lgs:
    N = int(lg[1:]) if lg.startswith("n") else 1 << int(lg)
    cap = 1
    while cap < N:
        cap <<= 1
    t0 = time.perf_counter()
    gens = R.Gens(ctx, cap)
    t_gens = time.perf_counter() - t0
    wit = R.ChaChaRng(bytes([3] * 32))
    x0_raw = wit.scalars_raw(curve, 1)
    ks_raw = wit.scalars_raw(curve, N)
    for rep in range(2):
        rng = R.ChaChaRng(bytes(range(32)))
        p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
        blind = rng.scalar(curve)
        t0 = time.perf_counter()
        com, var = p.commit(codec.dec_fe(x0_raw, r), blind)
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        proof = p.prove(rng)
        t_prove = time.perf_counter() - t0
        st_p = ctx.last_stage_ms()
        v = R.Verifier(ctx, R.Transcript(b"ChainCircuit"))
        vv = v.commit(com)
        v.chain_circuit_raw(vv, N, ks_raw, None)
        t0 = time.perf_counter()
        v.verify(proof, gens)
        t_verify = time.perf_counter() - t0
        st_v = ctx.last_stage_ms()
    row = {"curve": curve, "lg_n": lg if lg.startswith("n") else int(lg), "gens_s": round(t_gens, 2), "build_ms": round(t_build * 1e3, 1), "prove_ms": round(t_prove * 1e3, 2),
           "verify_ms": round(t_verify * 1e3, 2), "prove_stages": {k: v for k, v in st_p.items() if v}, "verify_stages": {k: v for k, v in st_v.items() if v},
           "proof_bytes": len(proof.to_bytes())}
    print(json.dumps(row), flush=True)
    res.append(row)
    del gens
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "prove_quick_%s.json" % curve), "w"), indent=1)

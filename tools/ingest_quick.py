"""Proof ingest timing: n x bp_proof_from_bytes (host decompression) against one bp_proofs_from_bytes_batch (GPU)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context, codec  # noqa: E402
from ark_bulletproofs_b200 import r1cs as R  # noqa: E402

curve = "secq256k1"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = Context(curve, 0)
N = 1 << 16
gens = R.Gens(ctx, N)
wit = R.ChaChaRng(bytes([3] * 32))
x0_raw, ks_raw = wit.scalars_raw(curve, 1), wit.scalars_raw(curve, N)
rng = R.ChaChaRng(bytes(range(32)))
p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
com, var = p.commit(codec.dec_fe(x0_raw, codec.MODULI[curve][1]), rng.scalar(curve))
p.chain_circuit_raw(var, N, ks_raw, x0_raw)
blob = p.prove(rng).to_bytes()
blobs = [blob] * n
R.Proof.from_bytes_batch(ctx, blobs[:4])
t0 = time.perf_counter()
a = [R.Proof.from_bytes(curve, b) for b in blobs]
t_host = time.perf_counter() - t0
t0 = time.perf_counter()
b = R.Proof.from_bytes_batch(ctx, blobs)
t_gpu = time.perf_counter() - t0
assert all(x is not None for x in b) and b[0].to_bytes() == a[0].to_bytes()
print(json.dumps({"proofs": n, "proof_bytes": len(blob), "points_per_proof": 11 + 2 * 16, "from_bytes_host_ms": round(t_host * 1e3, 1),
                  "from_bytes_batch_gpu_ms": round(t_gpu * 1e3, 1)}))
